cd /root/repo
(cd oracle && make -s)
echo "== pytest =="; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
for s in tiny_tc tiny_ml; do timeout 300 python tools/gpu_bringup.py $s 2>&1 | tail -1; done
