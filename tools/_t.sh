cd /root/repo
(cd oracle && make -s)
python tools/det.py
echo "== pytest =="; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python tools/gpu_bringup.py bench 2>&1 | tail -1 | grep -o "| [0-9.]*ms timings.*" | tail -c 110
python tools/trace_step.py base.en 32 24 2>&1 | tail -4
