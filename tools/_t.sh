cd /root/repo
(cd oracle && make -s)
echo "== pytest beam =="; timeout 900 python -m pytest tests -m gpu -x -q -k "beam" 2>&1 | tail -12
echo "== pytest all =="; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
