#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(cd oracle && make -s)
for s in tiny_tc tiny_ml; do timeout 300 python tools/gpu_bringup.py $s 2>&1 | tail -4; done
echo "== bench PDL on =="; timeout 300 python tools/gpu_bringup.py bench 2>&1 | tail -3
echo "== bench PDL off =="; WB_PDL=0 timeout 300 python tools/gpu_bringup.py bench 2>&1 | tail -3
echo "== bench no graph =="; WB_NO_GRAPH=1 timeout 300 python tools/gpu_bringup.py bench 2>&1 | tail -3
echo "== pytest =="; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
