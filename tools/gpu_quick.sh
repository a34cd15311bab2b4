#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(cd oracle && make -s)
for s in tiny_tc tiny_ml; do timeout 300 python tools/gpu_bringup.py $s 2>&1 | tail -1; done
echo "== bench default =="; timeout 300 python tools/gpu_bringup.py bench 2>&1 | tail -1
echo "== bench PDL off =="; WB_PDL=0 timeout 300 python tools/gpu_bringup.py bench 2>&1 | tail -1
echo "== bench attn=cross =="; WB_ATTN_IMPL=cross timeout 300 python tools/gpu_bringup.py bench 2>&1 | tail -1
echo "== bench attn=reg =="; WB_ATTN_IMPL=reg timeout 300 python tools/gpu_bringup.py bench 2>&1 | tail -1
echo "== pytest =="; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== bench.py =="; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_quick.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','phase_ms','roofline')}); print(d['e2e'])"
bash tools/gpu_launchlist.sh
