#!/bin/bash
# End-of-round check on one GPU: the GPU suite, smoke(), the bench line (-> profiles/r02_bench_1gpu.json).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(cd oracle && make -s)
echo "== pytest -m gpu ==" ; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== smoke ==" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== bench ==" ; timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench --impl reference (one step) =="; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -c 700
