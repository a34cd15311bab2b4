#!/bin/bash
# A/B of an environment knob on the decode step: parity tests, then bench with the knob off and on, then a step trace.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(cd oracle && make -s)
KNOB=${1:-WB_FUSE_OUT}
echo "== pytest -m gpu ==" ; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
for v in 0 1; do
  echo "== bench $KNOB=$v =="
  env $KNOB=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --configs headline 2>gpurun_out/bench_$v.err | tee gpurun_out/bench_$KNOB$v.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','phase_ms')}); print(d['roofline']['us_per_launch'], d['roofline']['frac'], d['e2e']['value'])"
done
echo "== trace =="; timeout 300 python tools/trace_step.py base.en 32 24 2>&1 | tail -40 | tee gpurun_out/trace_step.txt
echo "== trace at t~215 =="; timeout 300 python tools/trace_step.py base.en 32 216 2>&1 | tail -24 | tee gpurun_out/trace_step_t215.txt
