#!/bin/bash
# Development: bench.py under several environment-knob settings. Each argument is a comma-separated list of VAR=VALUE.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "$@"; do
  echo "== $cfg =="
  env ${cfg//,/ } timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f  step %.2f ms  decode %.2f ms (%.1f us/step)  launches %d  xattn %.2f us frac %.3f  e2e %.0f' % (d['value'], d['ms_per_step'], d['phase_ms']['decode'], d['phase_ms']['decode']*1000/d['phase_ms']['decode_steps'], d['gpu_launches'], d['roofline']['us_per_launch'], d['roofline']['frac'], d['e2e']['value']))"
  grep "\[wb\]" gpurun_out/sweep.err | sort | uniq -c | head -4
done
