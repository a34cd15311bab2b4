#!/bin/bash
cd "$(dirname "$0")/.." 2>/dev/null || true
mkdir -p gpurun_out
(cd oracle && make -s)
echo "== pytest -m gpu ==" ; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== smoke ==" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== memcheck: log-mel tests =="; timeout 600 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "logmel or legacy or spectrogram or stream_log" > gpurun_out/sanitizer_memcheck_logmel.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck_logmel.log
echo "== ncu launch list =="
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py base.en 32 4 > gpurun_out/profile_step.log 2>&1
tail -2 gpurun_out/profile_step.log
echo "== ncu full: log-mel =="
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"logmel" -c 2 -f -o gpurun_out/prof_logmel python tools/profile_step.py base.en 32 1 > gpurun_out/prof5.log 2>&1; tail -2 gpurun_out/prof5.log
