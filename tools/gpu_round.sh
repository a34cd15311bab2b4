#!/bin/bash
# One GPU call: parity tests, smoke, bench, ncu launch list and full captures of the top kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(cd oracle && make -s)
echo "== pytest -m gpu ==" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke ==" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench ==" ; timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
if [ "$1" != "noprof" ]; then
echo "== ncu launch list ==" 
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py base.en 32 4 > gpurun_out/profile_step.log 2>&1
tail -2 gpurun_out/profile_step.log
echo "== ncu full: attn_decode ==" 
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_decode_head -s 2 -c 2 -f -o gpurun_out/prof_attn_decode python tools/profile_step.py base.en 32 2 > gpurun_out/prof1.log 2>&1; tail -2 gpurun_out/prof1.log
echo "== ncu full: gemm_tc ==" 
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 4 -f -o gpurun_out/prof_gemm_tc python tools/profile_step.py base.en 32 1 > gpurun_out/prof2.log 2>&1; tail -2 gpurun_out/prof2.log
echo "== ncu full: decoder block kernels, logits, finish ==" 
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"self_block|post_block|logits_tc|step_finish" -s 4 -c 8 -f -o gpurun_out/prof_blocks python tools/profile_step.py base.en 32 2 > gpurun_out/prof3.log 2>&1; tail -2 gpurun_out/prof3.log
echo "== ncu full: encoder attention, layernorm, log-mel ==" 
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"encoder_attention|layernorm|logmel" -c 5 -f -o gpurun_out/prof_misc python tools/profile_step.py base.en 32 1 > gpurun_out/prof4.log 2>&1; tail -2 gpurun_out/prof4.log
echo "== trace ==" 
timeout 300 python tools/trace_step.py base.en 32 24 2>&1 | tail -22 > gpurun_out/trace_step.txt; tail -3 gpurun_out/trace_step.txt
fi
ls -la gpurun_out
