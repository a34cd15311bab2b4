#!/bin/bash
# Development: L2 hit rate and DRAM bytes of the decode kernels as they run INSIDE a step (no cache flush, single-pass
# metrics so that the kernel is not replayed): do the weights survive the 590 MB cross K/V stream in L2?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for pol in 1 0; do
WB_L2_POLICY=$pol timeout 600 ncu --profile-from-start off --cache-control none --clock-control none \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum \
  -k regex:"self_block|post_block|logits_tc|attn_decode_head|step_finish" --csv --log-file gpurun_out/l2_instep_policy$pol.csv \
  python tools/profile_step.py base.en 32 4 > gpurun_out/l2_instep.log 2>&1
tail -1 gpurun_out/l2_instep.log
done
