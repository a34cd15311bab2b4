#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py ${1:-base.en} ${2:-32} ${3:-3} > gpurun_out/profile_step.log 2>&1
tail -2 gpurun_out/profile_step.log
