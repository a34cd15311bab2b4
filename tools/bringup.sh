#!/bin/bash
# Runs each bring-up section in its own process with its own timeout, so a hung kernel cannot take the whole call down.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
(cd oracle && make -s)
for s in "$@"; do
  echo "=== $s ===" | tee -a gpurun_out/bringup.log
  timeout 420 python tools/gpu_bringup.py $s 2>&1 | tail -40 | tee -a gpurun_out/bringup.log
done
