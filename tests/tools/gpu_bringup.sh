#!/bin/bash
# Runs every bring-up group under its own timeout so that a hung kernel cannot eat the whole GPU call.
mkdir -p gpurun_out
LOG=gpurun_out/bringup.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
for g in ${@:-gemm attn forward sampler perf}; do
  timeout 300 python tests/tools/gpu_bringup.py $g >> $LOG 2>&1
  echo "[group $g exit $?]" >> $LOG
done
tail -n 150 $LOG
