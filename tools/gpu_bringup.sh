#!/bin/bash
# Runs every operator bring-up test in its own process (a trapped kernel poisons only its own CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/bringup_smi.txt 2>&1
for t in ${@:-gemm attn ln conv sample remask rvqtc}; do
  echo "=== $t ===" | tee -a gpurun_out/bringup.log
  timeout 300 python tools/bringup_ops.py $t >> gpurun_out/bringup.log 2>&1
  echo "exit=$?" | tee -a gpurun_out/bringup.log
done
tail -n 120 gpurun_out/bringup.log
