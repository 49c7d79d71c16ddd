#!/bin/bash
# ncu --set full captures of single launches (one ncu session per gpurun call, smallest useful case).
# usage: tools/gpu_ncu.sh <attn|gemm> <kernel-regex> <skip> <count> <outname>
mkdir -p gpurun_out
python tools/bringup_ops.py $1 > gpurun_out/plain_$1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/$5 -f python tools/bringup_ops.py $1 > gpurun_out/ncu_$5.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/ncu_$5.log; ls -la gpurun_out/*.ncu-rep
