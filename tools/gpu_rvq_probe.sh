#!/bin/bash
# RVQ search probes on the bring-up build
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -DEDM_BRINGUP -o gpurun_out/libedm_bringup.so edm_tts_b200/csrc/abi.cu || exit 1
for t in "$@"; do timeout 300 python tools/bringup_ops.py $t 2>&1 | tail -14; done
