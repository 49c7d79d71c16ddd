#!/bin/bash
# Kernel timeline of a single-utterance decode (bring-up build with -DEDM_KTRACE)
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -DEDM_KTRACE -o gpurun_out/libedm_ktrace.so edm_tts_b200/csrc/abi.cu || exit 1
timeout 300 python tools/ktrace.py gpurun_out/libedm_ktrace.so "$@" 2>&1 | tail -40 | tee gpurun_out/ktrace.txt
