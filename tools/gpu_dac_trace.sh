#!/bin/bash
# Builds a tracing variant of the library on the GPU box (not shipped) and prints the fused-unit timeline of CTA 0.
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -DEDM_DAC_TRACE -o gpurun_out/libedm_trace.so edm_tts_b200/csrc/abi.cu || exit 1
python tools/dac_trace.py
