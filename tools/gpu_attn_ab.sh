#!/bin/bash
# attention variants by compile-time macro: tools/gpu_attn_ab.sh "-DEDM_ATTN_PREP=0" "-DEDM_ATTN_PREP=1" ...
mkdir -p gpurun_out
i=0
for flags in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC $flags -o gpurun_out/libedm_attn_$i.so edm_tts_b200/csrc/abi.cu > /dev/null 2>&1 || { echo "build failed: $flags"; exit 1; }
  i=$((i+1))
done
for rep in 1 2; do
  i=0
  for flags in "$@"; do echo -n "$flags: "; python tools/attn_bench.py gpurun_out/libedm_attn_$i.so 2>&1 | tail -2 | tr '\n' ' '; echo; i=$((i+1)); done
done
rm -f gpurun_out/libedm_attn_*.so
