#!/bin/bash
# A/B of bring-up switches on one box: tools/gpu_ab.sh "EDM_PDL=0" "EDM_PDL=1" ...  (each setting is run twice, interleaved)
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -DEDM_BRINGUP -o gpurun_out/libedm_bringup.so edm_tts_b200/csrc/abi.cu || exit 1
for rep in 1 2; do
  for setting in "$@"; do
    env $setting python tools/ab_step.py 5 2>&1 | tail -1
  done
done
