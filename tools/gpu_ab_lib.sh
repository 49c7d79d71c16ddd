#!/bin/bash
# A/B of two builds of the library on one box: the working tree against a source snapshot under build/base_src (e.g. HEAD:
#   for f in $(git ls-tree --name-only HEAD edm_tts_b200/csrc/); do git show HEAD:$f > build/base_src/$f; done).
mkdir -p gpurun_out
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC"
$NV -o gpurun_out/libedm_new.so edm_tts_b200/csrc/abi.cu || exit 1
(cd build/base_src && $NV -o ../../gpurun_out/libedm_base.so edm_tts_b200/csrc/abi.cu) || exit 1
for rep in 1 2 3; do
  for lib in base new; do
    echo -n "$lib: "; EDM_AB_LIB=gpurun_out/libedm_$lib.so python tools/ab_step.py 5 2>&1 | tail -1
  done
done
