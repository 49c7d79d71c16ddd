#!/bin/bash
# text-to-semantic decode: A/B of the working tree against build/base_src, then the per-kernel launch list of one decode
mkdir -p gpurun_out
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC"
(cd build/base_src && $NV -o ../../gpurun_out/libedm_base.so edm_tts_b200/csrc/abi.cu > /dev/null 2>&1) || exit 1
for rep in 1 2; do
  EDM_AB_OLD=1 EDM_AB_LIB=gpurun_out/libedm_base.so python tools/t2s_bench.py 2>&1 | tail -1
  python tools/t2s_bench.py 2>&1 | tail -1
done
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:edm:: -s 8403 -c 2801 --csv \
  --log-file gpurun_out/t2s_launches.csv python tools/t2s_bench.py > gpurun_out/t2s_ncu.log 2>&1
python tools/ncu_launches.py gpurun_out/t2s_launches.csv | head -24
