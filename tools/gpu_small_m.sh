#!/bin/bash
# Small-M GEMM changes: unit tests, the whole GPU suite, then same-box A/B of the decode timings against build/base_src
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x 2>&1 | tail -6
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 900 bash tools/gpu_ab_lib.sh 2>&1 | tee gpurun_out/ab_small_m.txt
