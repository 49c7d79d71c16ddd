#!/bin/bash
# One GPU session: tests, smoke, bench, then the ncu launch list of the same bench command.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" == "ncu" ]; then
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_bf16|attention_fwd|layernorm_kernel|conv_module|conv_stream|sample_kernel|remask_kernel|inject_kernel|update_input|build_input|assemble_codes|fill_u8" -s 854 -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
  echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
fi
