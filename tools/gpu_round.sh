#!/bin/bash
# One GPU session: tests (per-stage parity numbers -> gpurun_out/parity_r02.json), smoke, bench, then optionally the ncu launch list of
# the same bench command and of smoke().
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" == "ncu" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/smoke_launches.csv python __graft_entry__.py smoke > gpurun_out/ncu_smoke.log 2>&1
  echo "ncu smoke rc=$?"; tail -2 gpurun_out/ncu_smoke.log
fi
