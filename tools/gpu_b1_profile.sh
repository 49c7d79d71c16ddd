#!/bin/bash
# Single-utterance decode: timings (small_batch_probe) and the isolated per-kernel durations of one decode (ncu launch list)
mkdir -p gpurun_out
python tools/small_batch_probe.py 2>&1 | tail -8 | tee gpurun_out/b1_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:edm:: -c 900 --csv \
  --log-file gpurun_out/b1_launches.csv python tools/b1_decode.py > gpurun_out/b1_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/b1_ncu.log
python tools/ncu_launches.py gpurun_out/b1_launches.csv | tee gpurun_out/b1_launches.txt
