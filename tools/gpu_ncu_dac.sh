#!/bin/bash
# ncu --set full captures of the DAC encoder kernels (2 x 60 s of audio: same tile shapes as the dump_tokens batch).
# The plain run of the same command goes first; nothing printed under ncu is a timing.
mkdir -p gpurun_out
export DAC_B=2
CMD="python tools/bringup_ops.py dacenc"
$CMD > gpurun_out/plain_dacenc.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_dacenc.log; exit 1; }
tail -32 gpurun_out/plain_dacenc.log
# the first forward of the tool launches: conv0, 3 x resunit64, conv, 3 x resunit<128>, conv, 6 convs (256 ch), conv, 6 convs (512), conv, conv
ncu --set full --clock-control none --import-source on -k regex:dac_ -c 26 -o gpurun_out/dac_enc -f $CMD > gpurun_out/ncu_dac.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_dac.log
ls -la gpurun_out/dac_enc.ncu-rep
