#!/bin/bash
# ncu evidence of the bench step (BASELINE config 2 shapes): (1) the launch list of one timed step, (2) --set full captures of the
# step's own kernels, one ncu session per kernel class. The plain run of the same command goes first; a number printed under ncu is
# never a bench value.     usage: tools/gpu_ncu_bench.sh [classes...]   (default: gemm attn ln conv)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_ncu_bench.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_ncu_bench.log; exit 1; }
tail -c 300 gpurun_out/plain_ncu_bench.log; echo
# launches per decode: 469 GEMMs, 56 attention, 238 LayerNorm, 56 conv-module, ... = 854; 3 warm-up decodes precede the timed one
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_bf16|attention_fwd|layernorm_kernel|conv_module|conv_stream|sample_kernel|remask_kernel|inject_kernel|update_input|build_input|assemble_codes|fill_u8|argmax_combine" \
    -s 2562 -c 854 --csv --log-file gpurun_out/launches_bench_step.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launch list rc=$?"
declare -A PAT=( [gemm]="gemm_bf16_tn_pair_kernel" [attn]="attention_fwd_kernel" [ln]="layernorm_kernel" [conv]="conv_stream_kernel" [rvq]="rvq_" )
declare -A SKIP=( [gemm]=1430 [attn]=170 [ln]=720 [conv]=170 [rvq]=8 )
declare -A CNT=( [gemm]=8 [attn]=1 [ln]=4 [conv]=1 [rvq]=2 )
for c in ${@:-gemm attn ln conv}; do
  ncu --set full --clock-control none --import-source on -k regex:${PAT[$c]} -s ${SKIP[$c]} -c ${CNT[$c]} -o gpurun_out/bench_$c -f $CMD > gpurun_out/ncu_bench_$c.log 2>&1
  echo "ncu $c rc=$?"; tail -2 gpurun_out/ncu_bench_$c.log
done
ls -la gpurun_out/bench_*.ncu-rep
