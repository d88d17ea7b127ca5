#!/bin/bash
# ncu evidence for the kernels that make up the step (1 GPU): one `--set full` capture per kernel + the launch list of
# one denoising step.  Summaries are extracted on the CPU box with tools/ncu_summary.py into profiles/r2_*.
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
$N -k regex:attn2_tc -s 2 -c 1 -o gpurun_out/r2_prof_attn -f python tools/prof_attn.py > gpurun_out/r2_ncu_attn.log 2>&1
M=131072 N=320 K=320 RES=1 $N -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/r2_prof_gemm_k320 -f python tools/prof_gemm.py > gpurun_out/r2_ncu_gemm_k320.log 2>&1
M=131072 N=2560 K=320 GEGLU=1 $N -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/r2_prof_gemm_geglu -f python tools/prof_gemm.py > gpurun_out/r2_ncu_gemm_geglu.log 2>&1
$N -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/r2_prof_conv_l0 -f python tools/prof_conv.py > gpurun_out/r2_ncu_conv.log 2>&1
$N -k regex:tattn -s 2 -c 1 -o gpurun_out/r2_prof_tattn -f python tools/prof_aux.py tattn > gpurun_out/r2_ncu_tattn.log 2>&1
$N -k regex:xattn -s 2 -c 1 -o gpurun_out/r2_prof_xattn -f python tools/prof_aux.py xattn > gpurun_out/r2_ncu_xattn.log 2>&1
$N -k regex:gn_apply -s 2 -c 1 -o gpurun_out/r2_prof_gn_apply -f python tools/prof_aux.py gn > gpurun_out/r2_ncu_gn_apply.log 2>&1
$N -k regex:gn_stats -s 2 -c 1 -o gpurun_out/r2_prof_gn_stats -f python tools/prof_aux.py gn > gpurun_out/r2_ncu_gn_stats.log 2>&1
$N -k regex:layernorm -s 2 -c 1 -o gpurun_out/r2_prof_ln -f python tools/prof_aux.py ln > gpurun_out/r2_ncu_ln.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_one_step.csv \
    python bench.py --steps 1 --warmup 0 --no-graph --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
ls -la gpurun_out/r2_prof_*.ncu-rep gpurun_out/r2_launches_one_step.csv
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_one_step_shard8.csv \
    python bench.py --steps 1 --warmup 0 --no-graph --no-cpu-baseline --emulate-shard 8 > gpurun_out/r2_ncu_bench_shard8.log 2>&1
