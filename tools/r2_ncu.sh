#!/bin/bash
# ncu evidence for the kernels that make up the step (1 GPU).  Reports stay on the box (/tmp): gpurun_out/ only receives
# the raw-metric CSV of each capture, the stall buckets of the attention kernel and the launch lists (64 MiB limit).
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"   # not N: prof_gemm.py reads M, N, K from the environment
cap() {   # name, kernel regex, command...
  local name=$1 rx=$2; shift 2
  $NCU -k regex:$rx -s 2 -c 1 -o /tmp/r2_prof_$name -f "$@" > gpurun_out/r2_ncu_$name.log 2>&1
  ncu -i /tmp/r2_prof_$name.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${name}_raw.csv 2>/dev/null
  python tools/ncu_stalls.py /tmp/r2_prof_$name.ncu-rep > gpurun_out/r2_ncu_${name}_stalls.txt 2>&1
}
cap attn attn2_tc python tools/prof_attn.py
(export M=131072 N=320 K=320 RES=1; cap gemm_k320 gemm_tc python tools/prof_gemm.py)
(export M=131072 N=2560 K=320 GEGLU=1; cap gemm_geglu gemm_tc python tools/prof_gemm.py)
cap conv_l0 gemm_tc python tools/prof_conv.py
cap tattn tattn python tools/prof_aux.py tattn
cap xattn xattn python tools/prof_aux.py xattn
cap gn_apply gn_apply python tools/prof_aux.py gn
cap gn_stats gn_stats python tools/prof_aux.py gn
cap ln layernorm python tools/prof_aux.py ln
cp /tmp/r2_prof_attn.ncu-rep gpurun_out/r2_prof_attn.ncu-rep
L="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
$L --log-file gpurun_out/r2_launches_one_step.csv python bench.py --profiler-range > gpurun_out/r2_ncu_bench.log 2>&1
$L --log-file gpurun_out/r2_launches_one_step_shard8.csv python bench.py --profiler-range --emulate-shard 8 > gpurun_out/r2_ncu_bench_shard8.log 2>&1
du -sh gpurun_out
python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2_smoke.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.log
tail -3 gpurun_out/r2_smoke.log
