#!/bin/bash
# Round-2 call 9 (1 GPU): attn_dyn A/B (parity + speed); ncu --set full of three GEMM shapes (summaries only come back).
mkdir -p gpurun_out
HALLO_B200_ATTN_DYN=1 timeout 300 python -m pytest tests/test_attention_gpu.py -q -m gpu -x -p no:cacheprovider > gpurun_out/r2j_attn_dyn_tests.log 2>&1
echo "attn_dyn tests exit $?" | tee gpurun_out/r2j_summary.txt
timeout 200 python tools/kbench.py attn > gpurun_out/r2j_kbench_attn_base.log 2>&1
HALLO_B200_ATTN_DYN=1 timeout 200 python tools/kbench.py attn > gpurun_out/r2j_kbench_attn_dyn.log 2>&1
grep -h "attn C" gpurun_out/r2j_kbench_attn_base.log gpurun_out/r2j_kbench_attn_dyn.log >> gpurun_out/r2j_summary.txt
HALLO_B200_ATTN_DYN=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_attn_dyn.json 2> gpurun_out/r2j_bench_attn_dyn.log
NCU="ncu --set full --clock-control none --import-source on"
cap() {   # name, then the command (environment for the command is exported by the caller)
  local name=$1; shift
  $NCU -k regex:gemm_tc -s 2 -c 1 -o /tmp/r2_prof_$name -f "$@" > gpurun_out/r2_ncu_$name.log 2>&1
  ncu -i /tmp/r2_prof_$name.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${name}_raw.csv 2>/dev/null
  python tools/ncu_stalls.py /tmp/r2_prof_$name.ncu-rep > gpurun_out/r2_ncu_${name}_stalls.txt 2>&1
}
( export M=131072 N=320 K=320 RES=1; cap gemm_k320 python tools/prof_gemm.py )
( export M=131072 N=2560 K=320 GEGLU=1; cap gemm_geglu python tools/prof_gemm.py )
( export M=8192 N=1280 K=1280 RES=1; cap gemm_l2 python tools/prof_gemm.py )
ls -la gpurun_out/*.csv | tail -5 >> gpurun_out/r2j_summary.txt
cat gpurun_out/r2j_summary.txt
