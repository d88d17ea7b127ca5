#!/bin/bash
# ncu evidence for the GEMM kernel after the compile-time epilogue specialisation (1 GPU): same shapes and commands as
# tools/r2_ncu.sh, summaries under gpurun_out/r2v_*.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"   # not N: the shape variables below are M, N, K
cap() {   # name, kernel regex, command...
  local name=$1 rx=$2; shift 2
  $NCU -k regex:$rx -s 2 -c 1 -o /tmp/r2v_prof_$name -f "$@" > gpurun_out/r2v_ncu_$name.log 2>&1
  ncu -i /tmp/r2v_prof_$name.ncu-rep --page raw --csv > gpurun_out/r2v_ncu_${name}_raw.csv 2>/dev/null
  python tools/ncu_stalls.py /tmp/r2v_prof_$name.ncu-rep > gpurun_out/r2v_ncu_${name}_stalls.txt 2>&1
}
(export M=131072 N=320 K=320 RES=1; cap gemm_k320 gemm_tc python tools/prof_gemm.py)
(export M=131072 N=2560 K=320 GEGLU=1; cap gemm_geglu gemm_tc python tools/prof_gemm.py)
(export M=131072 N=960 K=320 RES=0; cap gemm_qkv gemm_tc python tools/prof_gemm.py)
du -sh gpurun_out
tail -2 gpurun_out/r2v_ncu_gemm_k320.log
