#!/bin/bash
# ncu --set full captures of the kernels DESIGN.md section 4 reasons about, default vs staged variant.
# One GPU, never under torchrun.  Reports land in gpurun_out/ (copy the summaries you cite into profiles/).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_ncu.sh'
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -s 1 -c 1 -f"
$NCU -k regex:attn2_tc -o gpurun_out/r2_attn_default python tools/prof_attn.py > gpurun_out/r2_ncu_attn_default.log 2>&1
HALLO_B200_ATTN_CHUNK=1 $NCU -k regex:attn2_tc -o gpurun_out/r2_attn_chunk python tools/prof_attn.py > gpurun_out/r2_ncu_attn_chunk.log 2>&1
$NCU -k regex:gemm_tc -o gpurun_out/r2_gemm_k320_default python tools/prof_gemm.py > gpurun_out/r2_ncu_gemm_default.log 2>&1
HALLO_B200_GEMM_TEPI=1 $NCU -k regex:gemm_tc -o gpurun_out/r2_gemm_k320_tepi python tools/prof_gemm.py > gpurun_out/r2_ncu_gemm_tepi.log 2>&1
GEGLU=1 N=2560 $NCU -k regex:gemm_tc -o gpurun_out/r2_gemm_geglu_default python tools/prof_gemm.py > gpurun_out/r2_ncu_geglu_default.log 2>&1
$NCU -k regex:gemm_tc -o gpurun_out/r2_conv_l0_default python tools/prof_conv.py > gpurun_out/r2_ncu_conv_default.log 2>&1
for r in gpurun_out/r2_conv_l0_default gpurun_out/r2_attn_default gpurun_out/r2_attn_chunk gpurun_out/r2_gemm_k320_default gpurun_out/r2_gemm_k320_tepi gpurun_out/r2_gemm_geglu_default; do
  [ -f $r.ncu-rep ] || continue
  ncu -i $r.ncu-rep --page raw --csv 2>/dev/null | python - "$r" <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, vals = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum", "launch__registers_per_thread", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]
print(sys.argv[1])
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print(f"   {w:80s} {vals[i]}")
PY
done > gpurun_out/r2_ncu_summary.txt 2>&1
cat gpurun_out/r2_ncu_summary.txt
