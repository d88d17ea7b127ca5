#!/bin/bash
# Parity + speed of the opt-in kernels (include/hallo_b200.h: hallo_b200_set_option), one process per switch so a
# trapping kernel only takes its own run down.  Run on the GPU box:  bash tools/run_experimental.sh
# Logs land in gpurun_out/exp_*.log; promote a switch to default only when its parity run is green AND faster.
mkdir -p gpurun_out
: > gpurun_out/exp_summary.txt
run() {   # name, env assignment, test files...
  local name=$1 envs=$2; shift 2
  env $envs timeout 900 python -m pytest "$@" -q -m gpu -x -p no:cacheprovider -k "not oracle_port and not bf16" > gpurun_out/exp_${name}_tests.log 2>&1
  echo "$name tests exit $?" | tee -a gpurun_out/exp_summary.txt
}
run gemm_tepi  HALLO_B200_GEMM_TEPI=1  tests/test_gemm_gpu.py tests/test_aux_gpu.py tests/test_unet_gpu.py
run attn_chunk HALLO_B200_ATTN_CHUNK=1 tests/test_attention_gpu.py tests/test_unet_gpu.py
run attn_chunk2 HALLO_B200_ATTN_CHUNK=2 tests/test_attention_gpu.py tests/test_unet_gpu.py
run attn_v3    HALLO_B200_ATTN_V3=1    tests/test_attention_gpu.py tests/test_unet_gpu.py
run xattn_tc   HALLO_B200_XATTN_TC=1   tests/test_aux_gpu.py tests/test_unet_gpu.py
run gemm_fill  HALLO_B200_GEMM_FILL=1  tests/test_gemm_gpu.py tests/test_aux_gpu.py tests/test_unet_gpu.py
run gn_fused   HALLO_B200_GN_FUSED=1   tests/test_aux_gpu.py tests/test_unet_gpu.py
run tattn_mma  HALLO_B200_TATTN_MMA=1  tests/test_aux_gpu.py tests/test_unet_gpu.py tests/test_pipeline_gpu.py
# speed: baseline first, then each switch
timeout 600 python tools/kbench.py gemm conv attn > gpurun_out/exp_kbench_base.log 2>&1
HALLO_B200_GEMM_TEPI=1 timeout 600 python tools/kbench.py gemm conv > gpurun_out/exp_kbench_gemm_tepi.log 2>&1
HALLO_B200_ATTN_CHUNK=1 timeout 600 python tools/kbench.py attn > gpurun_out/exp_kbench_attn_chunk.log 2>&1
HALLO_B200_ATTN_CHUNK=1 HALLO_B200_ATTN_POLY=4 timeout 600 python tools/kbench.py attn > gpurun_out/exp_kbench_attn_chunk_poly4.log 2>&1
HALLO_B200_ATTN_CHUNK=2 timeout 600 python tools/kbench.py attn > gpurun_out/exp_kbench_attn_chunk2.log 2>&1
HALLO_B200_ATTN_V3=1 timeout 600 python tools/kbench.py attn > gpurun_out/exp_kbench_attn_v3.log 2>&1
HALLO_B200_ATTN_V3=1 HALLO_B200_ATTN_POLY=4 timeout 600 python tools/kbench.py attn > gpurun_out/exp_kbench_attn_v3_poly4.log 2>&1
HALLO_B200_ATTN_V3=1 HALLO_B200_ATTN_POLY=3 timeout 600 python tools/kbench.py attn > gpurun_out/exp_kbench_attn_v3_poly3.log 2>&1
HALLO_B200_ATTN_CHUNK=2 HALLO_B200_ATTN_POLY=4 timeout 600 python tools/kbench.py attn > gpurun_out/exp_kbench_attn_chunk2_poly4.log 2>&1
tail -n 30 gpurun_out/exp_kbench_*.log >> gpurun_out/exp_summary.txt
cat gpurun_out/exp_summary.txt
