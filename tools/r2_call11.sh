#!/bin/bash
# attn_split A/B: parity of the half-split head_dim-40 attention, then its time against the default schedule
mkdir -p gpurun_out
{
echo "== parity, HALLO_B200_ATTN_SPLIT=1"
HALLO_B200_ATTN_SPLIT=1 timeout 600 python -m pytest tests/test_attention_gpu.py -x -q -m gpu 2>&1 | tail -15
echo "== kbench attn, default"
timeout 300 python tools/kbench.py attn 2>&1 | grep -v sdpa
echo "== kbench attn, HALLO_B200_ATTN_SPLIT=1"
HALLO_B200_ATTN_SPLIT=1 timeout 300 python tools/kbench.py attn 2>&1 | grep -v sdpa
} > gpurun_out/r2k_attn_split.txt 2>&1
cat gpurun_out/r2k_attn_split.txt
