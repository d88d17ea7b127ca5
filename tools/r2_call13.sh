#!/bin/bash
# Round-2 call 13 (1 GPU): where the step goes inside the graph -- CUPTI kernel records of 3 replays (kernel-busy time
# vs idle gaps), unsharded and as rank 0 of an emulated 8-way shard; new split-K test expectations.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3 > gpurun_out/r2m_summary.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kineto gpurun_out/r2m_kineto_n1.txt > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err
timeout 400 python bench.py --emulate-shard 8 --steps 10 --warmup 3 --no-cpu-baseline --kineto gpurun_out/r2m_kineto_shard8.txt > gpurun_out/r2m_bench_shard8.json 2> gpurun_out/r2m_bench_shard8.err
head -3 gpurun_out/r2m_kineto_n1.txt gpurun_out/r2m_kineto_shard8.txt >> gpurun_out/r2m_summary.txt
tail -3 gpurun_out/r2m_bench_n1.err gpurun_out/r2m_bench_shard8.err >> gpurun_out/r2m_summary.txt
cat gpurun_out/r2m_summary.txt
