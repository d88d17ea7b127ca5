#!/bin/bash
# Round-2 call 15 (1 GPU): split-K threshold A/B on the per-rank step of emulated 8- and 4-way shards, and the in-graph
# per-launch durations (CUPTI) with the op list to match them to shapes.
mkdir -p gpurun_out
S=gpurun_out/r2o_summary.txt
: > $S
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -2 >> $S
for R in 8 4; do
for K in 0 16 32 48 80; do
  HALLO_B200_GEMM_SPLITK=$K timeout 300 python bench.py --emulate-shard $R --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2o_bench_shard${R}_splitk$K.json 2> gpurun_out/r2o_bench_shard${R}_splitk$K.err
  python - gpurun_out/r2o_bench_shard${R}_splitk$K.json <<'PY' >> $S
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 3), d["clocks"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
done
timeout 400 python bench.py --emulate-shard 8 --steps 5 --warmup 3 --no-cpu-baseline --kineto gpurun_out/r2o_kineto_shard8.txt > /dev/null 2> gpurun_out/r2o_k8.err
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kineto gpurun_out/r2o_kineto_n1.txt > /dev/null 2> gpurun_out/r2o_k1.err
wc -l gpurun_out/r2o_kineto_*.seq gpurun_out/r2o_kineto_*.ops >> $S
cat $S
