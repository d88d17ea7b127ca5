#!/bin/bash
# Round-2 call 20 (1 GPU): compile-time specialised GEMM epilogues -- full GPU suite, GEMM micro-bench, the step
# (unsharded + as rank 0 of emulated 8- and 4-way shards), in-graph kernel table, smoke.
mkdir -p gpurun_out
S=gpurun_out/r2t_summary.txt
: > $S
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -rA --deselect tests/test_multigpu_gpu.py 2>&1 | grep -v "^PASSED" > gpurun_out/r2t_tests.log
echo "GPU suite (1 GPU) exit ${PIPESTATUS[0]}" >> $S
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2t_tests.log | tail -8 >> $S
timeout 300 python tools/kbench.py gemm > gpurun_out/r2t_kbench_gemm.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --kineto gpurun_out/r2t_kineto_n1.txt > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err
for R in 8 4; do
  timeout 300 python bench.py --emulate-shard $R --steps 20 --warmup 5 --no-cpu-baseline --kineto gpurun_out/r2t_kineto_shard$R.txt > gpurun_out/r2t_bench_shard$R.json 2> gpurun_out/r2t_bench_shard$R.err
done
for f in gpurun_out/r2t_bench_n1.json gpurun_out/r2t_bench_shard8.json gpurun_out/r2t_bench_shard4.json; do
python - $f <<'PY' >> $S
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "frames/s", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "attn frac", round(d["roofline"]["frac"], 3), d["clocks"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
python __graft_entry__.py --smoke > gpurun_out/r2t_smoke.log 2>&1; tail -3 gpurun_out/r2t_smoke.log >> $S
cat $S
