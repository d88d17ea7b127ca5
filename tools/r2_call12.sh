#!/bin/bash
# Round-2 call 12 (1 GPU): split-K GEMM (option gemm_splitk): parity, then whole-suite with the option forced on, then
# the per-rank step of an 8-way frame shard (emulated on one GPU) and the unsharded step, each with and without it.
mkdir -p gpurun_out
S=gpurun_out/r2l_summary.txt
: > $S
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -p no:cacheprovider -rA -x 2>&1 | grep -v "^PASSED" > gpurun_out/r2l_gemm_tests.log
echo "gemm tests exit ${PIPESTATUS[0]}" >> $S
grep -E "S=|passed|failed|FAILED|ERROR" gpurun_out/r2l_gemm_tests.log | tail -60 >> $S
HALLO_B200_GEMM_SPLITK=1 timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -rA --deselect tests/test_multigpu_gpu.py 2>&1 | grep -v "^PASSED" > gpurun_out/r2l_tests_splitk.log
echo "GPU suite with gemm_splitk=1 exit ${PIPESTATUS[0]}" >> $S
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2l_tests_splitk.log | tail -8 >> $S
for R in 8 4; do
for K in 0 1; do
  HALLO_B200_GEMM_SPLITK=$K timeout 300 python bench.py --emulate-shard $R --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_bench_shard${R}_splitk$K.json 2> gpurun_out/r2l_bench_shard${R}_splitk$K.err
  python - gpurun_out/r2l_bench_shard${R}_splitk$K.json <<'PY' >> $S
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "launches", d.get("launches_per_step"), d["clocks"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
done
for K in 0 1; do
  HALLO_B200_GEMM_SPLITK=$K timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_bench_n1_splitk$K.json 2> gpurun_out/r2l_bench_n1_splitk$K.err
  python - gpurun_out/r2l_bench_n1_splitk$K.json <<'PY' >> $S
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "frames/s", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), d["clocks"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
cat $S
