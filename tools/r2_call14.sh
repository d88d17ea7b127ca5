#!/bin/bash
# Round-2 call 14 (1 GPU): programmatic dependent launch (option pdl) -- whole GPU suite with it forced on, the default
# build's GEMM / UNet parity (epilogue-vector prefetch, read-only wait on the last TMA stores), launch-latency micro-bench,
# and the step with / without pdl (unsharded and as rank 0 of an emulated 8-way shard).
mkdir -p gpurun_out
S=gpurun_out/r2n_summary.txt
: > $S
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_unet_gpu.py tests/test_pipeline_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4 >> $S
echo "defaults (pdl=0): exit ${PIPESTATUS[0]}" >> $S
HALLO_B200_PDL=1 timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -rA --deselect tests/test_multigpu_gpu.py 2>&1 | grep -v "^PASSED" > gpurun_out/r2n_tests_pdl.log
echo "GPU suite with pdl=1 exit ${PIPESTATUS[0]}" >> $S
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2n_tests_pdl.log | tail -8 >> $S
timeout 400 python tools/kbench_latency.py > gpurun_out/r2n_kbench_latency.log 2>&1
grep -E "M256 |M16384|M1024   N1280|floor" gpurun_out/r2n_kbench_latency.log >> $S
for R in 0 8; do
for K in 0 1; do
  EM=""; [ $R -gt 0 ] && EM="--emulate-shard $R"
  HALLO_B200_PDL=$K timeout 400 python bench.py $EM --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench_shard${R}_pdl$K.json 2> gpurun_out/r2n_bench_shard${R}_pdl$K.err
  python - gpurun_out/r2n_bench_shard${R}_pdl$K.json <<'PY' >> $S
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "frames/s", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), d["clocks"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
done
cat $S
