#!/bin/bash
# Round-2 call 10 (1 GPU): full GPU suite on the final kernels (deeper residual prefetch, cheaper SiLU), GEMM micro-bench, bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -rA 2>&1 | grep -v "^PASSED" > gpurun_out/r2k_tests.log
echo "GPU tests exit ${PIPESTATUS[0]}" | tee gpurun_out/r2k_summary.txt
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2k_tests.log | tail -8 >> gpurun_out/r2k_summary.txt
timeout 300 python tools/kbench.py gemm > gpurun_out/r2k_kbench_gemm.log 2>&1
grep -E "residual|geglu" gpurun_out/r2k_kbench_gemm.log >> gpurun_out/r2k_summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-ops > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench_ops.log
python - gpurun_out/r2k_bench.json <<'PY' >> gpurun_out/r2k_summary.txt
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "ms/step", round(d["ms_per_step"], 2), "frames/s", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "setup", round(d["e2e"]["window_setup_ms"], 1), "attn frac", round(d["roofline"]["frac"], 3), "launches", d["launches_per_step"], d["clocks"])
PY
tail -16 gpurun_out/r2k_bench_ops.log >> gpurun_out/r2k_summary.txt
python __graft_entry__.py --smoke > gpurun_out/r2k_smoke.log 2>&1; tail -3 gpurun_out/r2k_smoke.log >> gpurun_out/r2k_summary.txt
cat gpurun_out/r2k_summary.txt
