#!/bin/bash
# Round-2 call 4 (1 GPU): full GPU suite (new: ReferenceNet, emulated shard, scatter kernels, clip driver, AudioProj
# fixture), bench at configs[1] and configs[3] (96x96 bf16), a short clip run (configs[4] on one GPU).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -rA 2>&1 | grep -v "^PASSED" > gpurun_out/r2d_tests.log
echo "GPU tests exit ${PIPESTATUS[0]}" | tee gpurun_out/r2d_summary.txt
grep -E "rel L2|passed|failed|FAILED|ERROR" gpurun_out/r2d_tests.log | tail -60 >> gpurun_out/r2d_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-ops > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench_ops.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --size 96 --dtype bf16 > gpurun_out/r2d_bench_768_bf16.json 2> gpurun_out/r2d_bench_768_bf16.log
timeout 900 python bench.py --windows 2 > gpurun_out/r2d_clip_2win.json 2> gpurun_out/r2d_clip_2win.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --emulate-shard 8 --profile-ops > gpurun_out/r2d_bench_shard8.json 2> gpurun_out/r2d_bench_shard8_ops.log
HALLO_B200_ATTN_OCC2=2 timeout 300 python -m pytest tests/test_attention_gpu.py -q -m gpu -x -p no:cacheprovider > gpurun_out/r2d_attn_stg4_tests.log 2>&1
echo "attn occ2=2 (4-stage ring) tests exit $?" >> gpurun_out/r2d_summary.txt
timeout 300 python tools/kbench.py attn gemm > gpurun_out/r2d_kbench.log 2>&1
HALLO_B200_ATTN_OCC2=2 timeout 300 python tools/kbench.py attn > gpurun_out/r2d_kbench_attn_stg4.log 2>&1
grep -h "L4096" gpurun_out/r2d_kbench.log gpurun_out/r2d_kbench_attn_stg4.log >> gpurun_out/r2d_summary.txt
grep -h "geglu" -A1 gpurun_out/r2d_kbench.log >> gpurun_out/r2d_summary.txt
for f in gpurun_out/r2d_bench.json gpurun_out/r2d_bench_768_bf16.json gpurun_out/r2d_bench_shard8.json; do
python - $f <<'PY' >> gpurun_out/r2d_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 2), "frames/s", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "attn frac", round(d["roofline"]["frac"], 3), "launches", d["launches_per_step"], "whole-step frac", round(d["roofline"]["whole_step"]["frac"], 3))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
tail -3 gpurun_out/r2d_clip_2win.json gpurun_out/r2d_clip_2win.log >> gpurun_out/r2d_summary.txt
tail -22 gpurun_out/r2d_bench_ops.log >> gpurun_out/r2d_summary.txt
tail -40 gpurun_out/r2d_bench_shard8_ops.log >> gpurun_out/r2d_summary.txt
cat gpurun_out/r2d_summary.txt
