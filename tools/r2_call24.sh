#!/bin/bash
# Round-2 call 24 (1 GPU): configs[3] (768^2, bf16) and the 2-window clip (configs[4] shape) on the final kernels.
mkdir -p gpurun_out
timeout 240 python bench.py --steps 10 --warmup 3 --size 96 --dtype bf16 --no-cpu-baseline > gpurun_out/r2w_bench_768_bf16.json 2> gpurun_out/r2w_bench_768_bf16.err
timeout 300 python bench.py --windows 2 > gpurun_out/r2w_clip_2win.json 2> gpurun_out/r2w_clip_2win.err
python - <<'PY'
import json
for f in ("gpurun_out/r2w_bench_768_bf16.json", "gpurun_out/r2w_clip_2win.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "unit")}, "e2e", d.get("e2e", {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(f, "unreadable", e)
PY
