#!/bin/bash
# Round-2 call 7 (8 GPUs): 8-rank parity (peer memory), bench at N = 8, configs[3] at 8 GPUs, a clip at 8 GPUs.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -p no:cacheprovider -rA -s -k peer > gpurun_out/r2g_tests_mgpu8.log 2>&1
echo "8-rank parity test exit $?" | tee gpurun_out/r2g_summary.txt
grep -E "world|passed|failed|Error" gpurun_out/r2g_tests_mgpu8.log | tail -5 >> gpurun_out/r2g_summary.txt
timeout 420 $T --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 --profile-ops > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8_ops.log
echo "bench n8 exit $?" | tee -a gpurun_out/r2g_summary.txt
timeout 420 $T --nproc-per-node 8 --master-port 29553 bench.py --gpus 8 --steps 10 --warmup 3 --size 96 --dtype bf16 > gpurun_out/r2g_bench_n8_768_bf16.json 2> gpurun_out/r2g_bench_n8_768_bf16.log
echo "bench n8 768 bf16 exit $?" | tee -a gpurun_out/r2g_summary.txt
timeout 420 $T --nproc-per-node 8 --master-port 29554 bench.py --gpus 8 --windows 3 > gpurun_out/r2g_clip_n8.json 2> gpurun_out/r2g_clip_n8.log
echo "clip n8 exit $?" | tee -a gpurun_out/r2g_summary.txt
for f in gpurun_out/r2g_bench_n8.json gpurun_out/r2g_bench_n8_768_bf16.json; do
python - $f <<'PY' >> gpurun_out/r2g_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 2), "frames/s", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "shard err", d.get("sharded_vs_unsharded_rel_l2"), "clocks", d.get("clocks"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
tail -2 gpurun_out/r2g_clip_n8.json >> gpurun_out/r2g_summary.txt
tail -45 gpurun_out/r2g_bench_n8_ops.log >> gpurun_out/r2g_summary.txt
cat gpurun_out/r2g_summary.txt
