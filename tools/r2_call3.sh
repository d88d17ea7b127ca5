#!/bin/bash
# Round-2 call 3 (2 GPUs): single-GPU checks of the scatter kernels + emulated shard first; then the real 2-rank runs.
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_aux_gpu.py tests/test_unet_gpu.py tests/test_refnet_gpu.py -q -m gpu -x -p no:cacheprovider -rA 2>&1 | grep -v "^PASSED" > gpurun_out/r2c_tests_1gpu.log
rc=${PIPESTATUS[0]}
echo "1-GPU tests exit $rc" | tee gpurun_out/r2c_summary.txt
tail -5 gpurun_out/r2c_tests_1gpu.log >> gpurun_out/r2c_summary.txt
if [ "$rc" != "0" ]; then cat gpurun_out/r2c_summary.txt; exit 0; fi
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -p no:cacheprovider -rA -s > gpurun_out/r2c_tests_mgpu.log 2>&1
echo "multi-GPU tests exit $?" | tee -a gpurun_out/r2c_summary.txt
grep -E "MGPU|rel_l2|passed|failed|Error|error" gpurun_out/r2c_tests_mgpu.log | tail -20 >> gpurun_out/r2c_summary.txt
for ex in peer nccl; do
  HALLO_B200_EXCHANGE=$ex timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 10 --warmup 3 --profile-ops > gpurun_out/r2c_bench_n2_$ex.json 2> gpurun_out/r2c_bench_n2_${ex}_ops.log
  echo "bench n2 $ex exit $?" | tee -a gpurun_out/r2c_summary.txt
  python - gpurun_out/r2c_bench_n2_$ex.json <<'PY' >> gpurun_out/r2c_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 2), "frames/s", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "shard err", d.get("sharded_vs_unsharded_rel_l2"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
  tail -25 gpurun_out/r2c_bench_n2_${ex}_ops.log >> gpurun_out/r2c_summary.txt
done
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.log
python - gpurun_out/r2c_bench_n1.json <<'PY' >> gpurun_out/r2c_summary.txt
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "ms/step", round(d["ms_per_step"], 2), "frames/s", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "attn frac", round(d["roofline"]["frac"], 3), "launches", d["launches_per_step"])
PY
cat gpurun_out/r2c_summary.txt
