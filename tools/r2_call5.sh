#!/bin/bash
# Round-2 call 5 (2 GPUs): thread-rank sharded test on one GPU, then the real 2-rank runs (peer memory / NCCL).
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_gemm_gpu.py -q -m gpu -p no:cacheprovider -rA -k "sharded or geglu or golden" 2>&1 | grep -v "^PASSED" > gpurun_out/r2e_tests_1gpu.log
echo "1-GPU tests exit ${PIPESTATUS[0]}" | tee gpurun_out/r2e_summary.txt
grep -E "rel L2|passed|failed|FAILED" gpurun_out/r2e_tests_1gpu.log | tail -12 >> gpurun_out/r2e_summary.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -p no:cacheprovider -rA -s > gpurun_out/r2e_tests_mgpu.log 2>&1
echo "multi-GPU tests exit $?" | tee -a gpurun_out/r2e_summary.txt
grep -E "MGPU|rel_l2|passed|failed|Error|error" gpurun_out/r2e_tests_mgpu.log | tail -20 >> gpurun_out/r2e_summary.txt
for ex in peer nccl; do
  HALLO_B200_EXCHANGE=$ex timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 10 --warmup 3 --profile-ops > gpurun_out/r2e_bench_n2_$ex.json 2> gpurun_out/r2e_bench_n2_${ex}_ops.log
  echo "bench n2 $ex exit $?" | tee -a gpurun_out/r2e_summary.txt
  python - gpurun_out/r2e_bench_n2_$ex.json <<'PY' >> gpurun_out/r2e_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 2), "frames/s", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "shard err", d.get("sharded_vs_unsharded_rel_l2"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
  grep -E "peer_barrier|a2a_|groupnorm_scatter|add n" gpurun_out/r2e_bench_n2_${ex}_ops.log | head -12 >> gpurun_out/r2e_summary.txt
  tail -16 gpurun_out/r2e_bench_n2_${ex}_ops.log >> gpurun_out/r2e_summary.txt
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --windows 2 > gpurun_out/r2e_clip_n2.json 2> gpurun_out/r2e_clip_n2.log
echo "clip n2 exit $?" | tee -a gpurun_out/r2e_summary.txt
tail -2 gpurun_out/r2e_clip_n2.json >> gpurun_out/r2e_summary.txt
cat gpurun_out/r2e_summary.txt
