#!/bin/bash
# Round-2 call 21 (2 GPUs): multi-GPU parity (peer memory + NCCL exchange) and the 2-rank bench on the final kernels
# (specialised GEMM epilogues, split-K).
mkdir -p gpurun_out
S=gpurun_out/r2u_summary.txt
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -p no:cacheprovider -rA -s > gpurun_out/r2u_tests_mgpu2.log 2>&1
echo "2-rank tests exit $?" > $S
grep -E "world|passed|failed|Error|rel L2" gpurun_out/r2u_tests_mgpu2.log | tail -8 >> $S
timeout 420 $T --nproc-per-node 2 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2u_bench_n2.json 2> gpurun_out/r2u_bench_n2.err
echo "bench n2 exit $?" >> $S
python - gpurun_out/r2u_bench_n2.json <<'PY' >> $S
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 2), "frames/s", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "shard err", d.get("sharded_vs_unsharded_rel_l2"), "clocks", d.get("clocks"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
cat $S
