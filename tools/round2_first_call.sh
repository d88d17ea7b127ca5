#!/bin/bash
# First GPU call of a round: (1) default-path parity, (2) every staged switch on its own (parity + speed),
# (3) op breakdown of a full step and of one rank's shard of an 8-rank job.  Everything lands in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/round2_first_call.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r2_tests_default.log 2>&1
echo "default tests exit $?" | tee gpurun_out/r2_first_call_summary.txt
bash tools/run_experimental.sh > gpurun_out/r2_experimental.log 2>&1
cat gpurun_out/exp_summary.txt >> gpurun_out/r2_first_call_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --profile-ops --no-cpu-baseline > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default_ops.log
ALL="HALLO_B200_GEMM_TEPI=1 HALLO_B200_ATTN_CHUNK=1 HALLO_B200_XATTN_TC=1 HALLO_B200_TATTN_MMA=1 HALLO_B200_GN_FUSED=1"
env $ALL timeout 600 python bench.py --steps 10 --warmup 3 --profile-ops --no-cpu-baseline > gpurun_out/r2_bench_all_switches.json 2> gpurun_out/r2_bench_all_switches_ops.log
timeout 600 python bench.py --steps 10 --warmup 3 --profile-ops --no-cpu-baseline --emulate-shard 8 > gpurun_out/r2_bench_shard8.json 2> gpurun_out/r2_bench_shard8_ops.log
env $ALL HALLO_B200_GEMM_FILL=1 timeout 600 python bench.py --steps 10 --warmup 3 --profile-ops --no-cpu-baseline --emulate-shard 8 > gpurun_out/r2_bench_shard8_all_switches.json 2> gpurun_out/r2_bench_shard8_all_switches_ops.log
for f in gpurun_out/r2_bench_default.json gpurun_out/r2_bench_all_switches.json gpurun_out/r2_bench_shard8.json gpurun_out/r2_bench_shard8_all_switches.json; do
  python - "$f" <<'PY' >> gpurun_out/r2_first_call_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 2), "frames/s", round(d["value"], 3), "attn frac", round(d["roofline"]["frac"], 3))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
cat gpurun_out/r2_first_call_summary.txt
# Multi-GPU follow-up (separate call, charged N x):
#   /usr/local/graft/bin/gpurun --gpus 4 --timeout 1200 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 \
#      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 --profile-ops \
#      > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4_ops.log; \
#    HALLO_B200_MOTION_A2A=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
#      --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --profile-ops \
#      > gpurun_out/r2_bench_n4_a2a.json 2> gpurun_out/r2_bench_n4_a2a_ops.log'
# (the per-op table now includes kv_allgather_* / a2a_* / cfg_exchange_nccl rows; parity of the a2a path first:
#  HALLO_B200_MOTION_A2A=1 torchrun --nproc-per-node 2 tools/debug_gather.py)
timeout 300 python tools/ubench.py > gpurun_out/r2_ubench.log 2>&1
