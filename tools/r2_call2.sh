#!/bin/bash
# Round-2 call 2 (1 GPU): full GPU suite with the promoted defaults, the occ2 attention variant (parity + speed), bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider -rA 2>&1 | grep -v "^PASSED" > gpurun_out/r2b_tests.log
echo "default tests exit ${PIPESTATUS[0]}" | tee gpurun_out/r2b_summary.txt
for v in "OCC2=1" "OCC2=1 POLY=3" "OCC2=1 POLY=4"; do
  envs=""; for kv in $v; do envs="$envs HALLO_B200_ATTN_$kv"; done
  tag=$(echo $v | tr ' =' '__')
  env $envs timeout 600 python -m pytest tests/test_attention_gpu.py -q -m gpu -x -p no:cacheprovider > gpurun_out/r2b_attn_${tag}_tests.log 2>&1
  echo "attn $v tests exit $?" | tee -a gpurun_out/r2b_summary.txt
  env $envs timeout 300 python tools/kbench.py attn > gpurun_out/r2b_kbench_attn_${tag}.log 2>&1
done
timeout 300 python tools/kbench.py attn > gpurun_out/r2b_kbench_attn_base.log 2>&1
HALLO_B200_ATTN_POLY=4 timeout 300 python tools/kbench.py attn > gpurun_out/r2b_kbench_attn_POLY_4.log 2>&1
grep -h "L4096" gpurun_out/r2b_kbench_attn_*.log | sed 's/^/  /' >> gpurun_out/r2b_summary.txt
for f in gpurun_out/r2b_kbench_attn_*.log; do echo "== $f"; cat $f; done >> gpurun_out/r2b_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --profile-ops --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench_ops.log
tail -20 gpurun_out/r2b_bench_ops.log >> gpurun_out/r2b_summary.txt
cat gpurun_out/r2b_summary.txt
