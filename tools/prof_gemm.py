"""Runs one GEMM shape a few times -- target of `ncu -k regex:gemm_tc`.
   M, N, K, RES (1 = +bias +residual), GEGLU via the environment; default: the L0 to_out projection."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallo_b200 import ops  # noqa: E402

M, N, K = (int(os.environ.get(k, d)) for k, d in (("M", 131072), ("N", 320), ("K", 320)))
geglu = os.environ.get("GEGLU", "0") not in ("", "0")
res = os.environ.get("RES", "1") not in ("", "0") and not geglu
dev = "cuda"
a = torch.randn(M, K, device=dev, dtype=torch.float16)
w = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.02
out = torch.empty(M, N // 2 if geglu else N, device=dev, dtype=torch.float16)
bias = torch.randn(N, device=dev, dtype=torch.float16)
r = torch.randn(M, N, device=dev, dtype=torch.float16) if res else None
for _ in range(int(os.environ.get("REPS", 3))):
    ops.gemm(a, w, out, bias=bias, residual=r, geglu=geglu)
torch.cuda.synchronize()
print("done")
