"""Runs the L0 3x3 conv (implicit GEMM) a few times -- target of `ncu -k regex:gemm_tc`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallo_b200 import ops  # noqa: E402

n, h, cin, cout = (int(os.environ.get(k, d)) for k, d in (("N", 32), ("H", 64), ("CIN", 320), ("COUT", 320)))
x = torch.randn(n, h, h, cin, device="cuda", dtype=torch.float16)
w = torch.randn(cout, 9 * cin, device="cuda", dtype=torch.float16) * 0.01
out = torch.empty(n * h * h, cout, device="cuda", dtype=torch.float16)
for _ in range(int(os.environ.get("REPS", 3))):
    ops.conv3x3(x, w, out)
torch.cuda.synchronize()
print("done")
