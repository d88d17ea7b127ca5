"""Runs one of the non-GEMM kernels a few times at its L0 shape -- target of `ncu -k regex:<kernel>`.
   python tools/prof_aux.py tattn | xattn | gn | ln"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallo_b200 import ops  # noqa: E402

which = sys.argv[1]
dev, dt = "cuda", torch.float16
C, L, H = 320, 4096, 8
reps = int(os.environ.get("REPS", 3))
if which == "tattn":        # temporal attention of one motion-module attention block at L0: 2 halves x 18 frames
    F = 18
    qkv = torch.randn(2 * F * L, 3 * C, device=dev, dtype=dt)
    o = torch.empty(2 * F * L, C, device=dev, dtype=dt)
    for _ in range(reps):
        ops.temporal_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, batch=2, fq=F, fk=F, tokens=L, heads=H)
elif which == "xattn":      # the three audio cross-attentions of one block at L0: 32 frames, 32 keys, 3 regions
    B = 32
    q3 = torch.randn(B * L, 3 * C, device=dev, dtype=dt)
    kva = torch.randn(B * 32, 6 * C, device=dev, dtype=dt)
    a3 = torch.empty(B * L, 3 * C, device=dev, dtype=dt)
    for _ in range(reps):
        ops.cross_attention(q3, kva[:, :C], kva[:, C:2 * C], a3, frames=B, tokens=L, heads=H, head_dim=C // H, n_keys=32,
                            kv_frame_div=1, regions=3, q_region_stride=C, kv_region_stride=2 * C, o_region_stride=C)
elif which == "gn":         # GroupNorm + SiLU of a ResNet block input at L0: 32 frames
    B = 32
    x = torch.randn(B * L, C, device=dev, dtype=dt)
    g, b = torch.ones(C, device=dev, dtype=dt), torch.zeros(C, device=dev, dtype=dt)
    out = torch.empty_like(x)
    ws = torch.empty(ops.gn_workspace_floats(B, L, 32, C), device=dev, dtype=torch.float32)
    for _ in range(reps):
        ops.groupnorm(x, g, b, out, ws, n_frames=B, hw=L, eps=1e-5, silu=True)
elif which == "ln":
    x = torch.randn(32 * L, C, device=dev, dtype=dt)
    g, b = torch.ones(C, device=dev, dtype=dt), torch.zeros(C, device=dev, dtype=dt)
    out = torch.empty_like(x)
    for _ in range(reps):
        ops.layernorm(x, g, b, out)
torch.cuda.synchronize()
print("done")
