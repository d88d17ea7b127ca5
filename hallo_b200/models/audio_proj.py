"""AudioProjModel (hallo/models/audio_proj.py:40-124): wav2vec window (bz, f, 5, 12, 768) -> 32 context tokens
of 768 per frame.  Runs once per window (36 M parameters, M = 16 rows); its three Linear layers go through the
tcgen05 GEMM, ReLU / LayerNorm through PyTorch (outside the timed denoising loop)."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class AudioProjModel(nn.Module):
    def __init__(self, seq_len=5, blocks=12, channels=768, intermediate_dim=512, output_dim=768, context_tokens=32):
        super().__init__()
        self.seq_len, self.blocks, self.channels = seq_len, blocks, channels
        self.input_dim = seq_len * blocks * channels
        self.intermediate_dim, self.context_tokens, self.output_dim = intermediate_dim, context_tokens, output_dim
        self.proj1 = nn.Linear(self.input_dim, intermediate_dim)
        self.proj2 = nn.Linear(intermediate_dim, intermediate_dim)
        self.proj3 = nn.Linear(intermediate_dim, context_tokens * output_dim)
        self.norm = nn.LayerNorm(output_dim)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def _linear(self, x, lin):
        out = torch.empty(x.shape[0], lin.out_features, device=x.device, dtype=x.dtype)
        return ops.gemm(x.contiguous(), lin.weight, out, bias=lin.bias)

    @torch.no_grad()
    def forward(self, audio_embeds):
        bz, f = audio_embeds.shape[:2]
        x = audio_embeds.reshape(bz * f, self.input_dim)
        if x.dtype not in (torch.float16, torch.bfloat16) or not x.is_cuda:
            raise RuntimeError("hallo_b200.AudioProjModel runs in fp16/bf16 on CUDA only (no CPU path)")
        x = torch.relu(self._linear(x, self.proj1))
        x = torch.relu(self._linear(x, self.proj2))
        x = self._linear(x, self.proj3).reshape(bz * f, self.context_tokens, self.output_dim)
        x = self.norm(x)
        return x.reshape(bz, f, self.context_tokens, self.output_dim)
