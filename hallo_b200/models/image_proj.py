"""ImageProjModel (hallo/models/image_proj.py:23-76): face embedding (bs, 512) -> 4 context tokens of 768, LayerNorm.
One 512 x 3072 GEMM per clip -- plain PyTorch, outside the hot path; same state-dict keys (proj, norm)."""
from __future__ import annotations

import torch
from torch import nn


class ImageProjModel(nn.Module):
    def __init__(self, cross_attention_dim=1024, clip_embeddings_dim=1024, clip_extra_context_tokens=4):
        super().__init__()
        self.generator = None
        self.cross_attention_dim = cross_attention_dim
        self.clip_extra_context_tokens = clip_extra_context_tokens
        self.proj = nn.Linear(clip_embeddings_dim, clip_extra_context_tokens * cross_attention_dim)
        self.norm = nn.LayerNorm(cross_attention_dim)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @torch.no_grad()
    def forward(self, image_embeds):
        t = self.proj(image_embeds).reshape(-1, self.clip_extra_context_tokens, self.cross_attention_dim)
        return self.norm(t)
