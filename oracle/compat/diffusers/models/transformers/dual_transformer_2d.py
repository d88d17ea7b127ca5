"""DualTransformer2DModel: import-only (dual_cross_attention is False in SD-1.5, unet_2d_blocks.py:24)."""
from torch import nn


class DualTransformer2DModel(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("dual_cross_attention is not part of the SD-1.5 ReferenceNet")
