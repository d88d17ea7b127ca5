"""AdaLayerNormSingle: import-only in the reference's ReferenceNet path (transformer_2d.py:42, PixArt variant)."""
from torch import nn


class AdaLayerNormSingle(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("import-only in the reference's SD-1.5 ReferenceNet")
