"""LoRACompatibleConv / LoRACompatibleLinear (diffusers 0.27.2 models/lora.py) without an attached LoRA layer: plain
Conv2d / Linear that tolerate the extra `scale` argument the non-PEFT call sites pass."""
from torch import nn


class LoRACompatibleConv(nn.Conv2d):
    def __init__(self, *args, lora_layer=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_layer = lora_layer

    def forward(self, hidden_states, scale: float = 1.0):
        return super().forward(hidden_states)


class LoRACompatibleLinear(nn.Linear):
    def __init__(self, *args, lora_layer=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_layer = lora_layer

    def forward(self, hidden_states, scale: float = 1.0):
        return super().forward(hidden_states)
