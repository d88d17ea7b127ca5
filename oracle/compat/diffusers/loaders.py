"""UNet2DConditionLoadersMixin (diffusers 0.27.2 loaders): LoRA / IP-adapter weight loading helpers the reference's
UNet2DConditionModel inherits but never calls at inference (hallo/models/unet_2d_condition.py:54) -- an empty base."""


class UNet2DConditionLoadersMixin:
    pass
