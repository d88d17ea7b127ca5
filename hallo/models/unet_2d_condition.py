"""hallo.models.unet_2d_condition (scripts/inference.py:45) -> the engine-backed ReferenceNet
(hallo/models/unet_2d_condition.py).  Remove this file from the overlay to keep the reference's PyTorch ReferenceNet:
the write-mode ReferenceAttentionControl works with either."""
from hallo_b200.models.unet_2d_condition import UNet2DConditionModel, UNet2DConditionOutput  # noqa: F401
