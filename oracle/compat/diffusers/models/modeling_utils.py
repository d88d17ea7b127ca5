"""ModelMixin restated from diffusers 0.27.2: nn.Module + device/dtype + config attribute fallback
+ recursive gradient-checkpointing switch.  from_config leaves the module in train() mode."""
from functools import partial

import torch

from ..configuration_utils import ConfigMixin


class ModelMixin(torch.nn.Module, ConfigMixin):
    _supports_gradient_checkpointing = False

    def __getattr__(self, name):
        # diffusers: direct attribute access falls back to the registered config (deprecated path used
        # by hallo/animate/face_animate.py:315 `denoising_unet.in_channels`)
        is_in_config = "_internal_dict" in self.__dict__ and name in self.__dict__["_internal_dict"]
        if is_in_config and name not in self.__dict__:
            return self._internal_dict[name]
        return super().__getattr__(name)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def enable_gradient_checkpointing(self):
        if not self._supports_gradient_checkpointing:
            raise ValueError(f"{self.__class__.__name__} does not support gradient checkpointing.")
        self.apply(partial(self._set_gradient_checkpointing, value=True))

    def disable_gradient_checkpointing(self):
        if self._supports_gradient_checkpointing:
            self.apply(partial(self._set_gradient_checkpointing, value=False))
