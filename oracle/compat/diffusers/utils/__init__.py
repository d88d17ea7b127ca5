import logging as _pylogging

SAFETENSORS_WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"
WEIGHTS_NAME = "diffusion_pytorch_model.bin"


class BaseOutput:
    """Reference subclasses are @dataclass'es accessed by attribute (.sample)."""

    def __getitem__(self, k):
        if isinstance(k, str):
            return getattr(self, k)
        return tuple(getattr(self, f) for f in self.__dataclass_fields__)[k]


class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


logging = _Logging()

# the non-PEFT branch of the reference's 2D code (LoRACompatible* layers, explicit `scale` arguments): same arithmetic
USE_PEFT_BACKEND = False


def deprecate(*args, **kwargs):
    return None


def scale_lora_layers(model, weight):
    return None


def unscale_lora_layers(model, weight=None):
    return None


def is_torch_version(operation: str, version: str) -> bool:
    import operator
    import torch
    from packaging.version import parse
    ops = {">": operator.gt, ">=": operator.ge, "==": operator.eq, "!=": operator.ne, "<=": operator.le, "<": operator.lt}
    return ops[operation](parse(parse(torch.__version__).base_version), parse(version))
