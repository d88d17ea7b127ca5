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
