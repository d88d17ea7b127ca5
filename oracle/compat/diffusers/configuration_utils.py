"""ConfigMixin / register_to_config restated from diffusers 0.27.2 (configuration_utils.py)."""
import functools
import inspect
import json


class FrozenDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        kwargs.pop("kwargs", None)
        if not hasattr(self, "_internal_dict"):
            internal = kwargs
        else:
            internal = {**self._internal_dict, **kwargs}
        object.__setattr__(self, "_internal_dict", FrozenDict(internal))

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def load_config(cls, path, **kwargs):
        with open(path, "r", encoding="utf-8") as f:
            return json.load(f)

    @classmethod
    def from_config(cls, config, **kwargs):
        """Init args = signature-matching config keys overridden by kwargs; the remaining config
        keys stay reachable through ``.config`` (this is what makes ``config.center_input_sample``
        resolve in hallo/models/unet_3d.py:562)."""
        config = dict(config)
        sig = inspect.signature(cls.__init__).parameters
        expected = {k for k in sig if k not in ("self", "kwargs")}
        init_dict = {}
        for k in expected:
            if k in kwargs:
                init_dict[k] = kwargs.pop(k)
            elif k in config:
                init_dict[k] = config.pop(k)
        hidden = {k: v for k, v in config.items() if k not in init_dict and not k.startswith("_")}
        model = cls(**init_dict)
        model.register_to_config(**hidden)
        return model


def register_to_config(init):
    @functools.wraps(init)
    def inner_init(self, *args, **kwargs):
        init_kwargs = {k: v for k, v in kwargs.items() if not k.startswith("_")}
        config_init_kwargs = {k: v for k, v in kwargs.items() if k.startswith("_")}
        signature = inspect.signature(init)
        parameters = {name: p.default for i, (name, p) in enumerate(signature.parameters.items()) if i > 0}
        new_kwargs = {}
        for arg, name in zip(args, parameters.keys()):
            new_kwargs[name] = arg
        new_kwargs.update({k: init_kwargs.get(k, default) for k, default in parameters.items()
                           if k not in new_kwargs})
        new_kwargs = {**config_init_kwargs, **new_kwargs}
        getattr(self, "register_to_config")(**new_kwargs)
        init(self, *args, **init_kwargs)

    return inner_init
