import functools
import inspect


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    @property
    def config(self):
        return self._internal_dict

    def register_to_config(self, **kwargs):
        self._internal_dict = FrozenDict(kwargs)


def register_to_config(init):
    """diffusers.configuration_utils.register_to_config: store ctor arguments (with defaults) on self.config."""
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = {k: v.default for k, v in sig.parameters.items() if k != "self"}
        names = [k for k in sig.parameters if k != "self"]
        for n, a in zip(names, args):
            params[n] = a
        params.update(kwargs)
        init(self, *args, **kwargs)
        ConfigMixin.register_to_config(self, **params)
    return inner
