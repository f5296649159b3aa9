"""ConfigMixin / register_to_config stand-in (diffusers 0.24 semantics)."""
import functools
import inspect
from types import SimpleNamespace


class FrozenConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kw):
        cfg = dict(getattr(self, "_internal_dict", {}))
        cfg.update(kw)
        self._internal_dict = FrozenConfig(cfg)

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        cfg = dict(config)
        cfg.update(kwargs)
        sig = inspect.signature(cls.__init__)
        accepted = {k for k in sig.parameters if k != "self"}
        init_kw = {k: v for k, v in cfg.items() if k in accepted}
        model = cls(**init_kw)
        # undeclared keys are still registered on .config (how `center_input_sample` exists)
        extra = {k: v for k, v in cfg.items() if k not in accepted and not k.startswith("_")}
        model.register_to_config(**extra)
        return model


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = [p for p in sig.parameters.values() if p.name != "self"]
        cfg = {p.name: p.default for p in params if p.default is not inspect._empty}
        for p, a in zip(params, args):
            cfg[p.name] = a
        cfg.update(kwargs)
        init(self, *args, **kwargs)
        self.register_to_config(**cfg)

    return inner
