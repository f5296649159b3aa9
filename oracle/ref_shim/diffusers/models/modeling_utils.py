"""ModelMixin stand-in: nn.Module + config fall-through + device/dtype helpers."""
from functools import partial

import torch


class ModelMixin(torch.nn.Module):
    _supports_gradient_checkpointing = False

    def __getattr__(self, name):
        # diffusers lets `self.<cfg key>` resolve through `.config` (deprecated but relied upon:
        # transformer_3d.py:160 reads self.use_audio_module)
        try:
            return super().__getattr__(name)
        except AttributeError:
            d = self.__dict__.get("_internal_dict")
            if d is not None and name in d:
                return d[name]
            raise

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def enable_gradient_checkpointing(self):
        self.apply(partial(self._set_gradient_checkpointing, value=True))

    def disable_gradient_checkpointing(self):
        self.apply(partial(self._set_gradient_checkpointing, value=False))

    def _set_gradient_checkpointing(self, module, value=False):
        if hasattr(module, "gradient_checkpointing"):
            module.gradient_checkpointing = value
