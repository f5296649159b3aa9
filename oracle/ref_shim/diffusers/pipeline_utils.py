import torch
from tqdm import tqdm


class DiffusionPipeline:
    def register_modules(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def progress_bar(self, iterable=None, total=None):
        return tqdm(iterable, total=total, disable=True)

    @property
    def device(self):
        return torch.device("cpu")

    def to(self, *a, **k):
        return self
