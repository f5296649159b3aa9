"""Sliding context windows over the frame axis (host mirror of src/pipelines/context.py)."""
from typing import Callable, List, Optional

import numpy as np


def ordered_halving(val: int) -> float:
    """Bit-reversed 64-bit fraction of ``val`` (context.py:7-12)."""
    rev = 0
    for i in range(64):
        if (val >> i) & 1:
            rev |= 1 << (63 - i)
    return rev / (1 << 64)


def uniform(step: int = ..., num_steps: Optional[int] = None, num_frames: int = ..., context_size: Optional[int] = None,
            context_stride: int = 3, context_overlap: int = 4, closed_loop: bool = True):
    """Generator of frame-index windows (context.py:15-42)."""
    if num_frames <= context_size:
        yield list(range(num_frames))
        return
    context_stride = min(context_stride, int(np.ceil(np.log2(num_frames / context_size))) + 1)
    frac = ordered_halving(step)
    for k in range(context_stride):
        cstep = 1 << k
        pad = int(round(num_frames * frac))
        first = int(frac * cstep) + pad
        last = num_frames + pad + (0 if closed_loop else -context_overlap)
        for j in range(first, last, context_size * cstep - context_overlap):
            yield [e % num_frames for e in range(j, j + context_size * cstep, cstep)]


def get_context_scheduler(name: str) -> Callable:
    if name == "uniform":
        return uniform
    raise ValueError(f"Unknown context_overlap policy {name}")


def get_total_steps(scheduler, timesteps: List[int], num_steps: Optional[int] = None, num_frames: int = ...,
                    context_size: Optional[int] = None, context_stride: int = 3, context_overlap: int = 4,
                    closed_loop: bool = True):
    return sum(len(list(scheduler(i, num_steps, num_frames, context_size, context_stride, context_overlap)))
               for i in range(len(timesteps)))
