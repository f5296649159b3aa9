"""mmgt_b200 -- B200-native (sm_100a) implementation of MMGT's stage-2 denoising hot path.

Python host modules mirror the reference's interfaces (``UNet3DConditionModel.forward``,
``Pose2VideoPipeline.__call__``, ``ReferenceAttentionControl``) and dispatch every operator to the
hand-written CUDA kernels in ``libmmgt_b200.so`` (C ABI: include/mmgt_b200.h).  There is no CPU path.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
