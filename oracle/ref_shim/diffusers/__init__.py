"""Minimal stand-in for ``diffusers==0.24.0`` (reference pin: requirements.txt:36).

TEST INFRASTRUCTURE ONLY.  The reference's hot-path modules import diffusers, which is
not installable in this sandbox.  This package restates the handful of primitives they
use so that ``/root/reference/src/models/*.py`` can be imported
*unchanged* and serve as the ground truth that ``oracle/`` is pinned against
(see ``oracle/make_golden.py``).  Nothing in ``mmgt_b200/`` may import it.
"""
from .models.modeling_utils import ModelMixin  # noqa: F401
from .configuration_utils import ConfigMixin, register_to_config  # noqa: F401
from .pipeline_utils import DiffusionPipeline  # noqa: F401
