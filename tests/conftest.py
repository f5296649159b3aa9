import os
import sys

import pytest

# The single-GPU emulation of frame shards runs kernels that wait for each other on two streams: every kernel must be
# loaded before the first such wait (lazy module loading synchronises the context).  Must be set before CUDA starts.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run on the B200 box with -m gpu")


@pytest.fixture(scope="session")
def lib_built():
    from mmgt_b200 import build
    return build.build(verbose=False)
