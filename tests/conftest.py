import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The formatter picks smaller chunks for small shards (spmv.cu: format_host); nearly every matrix of
# this suite is small, so the suite pins the full 8-group chunks (the layout the big configurations
# get) and test_chunk_sizes_* cover the other sizes and the automatic choice.
os.environ.setdefault("GLB_SPMV_MAX_GROUPS", "8")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The parity oracle (test infrastructure): C restatement + compiled reference if present."""
    import oracle as o
    return o


@pytest.fixture(scope="session")
def ctx():
    """One glb context on cuda:0 for the -m gpu tests; the CUDA path must be the one that runs."""
    from graphlily_b200 import capi
    if capi.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests need a B200 (there is no CPU fallback)")
    c = capi.Context(0)
    yield c
    c.close()
