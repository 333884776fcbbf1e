import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SAMPLE_TIFF = os.path.join(GOLDEN, "SampleData_2Phase_stack_3d_1bit.tif")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def sample_phase():
    from oracle import oi_numpy as o
    return o.threshold(o.read_tiff_raw(SAMPLE_TIFF), 0.5)


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library (built here by nvcc cross-compilation)."""
    from openimpala_b200 import build, capi
    if not os.path.exists(capi.LIB_PATH):
        build.build_lib()
    return capi.load()
