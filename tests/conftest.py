import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU oracle check")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        # a lost fixture must not silently turn the parity suite into skips
        pytest.fail(f"golden fixture {name} missing (regenerate with oracle/make_golden.py in the build container)")
    return np.load(path)


@pytest.fixture(scope="session")
def lib_built():
    """Make sure libdiffroll_b200.so exists (nvcc cross-compiles without a GPU)."""
    from diffroll_b200 import build
    return build.build()


@pytest.fixture(scope="session", autouse=True)
def _library_present():
    """A checkout without the built artefact (the .so is git-ignored) compiles it once; an existing one is used as is."""
    from diffroll_b200 import build
    if not os.path.exists(build.LIB):
        build.build()
