import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _cuda_device_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: on a machine without one a plain `pytest tests` skips them instead of failing."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device (the gpu-marked tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """libpbr.so is a build artefact (git-ignored): build it in-tree if it is missing or stale (nvcc cross-compiles
    for sm_100a without a GPU).  The product itself never builds implicitly; this is for a fresh checkout under pytest."""
    import shutil
    from pyremot_b200 import build
    if build.needs_build() and not (os.path.exists(os.path.join(build.CUDA_HOME, "bin", "nvcc")) or shutil.which("nvcc")):
        pytest.skip("libpbr.so is not built and nvcc is not available")
    build.build_lib()
