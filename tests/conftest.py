import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU check")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build cookietts_b200/libcwg.so if the tree is fresh (nvcc cross-compiles without a GPU)."""
    lib = os.path.join(ROOT, "cookietts_b200", "libcwg.so")
    if not os.path.exists(lib) or not os.path.exists(os.path.join(ROOT, "cookietts_b200", "_cwg_torch.so")):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "cookietts_b200", "csrc"), "-j4"],
                              stdout=subprocess.DEVNULL)
    yield
