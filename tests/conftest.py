import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def built_lib():
    """libmrhash_b200.so built in-tree (compiles here without a GPU)."""
    from mrhash_b200 import _capi

    if not os.path.exists(_capi.LIB_PATH):
        import subprocess

        subprocess.check_call([os.path.join(ROOT, "mrhash_b200", "build.sh")])
    return _capi.LIB_PATH
