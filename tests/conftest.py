import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Native pieces present (build once per session when stale)."""
    import subprocess
    pkg = os.path.join(ROOT, "nemotron-speech.cpp_b200")
    if not os.path.exists(os.path.join(pkg, "libnsb200.so")):
        subprocess.check_call(["make", "-C", pkg, "-j8", "all"], stdout=subprocess.DEVNULL)
    import oracle as O
    O.build()
    return True
