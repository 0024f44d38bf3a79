import shutil
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_sessionstart(session):
    """A fresh checkout has no built library (it is git-ignored): build it once if a compiler is here, so that the C-ABI
    tests do not depend on somebody having run __graft_entry__.build() first.  Without nvcc they fail loudly, as they should."""
    from mvster_b200 import _lib
    if not _lib.LIB_PATH.exists() and (shutil.which("nvcc") or Path("/usr/local/cuda/bin/nvcc").exists()):
        from mvster_b200 import build
        build.build_library()


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
