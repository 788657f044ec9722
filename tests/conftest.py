import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with `pytest -m gpu` on the B200 box")


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI shared library, built in-tree (cross-compiles without a GPU)."""
    from ecamp_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import subprocess
        subprocess.check_call([os.path.join(ROOT, "ecamp_b200", "csrc", "build.sh")])
    return _lib
