import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_files():
    import glob

    files = sorted(f for f in glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")) if not os.path.basename(f).startswith("loss_"))
    assert files, "tests/golden/*.npz missing"
    return files
