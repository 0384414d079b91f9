import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """The oracle is the checker: build it if the prebuilt files are missing (gcc only, a second)."""
    b = os.path.join(ROOT, "oracle", "_build")
    if not (os.path.exists(os.path.join(b, "cross_match")) and os.path.exists(os.path.join(b, "librm_oracle.so"))):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    yield
