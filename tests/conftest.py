import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with -m gpu)")


@pytest.fixture(scope="session")
def vimg():
    from helpers import load_product
    return load_product()


@pytest.fixture(scope="session")
def man():
    from helpers import manifest
    return manifest()


@pytest.fixture(scope="session")
def refcpu():
    """The compiled reference (oracle/_ref); tests that need it skip when it was not built."""
    from oracle import refcpu as r
    if not r.available("parity"):
        pytest.skip("oracle/_ref/libvimg_ref_parity.so not built (make -C oracle)")
    return r
