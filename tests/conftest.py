import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.lib()
    return O


@pytest.fixture(scope="session")
def ossl():
    import ctypes as C

    from oracle import oracle as O

    O.build()
    path = os.path.join(ROOT, "oracle", "libosslref.so")
    if not os.path.exists(path):
        pytest.skip("libcrypto cross-check helper not built")
    return C.CDLL(path)
