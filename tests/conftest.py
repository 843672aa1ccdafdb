import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: needs the compiled unmodified reference (oracle/_ref)")


@pytest.fixture(scope="session")
def zlib():
    from zoic_b200 import capi
    return capi.load()


@pytest.fixture(scope="session")
def port():
    from oracle import port as p
    p.load()
    return p
