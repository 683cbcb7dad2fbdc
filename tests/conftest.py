import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "qp-zk-circuits-rm_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

# Reference fixtures are only present in the build container; tests that need them skip elsewhere.
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref_fixture():
    def _load(rel):
        path = os.path.join(REFERENCE, rel)
        if not os.path.exists(path):
            pytest.skip("reference fixture %s not present on this box" % rel)
        with open(path, "rb") as f:
            return f.read()

    return _load
