import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "qp-zk-circuits-rm_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

# The reference's shipped fixtures for this path are committed under tests/golden/ (copied, with
# provenance, by scripts/make_golden.py) so that every box - including the GPU box, where /root/reference
# does not exist - runs the same pins.
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = {
    "wormhole/bench-data/common.bin": "bench_common.bin",
    "wormhole/bench-data/verifier.bin": "bench_verifier.bin",
    "wormhole/bench-data/proof.bin": "bench_proof.bin",
    "wormhole/aggregator/data/dummy_proof.bin": "dummy_proof.bin",
    "wormhole/aggregator/data/dummy_proof_zk.bin": "dummy_proof_zk.bin",
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref_fixture():
    def _load(rel):
        path = os.path.join(GOLDEN, GOLDEN_NAMES[rel])
        if not os.path.exists(path):
            pytest.fail("golden fixture %s missing - run scripts/make_golden.py" % path)
        with open(path, "rb") as f:
            return f.read()

    return _load
