"""The bounds-checked build (`make libqpzk_checked.so`, -DQPZK_CHECKED: every tile, level, wire and table index
computed on the device is range-checked and a violation traps) through the small-case suites. The GPU pool does
not allow compute-sanitizer; this is its replacement. The checked library is selected with QPZK_LIB in a child
process so that the parent keeps the release build."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, "qp-zk-circuits-rm_b200", "libqpzk_checked.so")


@pytest.mark.gpu
def test_small_case_suites_pass_with_every_index_checked():
    if not os.path.exists(CHECKED):
        pytest.fail("%s missing - run `python -c 'import __graft_entry__ as g; g.build()'`" % CHECKED)
    if os.environ.get("QPZK_LIB"):
        pytest.skip("already running against an alternative library")
    env = dict(os.environ, QPZK_LIB=CHECKED)
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
           os.path.join(ROOT, "tests", "test_gpu_prover.py"), os.path.join(ROOT, "tests", "test_gpu_abi_hardening.py"),
           os.path.join(ROOT, "tests", "test_gpu_sharded_proof.py"), os.path.join(ROOT, "tests", "test_sharded_commit.py"),
           os.path.join(ROOT, "tests", "test_recursion_gates.py"),
           os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-k", "not 2p18 and not large and not microbench"]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "QPZK_CHECK failed" not in r.stdout + r.stderr, tail


def test_checked_library_is_built_and_exports_the_abi():
    import ctypes
    if not os.path.exists(CHECKED):
        pytest.fail("%s missing - run build()" % CHECKED)
    L = ctypes.CDLL(CHECKED)
    for sym in ("qpzk_prove_begin", "qpzk_prove_end", "qpzk_sprove_begin", "qpzk_batch_from_values", "qpzk_merkle_new"):
        assert hasattr(L, sym)
