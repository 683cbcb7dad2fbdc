"""Uploads from pageable host memory go through the context's pinned staging chunks filled by several host
threads (h2d_copy in csrc/qpzk.cu). The staged path must move exactly the bytes the driver's path moves: odd
lengths, lengths around the 8 MB chunk, several uploads in a row on one context, and a commit whose values come
from a pageable array against the same commit from pinned memory. Run in a subprocess because the knobs
(QPZK_H2D_THREADS, QPZK_H2D_MIN_MB) are read once per process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

SCRIPT = r"""
import numpy as np, ctypes, sys
sys.path.insert(0, 'tests'); sys.path.insert(0, 'qp-zk-circuits-rm_b200')
import qpzk
from helpers import rand_felts
ctx = qpzk.Context(0)
rng = np.random.default_rng(7)
CH = 8 << 20
for nbytes in (1, 63, 64, 65, 4097, CH - 8, CH, CH + 8, 2 * CH + 24, 3 * CH + 12345, 5 * CH + 1):
    src = rng.integers(0, 256, nbytes, dtype=np.uint8)          # pageable
    dev = ctx.dev_alloc(nbytes)
    ctx.h2d(dev, src)
    back = np.zeros(nbytes, np.uint8)
    ctx.d2h(back, dev)
    assert np.array_equal(src, back), nbytes
    ctx.dev_free(dev)
# a commit from pageable values == the same commit from pinned memory (rows of 2^15 x 40 = 10.5 MB > 1 chunk)
vals = rand_felts(np.random.default_rng(3), (40, 1 << 15))
pin = qpzk.PinnedBuffer(vals.shape)
pin.array[...] = vals
a = qpzk.PolynomialBatch.from_values(ctx, vals, 3, 4)
b = qpzk.PolynomialBatch.from_values(ctx, pin.array, 3, 4)
assert np.array_equal(a.cap, b.cap)
assert np.array_equal(a.polynomials, b.polynomials)
a.free(); b.free(); pin.free(); ctx.close()
print("staged-upload-ok")
"""


@pytest.mark.parametrize("threads", ["3", "4", "1"])
def test_staged_upload_moves_the_same_bytes(threads):
    env = dict(os.environ, QPZK_H2D_MIN_MB="0", QPZK_H2D_THREADS=threads)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", SCRIPT], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "staged-upload-ok" in r.stdout
