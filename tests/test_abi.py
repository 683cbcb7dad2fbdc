"""CPU-side checks of the drop-in boundary: libqpzk.so loads, exports every symbol include/qpzk.h
declares, generates the same Poseidon tables as the oracle, and fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import qpzk
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "qpzk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qpzk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = qpzk.load_library()
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), "libqpzk.so does not export %s" % s
    # and the Python mirror declares a signature for each of them
    assert set(syms) <= set(L._declared), set(syms) - set(L._declared)


def test_rust_sys_crate_declares_every_symbol():
    """rust/qpzk-sys/src/lib.rs (the FFI crate a patched qp-plonky2 links) must declare
    exactly the header's symbols; it cannot be compiled in this image, so at least keep it in sync."""
    rs = open(os.path.join(ROOT, "rust", "qpzk-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (qpzk_[a-z0-9_]+)\(", rs))
    assert declared == set(_header_symbols()), declared ^ set(_header_symbols())


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "qpzk.h")).read()
    assert "torch" not in src and "at::" not in src and "#include <cuda" not in src


def test_product_poseidon_tables_match_oracle():
    """Two independent derivations (oracle: column-vector form; product: transposed form) agree."""
    L = qpzk.load_library()
    L.qpzk_poseidon_tables_host.restype = ctypes.c_void_p
    p = L.qpzk_poseidon_tables_host()
    n = 360 + 12 + 22 + 121 + 242 + 242
    flat = np.frombuffer((ctypes.c_uint64 * n).from_address(p), dtype=np.uint64)
    t = orc.poseidon_tables()
    off = 0
    for key, cnt in (("rc", 360), ("fast_first", 12), ("fast_rc", 22), ("fast_init", 121), ("fast_w_hat", 242),
                     ("fast_v", 242)):
        assert np.array_equal(flat[off:off + cnt], t[key].ravel()), key
        off += cnt


def test_context_creation_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(qpzk.QpzkError):
        qpzk.Context(0)
