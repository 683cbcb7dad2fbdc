"""qpzk — host-side mirror of the plonky2 commitment API over the sm_100a C ABI (include/qpzk.h).

The reference is Rust and there is no Rust toolchain in this image, so this Python layer plays the
role the patched `qp-plonky2` crate plays in production: it keeps the names and
argument meaning of the functions whose bodies move to the GPU —
`PolynomialBatch::{from_values, from_coeffs, get_lde_values}`, `MerkleTree::{new, prove, cap}`,
`PoseidonHash::{hash_no_pad, two_to_one}` (qp-plonky2 1.1.1, reached from
/root/reference/wormhole/prover/src/lib.rs:233-237 and
/root/reference/wormhole/circuit/src/circuit.rs:98-108) — and calls the same C entry points a
`qpzk-sys` crate would bind. There is no CPU fallback: importing works anywhere, but creating a
`Context` without `libqpzk.so` or without a CUDA device raises.
"""
import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("QPZK_LIB") or os.path.join(_PKG, "libqpzk.so")   # QPZK_LIB: A/B builds of the kernels

P = 0xFFFFFFFF00000001
SALT_SIZE = 4
IMPORT_VERIFY = 1  # QPZK_IMPORT_VERIFY
STAGES = ("h2d", "ifft", "lde", "leaf_hash", "merkle_levels", "d2h")

_u64p = ctypes.POINTER(ctypes.c_uint64)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_vp = ctypes.c_void_p
_lib = None


class QpzkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("qpzk error %d: %s" % (code, msg))
        self.code = code


def load_library():
    """dlopen libqpzk.so and declare every symbol include/qpzk.h exports."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QpzkError(-2, "%s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    u32, u64, i = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
    sig = {
        "qpzk_ctx_create": (i, [i, u32, ctypes.POINTER(_vp)]),
        "qpzk_ctx_destroy": (None, [_vp]),
        "qpzk_last_error": (ctypes.c_char_p, []),
        "qpzk_ctx_sync": (i, [_vp]),
        "qpzk_ctx_stream": (_vp, [_vp]),
        "qpzk_ctx_stage_ms": (i, [_vp, ctypes.POINTER(ctypes.c_float)]),
        "qpzk_ctx_launch_count": (u64, [_vp]),
        "qpzk_host_alloc": (i, [ctypes.c_size_t, ctypes.POINTER(_vp)]),
        "qpzk_host_free": (None, [_vp]),
        "qpzk_dev_alloc": (i, [_vp, ctypes.c_size_t, ctypes.POINTER(_vp)]),
        "qpzk_dev_free": (None, [_vp, _vp]),
        "qpzk_memcpy_h2d": (i, [_vp, _vp, _vp, ctypes.c_size_t]),
        "qpzk_memcpy_d2h": (i, [_vp, _vp, _vp, ctypes.c_size_t]),
        "qpzk_poseidon_permute": (i, [_vp, _u64p, u64]),
        "qpzk_hash_no_pad": (i, [_vp, _u64p, u64, u32, _u64p]),
        "qpzk_two_to_one": (i, [_vp, _u64p, u64, _u64p]),
        "qpzk_merkle_new": (i, [_vp, _u64p, u64, u32, u32, ctypes.POINTER(_vp)]),
        "qpzk_tree_cap": (i, [_vp, _u64p]),
        "qpzk_tree_prove": (i, [_vp, u64, _u64p]),
        "qpzk_tree_digests": (i, [_vp, _u64p]),
        "qpzk_tree_free": (None, [_vp]),
        "qpzk_batch_from_values": (i, [_vp, _vp, u32, u32, u32, u32, _vp, u32, ctypes.POINTER(_vp)]),
        "qpzk_batch_from_coeffs": (i, [_vp, _vp, u32, u32, u32, u32, _vp, u32, ctypes.POINTER(_vp)]),
        "qpzk_batch_from_values_dev": (i, [_vp, _vp, u32, u32, u32, u32, _vp, u32, ctypes.POINTER(_vp)]),
        "qpzk_batch_from_coeffs_dev": (i, [_vp, _vp, u32, u32, u32, u32, _vp, u32, ctypes.POINTER(_vp)]),
        "qpzk_batch_from_values_shard_dev": (i, [_vp, _vp, u32, u32, u32, u32, _vp, u32, u32, u32,
                                                 ctypes.POINTER(_vp)]),
        "qpzk_batch_cap": (i, [_vp, _u64p]),
        "qpzk_batch_cap_dev": (_vp, [_vp]),
        "qpzk_batch_set_cap": (i, [_vp, _u64p]),
        "qpzk_batch_coeffs": (i, [_vp, _u64p]),
        "qpzk_batch_get_lde_rows": (i, [_vp, _u32p, u32, u32, _u64p]),
        "qpzk_batch_open": (i, [_vp, u64, _u64p, _u64p]),
        "qpzk_batch_export": (i, [_vp, _u64p, _u64p]),
        "qpzk_batch_eval_ext": (i, [_vp, _u64p, _u64p]),
        "qpzk_fri_pow": (i, [_vp, _u64p, u32, u32, ctypes.POINTER(u64)]),
        "qpzk_batch_ncols": (u32, [_vp]),
        "qpzk_batch_width": (u32, [_vp]),
        "qpzk_batch_degree_bits": (u32, [_vp]),
        "qpzk_batch_free": (None, [_vp]),
        "qpzk_measure_imad_peak": (i, [_vp, i, ctypes.POINTER(ctypes.c_double)]),
        "qpzk_batch_from_coeffs_shard_dev": (i, [_vp, _vp, u32, u32, u32, u32, _vp, u32, u32, u32,
                                                 ctypes.POINTER(_vp)]),
        "qpzk_batch_from_values_shard_dev_async": (i, [_vp, _vp, u32, u32, u32, u32, _vp, u32, u32, u32,
                                                       ctypes.POINTER(_vp)]),
        "qpzk_circuit_create": (i, [_vp, ctypes.c_char_p, ctypes.c_size_t, _u64p, _vp, ctypes.c_size_t,
                                    ctypes.POINTER(_vp)]),
        "qpzk_circuit_create_from_commitment": (i, [_vp, ctypes.c_char_p, ctypes.c_size_t, _u64p, ctypes.c_char_p, u64, u32,
                                                    ctypes.POINTER(_vp)]),
        "qpzk_circuit_commitment_size": (i, [_vp, ctypes.POINTER(u64)]),
        "qpzk_circuit_commitment_to_bytes": (i, [_vp, ctypes.c_char_p, u64]),
        "qpzk_batch_serialized_size": (i, [_vp, ctypes.POINTER(u64)]),
        "qpzk_batch_to_bytes": (i, [_vp, ctypes.c_char_p, u64]),
        "qpzk_batch_from_bytes": (i, [_vp, ctypes.c_char_p, u64, u32, ctypes.POINTER(_vp), ctypes.POINTER(u64)]),
        "qpzk_circuit_cap": (i, [_vp, _u64p, ctypes.c_size_t]),
        "qpzk_circuit_info": (i, [_vp, _u32p]),
        "qpzk_circuit_verifier_only": (ctypes.c_size_t, [_vp, ctypes.c_char_p, ctypes.c_size_t]),
        "qpzk_circuit_free": (None, [_vp]),
        "qpzk_prove": (i, [_vp, _vp, ctypes.c_size_t, _u64p, u32, _vp, _vp, _vp, ctypes.c_size_t, u32, ctypes.c_char_p,
                           ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
        "qpzk_prove_begin": (i, [_vp, _vp, ctypes.c_size_t, _u64p, u32, _vp, _vp, _vp, ctypes.c_size_t, u32]),
        "qpzk_prove_end": (i, [_vp, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
        "qpzk_sprove_begin": (i, [_vp, _vp, ctypes.c_size_t, _u64p, u32, _vp, _vp, _vp, ctypes.c_size_t, u32, u32, u32,
                                  ctypes.POINTER(_vp)]),
        "qpzk_sprove_next": (i, [_vp]),
        "qpzk_sprove_phase": (u32, [_vp]),
        "qpzk_sprove_exchange": (i, [_vp, u32, ctypes.POINTER(_vp), ctypes.POINTER(u64), ctypes.POINTER(u64),
                                     ctypes.POINTER(u64), ctypes.POINTER(u32)]),
        "qpzk_sprove_end": (i, [_vp, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
        "qpzk_zs_partial_products": (i, [_vp, _vp, _u64p, _u64p, _u64p]),
        "qpzk_quotient": (i, [_vp, _vp, _vp, _u64p, _u64p, _u64p, _u64p, _u64p]),
        "qpzk_fri_begin": (i, [_vp, _vp, _vp, _vp, _u64p, _u64p, ctypes.POINTER(_vp)]),
        "qpzk_fri_num_rounds": (u32, [_vp]),
        "qpzk_fri_commit_round": (i, [_vp, _u64p]),
        "qpzk_fri_fold": (i, [_vp, _u64p]),
        "qpzk_fri_final_poly": (i, [_vp, _u64p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
        "qpzk_fri_query": (i, [_vp, u64, _u64p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
        "qpzk_fri_free": (None, [_vp]),
        "qpzk_prove_trace": (ctypes.c_size_t, [_vp, i, _u64p]),
        "qpzk_prove_stage_ms": (i, [_vp, ctypes.POINTER(ctypes.c_float)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    L._declared = sorted(sig)
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise QpzkError(rc, load_library().qpzk_last_error().decode(errors="replace"))


def _arr(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


def _ptr(a):
    return a.ctypes.data_as(_u64p)


class Context:
    """One CUDA stream + caches (twiddle tables, coset pre-multipliers) on one device."""

    def __init__(self, device=0, blocking_sync=False, yield_sync=False):
        """blocking_sync / yield_sync: QPZK_CTX_BLOCKING_SYNC / QPZK_CTX_YIELD_SYNC (include/qpzk.h)."""
        L = load_library()
        h = _vp()
        _check(L.qpzk_ctx_create(device, 1 if blocking_sync else (2 if yield_sync else 0), ctypes.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            load_library().qpzk_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _check(load_library().qpzk_ctx_sync(self._h))

    @property
    def stream(self):
        return load_library().qpzk_ctx_stream(self._h)

    def stage_ms(self):
        out = (ctypes.c_float * len(STAGES))()
        _check(load_library().qpzk_ctx_stage_ms(self._h, out))
        return dict(zip(STAGES, [float(x) for x in out]))

    def launch_count(self):
        return int(load_library().qpzk_ctx_launch_count(self._h))

    def measure_imad_peak(self, kind=1):
        out = ctypes.c_double()
        _check(load_library().qpzk_measure_imad_peak(self._h, kind, ctypes.byref(out)))
        return out.value

    # -- raw device memory (bench / multi-GPU plumbing) --
    def dev_alloc(self, nbytes):
        p = _vp()
        _check(load_library().qpzk_dev_alloc(self._h, nbytes, ctypes.byref(p)))
        return p.value

    def dev_free(self, p):
        load_library().qpzk_dev_free(self._h, _vp(p))

    def h2d(self, dev_ptr, host_array):
        a = np.ascontiguousarray(host_array)
        _check(load_library().qpzk_memcpy_h2d(self._h, _vp(dev_ptr), a.ctypes.data_as(_vp), a.nbytes))

    def d2h(self, host_array, dev_ptr):
        _check(load_library().qpzk_memcpy_d2h(self._h, host_array.ctypes.data_as(_vp), _vp(dev_ptr),
                                              host_array.nbytes))

    def fri_pow(self, sponge_state, input_pos, min_leading_zeros):
        """`fri_proof_of_work`: smallest valid witness for the given duplex-sponge state."""
        st = _arr(sponge_state)
        out = ctypes.c_uint64()
        _check(load_library().qpzk_fri_pow(self._h, _ptr(st), input_pos, min_leading_zeros, ctypes.byref(out)))
        return out.value

    # -- PoseidonHash --
    def poseidon_permute(self, states):
        s = _arr(states).copy().reshape(-1, 12)
        _check(load_library().qpzk_poseidon_permute(self._h, _ptr(s), s.shape[0]))
        return s

    def hash_no_pad(self, inputs):
        x = _arr(inputs)
        x2 = x.reshape(1, -1) if x.ndim == 1 else x
        out = np.zeros((x2.shape[0], 4), np.uint64)
        _check(load_library().qpzk_hash_no_pad(self._h, _ptr(x2), x2.shape[0], x2.shape[1], _ptr(out)))
        return out[0] if x.ndim == 1 else out

    def two_to_one(self, left, right):
        l, r = _arr(left).reshape(-1, 4), _arr(right).reshape(-1, 4)
        pairs = np.ascontiguousarray(np.concatenate([l, r], axis=1))
        out = np.zeros((pairs.shape[0], 4), np.uint64)
        _check(load_library().qpzk_two_to_one(self._h, _ptr(pairs), pairs.shape[0], _ptr(out)))
        return out[0] if np.ndim(left) == 1 else out


class PinnedBuffer:
    """Page-locked host array (optional fast path for callers that own their trace buffers)."""

    def __init__(self, shape, dtype=np.uint64):
        L = load_library()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = _vp()
        _check(L.qpzk_host_alloc(n, ctypes.byref(p)))
        self._p = p
        buf = (ctypes.c_uint8 * n).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            load_library().qpzk_host_free(self._p)
            self._p = None


class MerkleTree:
    """`MerkleTree::new(leaves, cap_height)` on the device; `cap`, `prove`, `digests` read back."""

    def __init__(self, ctx, leaves, cap_height):
        lv = _arr(leaves)
        if lv.ndim != 2:
            raise ValueError("leaves must be [nleaves][leaf_len]")
        self.ctx, self.nleaves, self.leaf_len, self.cap_height = ctx, lv.shape[0], lv.shape[1], cap_height
        h = _vp()
        _check(load_library().qpzk_merkle_new(ctx._h, _ptr(lv), lv.shape[0], lv.shape[1], cap_height,
                                              ctypes.byref(h)))
        self._h = h

    @property
    def cap(self):
        out = np.zeros((1 << self.cap_height, 4), np.uint64)
        _check(load_library().qpzk_tree_cap(self._h, _ptr(out)))
        return out

    def prove(self, leaf_index):
        nl = (self.nleaves.bit_length() - 1) - self.cap_height
        out = np.zeros((max(nl, 1), 4), np.uint64)
        _check(load_library().qpzk_tree_prove(self._h, leaf_index, _ptr(out)))
        return out[:nl]

    @property
    def digests(self):
        out = np.zeros((max(2 * (self.nleaves - (1 << self.cap_height)), 1), 4), np.uint64)
        _check(load_library().qpzk_tree_digests(self._h, _ptr(out)))
        return out[:2 * (self.nleaves - (1 << self.cap_height))]

    def free(self):
        if getattr(self, "_h", None):
            load_library().qpzk_tree_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PolynomialBatch:
    """Device-resident `PolynomialBatch` (coefficients, LDE leaves, Merkle tree)."""

    def __init__(self, ctx, handle, ncols, degree_bits, rate_bits, cap_height, salt_cols):
        self.ctx, self._h = ctx, handle
        self.ncols, self.degree_bits, self.rate_bits = ncols, degree_bits, rate_bits
        self.cap_height, self.salt_cols = cap_height, salt_cols

    @staticmethod
    def _make(ctx, fn, data, rate_bits, cap_height, salts, dev=False, shape=None):
        L = load_library()
        if dev:
            ncols, n = shape
            dptr = _vp(data)
            sptr = _vp(salts[0]) if salts is not None else None
            salt_cols = salts[1] if salts is not None else 0
            keep = None
        else:
            a = _arr(data)
            if a.ndim != 2:
                raise ValueError("expected column-major [ncols][n]")
            ncols, n = a.shape
            dptr = a.ctypes.data_as(_vp)
            keep = a
            sptr, salt_cols = None, 0
            if salts is not None:
                s = _arr(salts)
                if s.shape != (s.shape[0], n << rate_bits):
                    raise ValueError("salts must be [salt_cols][n << rate_bits]")
                sptr, salt_cols, keep = s.ctypes.data_as(_vp), s.shape[0], (a, s)
        if n & (n - 1) or n == 0:
            raise ValueError("polynomial length must be a power of two")
        k = n.bit_length() - 1
        h = _vp()
        _check(getattr(L, fn)(ctx._h, dptr, ncols, k, rate_bits, cap_height, sptr, salt_cols, ctypes.byref(h)))
        del keep
        return PolynomialBatch(ctx, h, ncols, k, rate_bits, cap_height, salt_cols)

    @classmethod
    def from_values(cls, ctx, values, rate_bits, cap_height, salts=None):
        """values: [ncols][n] evaluations on the subgroup. `blinding` = salts is not None."""
        return cls._make(ctx, "qpzk_batch_from_values", values, rate_bits, cap_height, salts)

    @classmethod
    def from_coeffs(cls, ctx, coeffs, rate_bits, cap_height, salts=None):
        return cls._make(ctx, "qpzk_batch_from_coeffs", coeffs, rate_bits, cap_height, salts)

    @classmethod
    def from_values_dev(cls, ctx, dev_ptr, ncols, n, rate_bits, cap_height, salts=None):
        return cls._make(ctx, "qpzk_batch_from_values_dev", dev_ptr, rate_bits, cap_height, salts, dev=True,
                         shape=(ncols, n))

    @classmethod
    def from_coeffs_dev(cls, ctx, dev_ptr, ncols, n, rate_bits, cap_height, salts=None):
        return cls._make(ctx, "qpzk_batch_from_coeffs_dev", dev_ptr, rate_bits, cap_height, salts, dev=True,
                         shape=(ncols, n))

    @classmethod
    def from_values_shard_dev(cls, ctx, dev_ptr, ncols, n, rate_bits, cap_height, subtree_begin, subtree_end,
                              salts=None, enqueue_only=False):
        """One rank's part of a multi-GPU commit: cap subtrees [subtree_begin, subtree_end) only.
        enqueue_only: return without waiting for the stream (qpzk_batch_from_values_shard_dev_async)."""
        L = load_library()
        k = n.bit_length() - 1
        sptr = _vp(salts[0]) if salts is not None else None
        salt_cols = salts[1] if salts is not None else 0
        h = _vp()
        fn = L.qpzk_batch_from_values_shard_dev_async if enqueue_only else L.qpzk_batch_from_values_shard_dev
        _check(fn(ctx._h, _vp(dev_ptr), ncols, k, rate_bits, cap_height, sptr, salt_cols, subtree_begin, subtree_end,
                  ctypes.byref(h)))
        b = PolynomialBatch(ctx, h, ncols, k, rate_bits, cap_height, salt_cols)
        b.subtrees = (subtree_begin, subtree_end)
        return b

    def to_bytes(self):
        """`Write::write_polynomial_batch` bytes (polynomials, merkle_tree, degree_log, rate_bits, blinding)."""
        L = load_library()
        need = ctypes.c_uint64(0)
        _check(L.qpzk_batch_serialized_size(self._h, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        _check(L.qpzk_batch_to_bytes(self._h, buf, need.value))
        return buf.raw

    @classmethod
    def from_bytes(cls, ctx, data, verify=False):
        """`Read::read_polynomial_batch`: the commitment goes to the device as stored, nothing is recomputed
        (verify=True: the LDE and the tree are recomputed on the device and compared). Returns (batch, bytes read)."""
        L = load_library()
        data = bytes(data)
        h, used = _vp(), ctypes.c_uint64(0)
        _check(L.qpzk_batch_from_bytes(ctx._h, data, len(data), IMPORT_VERIFY if verify else 0, ctypes.byref(h),
                                       ctypes.byref(used)))
        b = PolynomialBatch(ctx, h, int(L.qpzk_batch_ncols(h)), int(L.qpzk_batch_degree_bits(h)), 0, 0, 0)
        b.salt_cols = int(L.qpzk_batch_width(h)) - b.ncols
        # rate_bits and cap_height sit at fixed places in the layout
        n, w = 1 << b.degree_bits, b.ncols + b.salt_cols
        off = 8 + b.ncols * 8 * (n + 1)
        N = int.from_bytes(data[off:off + 8], "little")
        b.rate_bits = N.bit_length() - 1 - b.degree_bits
        off += 8 + N * 8 * (w + 1)
        nd = int.from_bytes(data[off:off + 8], "little")
        b.cap_height = int.from_bytes(data[off + 8 + 32 * nd:off + 16 + 32 * nd], "little")
        return b, used.value

    def set_cap(self, cap):
        c = _arr(cap)
        if c.shape != (1 << self.cap_height, 4):
            raise ValueError("cap must be [2^cap_height][4]")
        _check(load_library().qpzk_batch_set_cap(self._h, _ptr(c)))

    @property
    def width(self):
        return self.ncols + self.salt_cols

    @property
    def lde_size(self):
        return 1 << (self.degree_bits + self.rate_bits)

    @property
    def cap(self):
        out = np.zeros((1 << self.cap_height, 4), np.uint64)
        _check(load_library().qpzk_batch_cap(self._h, _ptr(out)))
        return out

    @property
    def cap_dev(self):
        return load_library().qpzk_batch_cap_dev(self._h)

    @property
    def polynomials(self):
        out = np.zeros((self.ncols, 1 << self.degree_bits), np.uint64)
        _check(load_library().qpzk_batch_coeffs(self._h, _ptr(out)))
        return out

    def get_lde_values(self, index, step=1):
        idx = np.ascontiguousarray(np.atleast_1d(index), dtype=np.uint32)
        out = np.zeros((idx.size, self.ncols), np.uint64)
        _check(load_library().qpzk_batch_get_lde_rows(self._h, idx.ctypes.data_as(_u32p), idx.size, step,
                                                      _ptr(out)))
        return out[0] if np.ndim(index) == 0 else out

    def eval_ext(self, point):
        """Every committed polynomial at an extension point (a, b): [ncols][2] (`OpeningSet::new`)."""
        pt = _arr(point)
        out = np.zeros((self.ncols, 2), np.uint64)
        _check(load_library().qpzk_batch_eval_ext(self._h, _ptr(pt), _ptr(out)))
        return out

    def open(self, leaf_index):
        """(merkle_tree.leaves[leaf_index], merkle_tree.prove(leaf_index).siblings)"""
        nl = self.degree_bits + self.rate_bits - self.cap_height
        leaf = np.zeros(self.width, np.uint64)
        sib = np.zeros((max(nl, 1), 4), np.uint64)
        _check(load_library().qpzk_batch_open(self._h, leaf_index, _ptr(leaf), _ptr(sib)))
        return leaf, sib[:nl]

    def export(self, leaves=True, digests=True):
        N = self.lde_size
        lv = np.zeros((N, self.width), np.uint64) if leaves else None
        nd = 2 * (N - (1 << self.cap_height))
        dg = np.zeros((max(nd, 1), 4), np.uint64) if digests else None
        _check(load_library().qpzk_batch_export(self._h, _ptr(lv) if leaves else None,
                                                _ptr(dg) if digests else None))
        return lv, (dg[:nd] if digests else None)

    def free(self):
        if getattr(self, "_h", None):
            load_library().qpzk_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Fri:
    """The FRI prover driven step by step (the caller owns the transcript): see include/qpzk.h."""

    def __init__(self, circuit, wires_batch, zs_batch, quotient_batch, zeta, alpha, cap_height=4):
        z, a = _arr(zeta), _arr(alpha)
        h = _vp()
        _check(load_library().qpzk_fri_begin(circuit._h, wires_batch._h, zs_batch._h, quotient_batch._h, _ptr(z), _ptr(a),
                                             ctypes.byref(h)))
        self._h, self.cap_height = h, cap_height
        self.num_rounds = int(load_library().qpzk_fri_num_rounds(h))

    def commit_round(self):
        cap = np.zeros((1 << self.cap_height, 4), np.uint64)
        _check(load_library().qpzk_fri_commit_round(self._h, _ptr(cap)))
        return cap

    def fold(self, beta):
        b = _arr(beta)
        _check(load_library().qpzk_fri_fold(self._h, _ptr(b)))

    def _sized(self, fn, *args):
        ln = ctypes.c_size_t(0)
        _check(fn(self._h, *args, None, 0, ctypes.byref(ln)))
        out = np.zeros(max(ln.value, 1), np.uint64)
        _check(fn(self._h, *args, _ptr(out), out.size, ctypes.byref(ln)))
        return out[:ln.value]

    def final_poly(self):
        return self._sized(load_library().qpzk_fri_final_poly).reshape(-1, 2)

    def query(self, x_index):
        return self._sized(load_library().qpzk_fri_query, int(x_index))

    def free(self):
        if getattr(self, "_h", None):
            load_library().qpzk_fri_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ShardedProof:
    """One rank's part of a proof spread over several GPUs (include/qpzk.h, qpzk_sprove_*): run `next()` until
    `phase == 6`, performing the exchanges `exchanges()` lists between phases (qpzk.dist has the NCCL and the
    in-process spellings), then `end()`."""

    ALLGATHER, SUM = 1, 2

    def __init__(self, circuit, handle, keep):
        self.circuit, self._h, self._keep = circuit, handle, keep

    @property
    def phase(self):
        return int(load_library().qpzk_sprove_phase(self._h))

    def exchanges(self):
        """[(kind, device pointer, words, own_begin, own_end)] due after the phase just run."""
        L = load_library()
        out = []
        idx = 0
        while True:
            p, w, b, e, k = _vp(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint32()
            _check(L.qpzk_sprove_exchange(self._h, idx, ctypes.byref(p), ctypes.byref(w), ctypes.byref(b), ctypes.byref(e),
                                          ctypes.byref(k)))
            if k.value == 0:
                return out
            out.append((k.value, p.value, w.value, b.value, e.value))
            idx += 1

    def next(self):
        _check(load_library().qpzk_sprove_next(self._h))

    def end(self):
        buf = self.circuit._buf
        ln = ctypes.c_size_t(0)
        h, self._h = self._h, None
        _check(load_library().qpzk_sprove_end(h, buf, len(buf), ctypes.byref(ln)))
        self._keep = None
        return buf.raw[:ln.value]


PROVE_STAGES = ("commit_wires", "zs_partial_products_commit", "quotient_commit", "openings", "fri_combine",
                "fri_commit_phase", "proof_of_work", "queries")


class Circuit:
    """Prover-side circuit data (`ProverCircuitData` after `build()`): the constants|sigmas batch is
    committed once here and stays on the device; `prove(wires, public_inputs)` then mirrors
    `ProverCircuitData::prove` after witness generation and returns `ProofWithPublicInputs` bytes.
    One proof at a time per Circuit; `prove_begin` / `prove_end` split the call so that one host thread can
    keep several contexts busy."""

    INFO = ("degree_bits", "rate_bits", "cap_height", "num_wires", "num_routed", "num_challenges", "salt_cols",
            "num_public_inputs")

    def __init__(self, ctx, common_bytes, circuit_digest, constants_sigmas=None, commitment=None, verify=False):
        """constants_sigmas: the value columns (committed here, `build()`), or commitment: the serialized
        `constants_sigmas_commitment` of a prover restored from files (uploaded as stored, `new_from_files`)."""
        L = load_library()
        dg = _arr(circuit_digest)
        h = _vp()
        cb = bytes(common_bytes)
        if commitment is not None:
            cm = bytes(commitment)
            _check(L.qpzk_circuit_create_from_commitment(ctx._h, cb, len(cb), _ptr(dg), cm, len(cm),
                                                         IMPORT_VERIFY if verify else 0, ctypes.byref(h)))
        else:
            cs = _arr(constants_sigmas)
            _check(L.qpzk_circuit_create(ctx._h, cb, len(cb), _ptr(dg), cs.ctypes.data_as(_vp), cs.size, ctypes.byref(h)))
        self.ctx, self._h, self.common = ctx, h, cb
        info = (ctypes.c_uint32 * 8)()
        _check(L.qpzk_circuit_info(h, info))
        self.info = dict(zip(self.INFO, [int(x) for x in info]))
        self.n = 1 << self.info["degree_bits"]
        self.wires_words = self.info["num_wires"] * self.n
        self.salt_words = SALT_SIZE << (self.info["degree_bits"] + self.info["rate_bits"])
        self._keep = None
        self._buf = ctypes.create_string_buffer(1 << 20)

    @property
    def constants_sigmas_cap(self):
        ncap = 1 << self.info["cap_height"]
        out = np.zeros((ncap, 4), np.uint64)
        _check(load_library().qpzk_circuit_cap(self._h, _ptr(out), out.size))
        return out

    def commitment_bytes(self):
        """The serialized `constants_sigmas_commitment` (the PolynomialBatch field of `ProverOnlyCircuitData::to_bytes`)."""
        L = load_library()
        need = ctypes.c_uint64(0)
        _check(L.qpzk_circuit_commitment_size(self._h, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        _check(L.qpzk_circuit_commitment_to_bytes(self._h, buf, need.value))
        return buf.raw

    def verifier_only_bytes(self):
        L = load_library()
        need = L.qpzk_circuit_verifier_only(self._h, None, 0)
        buf = ctypes.create_string_buffer(max(need, 1))
        k = L.qpzk_circuit_verifier_only(self._h, buf, need)
        return buf.raw[:k]

    def _host_args(self, wires, public_inputs, salts):
        w = _arr(wires)
        pi = _arr(public_inputs)
        sp, keep, salt_words = [None, None, None], [w, pi], 0
        if salts is not None:
            for j in range(3):
                a = _arr(salts[j])
                keep.append(a)
                sp[j] = a.ctypes.data_as(_vp)
                salt_words = a.size
        return w.ctypes.data_as(_vp), w.size, pi, sp, salt_words, keep

    def prove(self, wires, public_inputs, salts=None, trace=False, seed=None):
        """salts: the three [4][N] blinding arrays of a hiding circuit, or None with `seed` = 4 u64 words: the
        salts are then drawn on the device (QPZK_PROVE_SEEDED_SALTS; qpzk.synth.seeded_salts restates them)."""
        self.prove_begin(wires, public_inputs, salts, trace=trace, seed=seed)
        return self.prove_end()

    def prove_dev(self, wires_dev, public_inputs, salts_dev=None, seed=None):
        """Same as prove() with the witness matrix (and salts) already resident on the device."""
        self.prove_begin_dev(wires_dev, public_inputs, salts_dev, seed=seed)
        return self.prove_end()

    def prove_begin(self, wires, public_inputs, salts=None, trace=False, seed=None):
        """Enqueue one proof (qpzk_prove_begin); the host arrays must stay alive until prove_end()."""
        wp, wn, pi, sp, sn, keep = self._host_args(wires, public_inputs, salts)
        flags = 1 if trace else 0
        if seed is not None:
            sd = _arr(seed)
            if sd.size != 4:
                raise ValueError("seed must be 4 u64 words")
            keep.append(sd)
            sp, sn, flags = [sd.ctypes.data_as(_vp), None, None], 4, flags | 4
        _check(load_library().qpzk_prove_begin(self._h, wp, wn, _ptr(pi), pi.size, sp[0], sp[1], sp[2], sn, flags))
        self._keep = keep

    def prove_begin_dev(self, wires_dev, public_inputs, salts_dev=None, seed=None):
        pi = _arr(public_inputs)
        sp = [None, None, None] if salts_dev is None else [_vp(x) for x in salts_dev]
        sn, flags, keep = (self.salt_words if salts_dev is not None else 0), 2, [pi]
        if seed is not None:
            sd = _arr(seed)
            keep.append(sd)
            sp, sn, flags = [sd.ctypes.data_as(_vp), None, None], 4, 6
        _check(load_library().qpzk_prove_begin(self._h, _vp(wires_dev), self.wires_words, _ptr(pi), pi.size, sp[0], sp[1],
                                               sp[2], sn, flags))
        self._keep = keep

    def prove_end(self):
        ln = ctypes.c_size_t(0)
        _check(load_library().qpzk_prove_end(self._h, self._buf, len(self._buf), ctypes.byref(ln)))
        self._keep = None
        return self._buf.raw[:ln.value]

    def sprove_begin(self, wires, public_inputs, salts, subtree_begin, subtree_end, on_device=False):
        """First phase of one rank's part of a multi-GPU proof (qpzk_sprove_begin); returns a ShardedProof."""
        if on_device:
            pi = _arr(public_inputs)
            wp, wn, keep = _vp(wires), self.wires_words, [pi]
            sp = [None, None, None] if salts is None else [_vp(x) for x in salts]
            sn = self.salt_words if salts is not None else 0
        else:
            wp, wn, pi, sp, sn, keep = self._host_args(wires, public_inputs, salts)
        h = _vp()
        _check(load_library().qpzk_sprove_begin(self._h, wp, wn, _ptr(pi), pi.size, sp[0], sp[1], sp[2], sn,
                                                2 if on_device else 0, subtree_begin, subtree_end, ctypes.byref(h)))
        return ShardedProof(self, h, keep)

    def zs_partial_products(self, wires, betas, gammas, nch=2, npp=9):
        """H8 as a stand-alone stage: [nch*(1+npp)][n]."""
        w, b, g = _arr(wires), _arr(betas), _arr(gammas)
        out = np.zeros((nch * (1 + npp), self.n), np.uint64)
        _check(load_library().qpzk_zs_partial_products(self._h, w.ctypes.data_as(_vp), _ptr(b), _ptr(g), _ptr(out)))
        return out

    def quotient(self, wires_batch, zs_batch, pi_hash, betas, gammas, alphas, nch=2, qdf=8):
        """H9 as a stand-alone stage: quotient chunk coefficients [nch*qdf][n]."""
        ph, b, g, a = _arr(pi_hash), _arr(betas), _arr(gammas), _arr(alphas)
        out = np.zeros((nch * qdf, self.n), np.uint64)
        _check(load_library().qpzk_quotient(self._h, wires_batch._h, zs_batch._h, _ptr(ph), _ptr(b), _ptr(g), _ptr(a),
                                            _ptr(out)))
        return out

    def trace(self, which):
        L = load_library()
        k = L.qpzk_prove_trace(self._h, which, None)
        out = np.zeros(max(k, 1), np.uint64)
        L.qpzk_prove_trace(self._h, which, _ptr(out))
        return out[:k]

    def stage_ms(self):
        out = (ctypes.c_float * 16)()
        _check(load_library().qpzk_prove_stage_ms(self._h, out))
        return dict(zip(PROVE_STAGES, [float(x) for x in out][:len(PROVE_STAGES)]))

    def free(self):
        if getattr(self, "_h", None):
            load_library().qpzk_circuit_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
