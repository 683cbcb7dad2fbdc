"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes loader for ``oracle/liboracle.so`` (the CPU restatement of the qp-plonky2 1.1.1 arithmetic
the reference calls into; see the headers of oracle/*.hpp for the reference file:line each piece
follows).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

P = 0xFFFFFFFF00000001
GEN = 14293326489335486720

u64p = ctypes.POINTER(ctypes.c_uint64)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        L.orc_last_error.restype = ctypes.c_char_p
        for name in ("orc_mul", "orc_inv", "orc_pow", "orc_root_of_unity", "orc_challenger_get"):
            getattr(L, name).restype = ctypes.c_uint64
        L.orc_mul.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        L.orc_pow.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        L.orc_inv.argtypes = [ctypes.c_uint64]
        L.orc_root_of_unity.argtypes = [ctypes.c_uint]
        L.orc_challenger_new.restype = ctypes.c_void_p
        L.orc_challenger_free.argtypes = [ctypes.c_void_p]
        L.orc_challenger_observe.argtypes = [ctypes.c_void_p, u64p, ctypes.c_uint64]
        L.orc_challenger_get.argtypes = [ctypes.c_void_p]
        _LIB = L
    return _LIB


def _a(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


def _p(a):
    return a.ctypes.data_as(u64p)


def selfcheck():
    return lib().orc_selfcheck()


def poseidon(state, naive=False):
    s = _a(state).copy()
    assert s.shape == (12,)
    (lib().orc_poseidon_naive if naive else lib().orc_poseidon)(_p(s))
    return s


def poseidon_tables():
    rc, first, frc = np.zeros(360, np.uint64), np.zeros(12, np.uint64), np.zeros(22, np.uint64)
    init, what, v = np.zeros((11, 11), np.uint64), np.zeros((22, 11), np.uint64), np.zeros((22, 11), np.uint64)
    lib().orc_poseidon_tables(_p(rc), _p(first), _p(frc), _p(init), _p(what), _p(v))
    return dict(rc=rc, fast_first=first, fast_rc=frc, fast_init=init, fast_w_hat=what, fast_v=v)


def hash_no_pad(x):
    x = _a(x)
    out = np.zeros(4, np.uint64)
    lib().orc_hash_no_pad(_p(x), ctypes.c_uint64(x.size), _p(out))
    return out


def hash_or_noop(x):
    x = _a(x)
    out = np.zeros(4, np.uint64)
    lib().orc_hash_or_noop(_p(x), ctypes.c_uint64(x.size), _p(out))
    return out


def two_to_one(l, r):
    l, r = _a(l), _a(r)
    out = np.zeros(4, np.uint64)
    lib().orc_two_to_one(_p(l), _p(r), _p(out))
    return out


def mul(a, b):
    return lib().orc_mul(a, b)


def inv(a):
    return lib().orc_inv(a)


def fpow(a, e):
    return lib().orc_pow(a, e)


def root_of_unity(bits):
    return lib().orc_root_of_unity(bits)


def fft(a):
    a = _a(a).copy()
    lib().orc_fft(_p(a), ctypes.c_uint64(a.size))
    return a


def ifft(a):
    a = _a(a).copy()
    lib().orc_ifft(_p(a), ctypes.c_uint64(a.size))
    return a


def coset_fft(a, shift=GEN):
    a = _a(a).copy()
    lib().orc_coset_fft(_p(a), ctypes.c_uint64(a.size), ctypes.c_uint64(shift))
    return a


def naive_coset_eval(coeffs, npoints, shift):
    c = _a(coeffs)
    out = np.zeros(npoints, np.uint64)
    lib().orc_naive_coset_eval(_p(c), ctypes.c_uint64(c.size), ctypes.c_uint64(npoints), ctypes.c_uint64(shift), _p(out))
    return out


def have_avx512():
    """True when the batched AVX-512 Poseidon / NTT paths are in use on this CPU."""
    return bool(lib().orc_have_avx512())


def merkle_new(leaves, cap_height, threads=1):
    """leaves: [nleaves][leaf_len] row-major. Returns (digests[.,4] plonky2 layout, cap[2^h,4])."""
    lv = _a(leaves)
    n, w = lv.shape
    dig = np.zeros((2 * (n - (1 << cap_height)), 4), np.uint64)
    cap = np.zeros((1 << cap_height, 4), np.uint64)
    lib().orc_merkle_new(_p(lv), ctypes.c_uint64(n), ctypes.c_uint64(w), ctypes.c_uint(cap_height),
                         ctypes.c_uint(threads), _p(dig) if dig.size else None, _p(cap))
    return dig, cap


def merkle_prove(digests, nleaves, cap_height, leaf_index):
    d = _a(digests)
    nl = (int(nleaves).bit_length() - 1) - cap_height
    sib = np.zeros((max(nl, 1), 4), np.uint64)
    k = lib().orc_merkle_prove(_p(d), ctypes.c_uint64(nleaves), ctypes.c_uint(cap_height),
                               ctypes.c_uint64(leaf_index), _p(sib))
    return sib[:k]


def merkle_verify(leaf, leaf_index, cap, siblings):
    leaf, cap, sib = _a(leaf), _a(cap), _a(siblings)
    return bool(lib().orc_merkle_verify(_p(leaf), ctypes.c_uint64(leaf.size), ctypes.c_uint64(leaf_index),
                                        _p(cap), _p(sib) if sib.size else None,
                                        ctypes.c_uint64(sib.shape[0] if sib.ndim == 2 else 0)))


def batch_commit(cols, rate_bits, cap_height, is_coeffs=False, salts=None, threads=1,
                 want_leaves=True, want_digests=True):
    """PolynomialBatch::from_values / from_coeffs over column-major ``cols`` [ncols][n].

    Returns dict(coeffs [ncols][n], leaves [N][ncols+salt] row-major bit-reversed, digests, cap)."""
    cols = _a(cols)
    ncols, n = cols.shape
    k = n.bit_length() - 1
    N = n << rate_bits
    salt_cols = 0
    sp = None
    if salts is not None:
        salts = _a(salts)
        salt_cols = salts.shape[0]
        assert salts.shape[1] == N
        sp = _p(salts)
    coeffs = np.zeros((ncols, n), np.uint64)
    leaves = np.zeros((N, ncols + salt_cols), np.uint64) if want_leaves else None
    dig = np.zeros((2 * (N - (1 << cap_height)), 4), np.uint64) if want_digests else None
    cap = np.zeros((1 << cap_height, 4), np.uint64)
    rc = lib().orc_batch_commit(_p(cols), ctypes.c_int(1 if is_coeffs else 0), ctypes.c_uint64(ncols),
                                ctypes.c_uint(k), ctypes.c_uint(rate_bits), ctypes.c_uint(cap_height), sp,
                                ctypes.c_uint(salt_cols), ctypes.c_uint(threads), _p(coeffs),
                                _p(leaves) if want_leaves else None,
                                _p(dig) if (want_digests and dig.size) else None, _p(cap))
    if rc != 0:
        raise RuntimeError(lib().orc_last_error().decode())
    return dict(coeffs=coeffs, leaves=leaves, digests=dig, cap=cap)


def batch_to_bytes(commit, rate_bits, blinding):
    """`Write::write_polynomial_batch` of qp-plonky2 1.1.1 (util/serialization, un-vendored; restated - no
    fixture of a serialized ProverOnlyCircuitData ships with the reference, so this layout is UNPINNED): the
    `constants_sigmas_commitment` field read back by `WormholeProver::new_from_files`
    (/root/reference/wormhole/prover/src/lib.rs:105-187). ``commit`` = the dict batch_commit returns.
    usize = u64 little-endian, bool = 1 byte; vectors of field elements carry a length, the cap does not."""
    coeffs, leaves, dig, cap = commit["coeffs"], commit["leaves"], commit["digests"], commit["cap"]
    ncols, n = coeffs.shape
    u = lambda v: int(v).to_bytes(8, "little")
    out = [u(ncols)]
    for c in range(ncols):
        out += [u(n), coeffs[c].astype("<u8").tobytes()]
    N, w = leaves.shape
    rec = np.empty((N, w + 1), "<u8")
    rec[:, 0] = w
    rec[:, 1:] = leaves
    out += [u(N), rec.tobytes(), u(dig.shape[0] if dig is not None and dig.size else 0)]
    if dig is not None and dig.size:
        out.append(dig.astype("<u8").tobytes())
    out += [u(cap.shape[0].bit_length() - 1), cap.astype("<u8").tobytes(), u(n.bit_length() - 1), u(rate_bits),
            b"\x01" if blinding else b"\x00"]
    return b"".join(out)


class Challenger:
    def __init__(self):
        self._h = lib().orc_challenger_new()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_challenger_free(self._h)
            self._h = None

    def observe(self, xs):
        xs = _a(np.atleast_1d(xs)).ravel()
        lib().orc_challenger_observe(self._h, _p(xs), ctypes.c_uint64(xs.size))

    def get(self):
        return lib().orc_challenger_get(self._h)

    def get_n(self, n):
        return [self.get() for _ in range(n)]


def verify(common, vonly, proof):
    """Restated plonky2 verifier. Returns (code, challenges dict); code 0 = accepted."""
    ch = np.zeros(11 + 4096, np.uint64)
    rc = lib().orc_verify(common, ctypes.c_uint64(len(common)), vonly, ctypes.c_uint64(len(vonly)),
                          proof, ctypes.c_uint64(len(proof)), _p(ch))
    if rc < 0:
        raise RuntimeError(lib().orc_last_error().decode())
    d = dict(betas=ch[0:2], gammas=ch[2:4], alphas=ch[4:6], zeta=ch[6:8], fri_alpha=ch[8:10],
             pow_response=int(ch[10]), query_indices=ch[11:])
    return rc, d


def proof_roundtrip(common, proof):
    return lib().orc_proof_roundtrip(common, ctypes.c_uint64(len(common)), proof, ctypes.c_uint64(len(proof)))


def check_proof_paths(common, proof, x_indices):
    xi = _a(x_indices)
    return lib().orc_check_proof_paths(common, ctypes.c_uint64(len(common)), proof,
                                       ctypes.c_uint64(len(proof)), _p(xi))


class Circuit:
    """Prover-side circuit data for the restated `prove()` (oracle/prover.hpp)."""

    def __init__(self, common, digest, constants_sigmas, threads=1):
        L = lib()
        L.orc_circuit_new.restype = ctypes.c_void_p
        self.common = bytes(common)
        cs = _a(constants_sigmas)
        dg = _a(digest)
        self._h = L.orc_circuit_new(self.common, ctypes.c_uint64(len(self.common)), _p(dg), _p(cs),
                                    ctypes.c_uint(threads))
        if not self._h:
            raise RuntimeError(L.orc_last_error().decode())
        self.ncs, self.n = cs.shape
        self.threads = threads

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_circuit_free(ctypes.c_void_p(self._h))
            self._h = None

    def verifier_only_bytes(self):
        L = lib()
        L.orc_circuit_verifier_only.restype = ctypes.c_uint64
        buf = ctypes.create_string_buffer(4096)
        k = L.orc_circuit_verifier_only(ctypes.c_void_p(self._h), buf, ctypes.c_uint64(4096))
        return buf.raw[:k]

    def cs_coeffs(self):
        out = np.zeros((self.ncs, self.n), np.uint64)
        lib().orc_circuit_cs_coeffs(ctypes.c_void_p(self._h), _p(out))
        return out

    def prove(self, wires, public_inputs, salts=None):
        L = lib()
        L.orc_prove.restype = ctypes.c_int64
        w, pi = _a(wires), _a(public_inputs)
        sp = [None, None, None]
        keep = []
        if salts is not None:
            for i in range(3):
                a = _a(salts[i])
                keep.append(a)
                sp[i] = _p(a)
        cap = 1 << 20
        buf = ctypes.create_string_buffer(cap)
        k = L.orc_prove(ctypes.c_void_p(self._h), _p(w), _p(pi), ctypes.c_uint64(pi.size), sp[0], sp[1], sp[2],
                        ctypes.c_uint(self.threads), buf, ctypes.c_uint64(cap))
        if k < 0:
            raise RuntimeError(L.orc_last_error().decode())
        return buf.raw[:k]

    def trace(self, nch=2, npp=9, qdf=8, rounds=3):
        L = lib()
        h = ctypes.c_void_p(self._h)
        ch = np.zeros(3 * nch + 4 + 2 * rounds, np.uint64)
        L.orc_trace_challenges(h, _p(ch))
        zs = np.zeros((nch * (1 + npp), self.n), np.uint64)
        L.orc_trace_zs_pp(h, _p(zs))
        q = np.zeros((nch * qdf, self.n), np.uint64)
        L.orc_trace_quotient_chunks(h, _p(q))
        fp = np.zeros((self.n, 2), np.uint64)
        L.orc_trace_final_poly(h, _p(fp))
        return dict(betas=ch[0:nch], gammas=ch[nch:2 * nch], alphas=ch[2 * nch:3 * nch],
                    zeta=ch[3 * nch:3 * nch + 2], fri_alpha=ch[3 * nch + 2:3 * nch + 4],
                    fri_betas=ch[3 * nch + 4:].reshape(-1, 2), zs_pp=zs, quotient_chunks=q, final_poly=fp)
