"""Oracle self-consistency for the commit path (CPU): radix-2 NTT vs naive evaluation, the plonky2
digest layout vs MerkleTree::prove, and from_values semantics (H2-H7 in SURVEY.md §8(a))."""
import numpy as np

from helpers import GEN, P, bitrev, rand_felts
from oracle import oracle as orc


def test_ntt_matches_naive_evaluation():
    rng = np.random.default_rng(3)
    for k in (0, 1, 3, 6):
        n = 1 << k
        c = rand_felts(rng, n)
        assert np.array_equal(orc.fft(c), orc.naive_coset_eval(c, n, 1))
        assert np.array_equal(orc.coset_fft(c, GEN), orc.naive_coset_eval(c, n, GEN))
        assert np.array_equal(orc.ifft(orc.fft(c)), c)


def test_root_of_unity_convention():
    # primitive_root_of_unity(k) = 7277203076849721926^(2^(32-k))  
    assert orc.root_of_unity(32) == 7277203076849721926
    assert orc.fpow(GEN, (P - 1) >> 32) == 7277203076849721926
    w = orc.root_of_unity(5)
    assert orc.fpow(w, 32) == 1 and orc.fpow(w, 16) == P - 1


def test_merkle_layout_prove_verify_roundtrip():
    rng = np.random.default_rng(4)
    for log_n, cap_h, width in ((4, 0, 7), (5, 2, 3), (6, 4, 20), (3, 3, 9)):
        n = 1 << log_n
        leaves = rand_felts(rng, (n, width))
        dig, cap = orc.merkle_new(leaves, cap_h)
        assert dig.shape[0] == 2 * (n - (1 << cap_h))
        dig8, cap8 = orc.merkle_new(leaves, cap_h, threads=8)
        assert np.array_equal(dig, dig8) and np.array_equal(cap, cap8)
        for i in range(n):
            sib = orc.merkle_prove(dig, n, cap_h, i)
            assert sib.shape[0] == log_n - cap_h
            assert orc.merkle_verify(leaves[i], i, cap, sib)
            if sib.shape[0]:
                bad = sib.copy()
                bad[0, 0] ^= np.uint64(1)
                assert not orc.merkle_verify(leaves[i], i, cap, bad)


def test_from_values_semantics():
    rng = np.random.default_rng(5)
    k, r, cap_h, ncols = 4, 3, 2, 5
    n, N = 1 << k, 1 << (k + r)
    vals = rand_felts(rng, (ncols, n))
    salts = rand_felts(rng, (4, N))
    out = orc.batch_commit(vals, r, cap_h, salts=salts, threads=2)
    wN = orc.root_of_unity(k + r)
    for c in range(ncols):
        assert np.array_equal(orc.fft(out["coeffs"][c]), vals[c])  # coefficients interpolate the values
        ev = orc.naive_coset_eval(out["coeffs"][c], N, GEN)         # p(g * w_N^i), natural order
        for i in range(N):
            assert out["leaves"][bitrev(i, k + r), c] == ev[i]
    for s in range(4):
        for i in range(0, N, 7):
            assert out["leaves"][bitrev(i, k + r), ncols + s] == salts[s, i]
    # from_coeffs on the coefficients gives the same tree
    out2 = orc.batch_commit(out["coeffs"], r, cap_h, is_coeffs=True, salts=salts)
    assert np.array_equal(out2["cap"], out["cap"]) and np.array_equal(out2["digests"], out["digests"])
    assert wN == orc.fpow(orc.root_of_unity(32), 1 << (32 - k - r))
