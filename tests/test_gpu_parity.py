"""GPU parity: every result that crosses the C ABI is compared bit-for-bit with the CPU oracle on
the same seeded inputs (integer work: exact equality is the bar), plus size-independent properties
at the full microbench size. All calls go through libqpzk.so's C entry points."""
import numpy as np
import pytest

from helpers import GEN, P, bitrev, rand_felts, splitmix64
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import qpzk
    c = qpzk.Context(0)
    yield c
    c.close()


def test_permutation_kats(ctx):
    states = np.stack([np.zeros(12, np.uint64), np.arange(12, dtype=np.uint64), np.full(12, P - 1, np.uint64)])
    out = ctx.poseidon_permute(states)
    assert [int(x) for x in out[0][:4]] == [0x3c18a9786cb0b359, 0xc4055e3364a246c3, 0x7953db0ab48808f4,
                                            0xc71603f33a1144ca]
    assert int(out[1][0]) == 0xd64e1e3efc5b8e9e and int(out[1][11]) == 0x5c0a27fcb0e1459b
    assert int(out[2][0]) == 0xbe0085cfc57a8357


def test_permutation_random_and_noncanonical(ctx):
    rng = np.random.default_rng(11)
    s = rand_felts(rng, (1000, 12))
    # non-canonical representatives (>= p) and extreme words must reduce to the same answer
    s[0, :] = np.uint64(0xFFFFFFFFFFFFFFFF)
    s[1, :] = np.uint64(P)
    s[2, ::2] = np.uint64(P - 1)
    s[3, :] = np.uint64(0xFFFFFFFF)
    s[4, :] = np.uint64(0xFFFFFFFF00000000)
    out = ctx.poseidon_permute(s)
    for i in range(s.shape[0]):
        assert np.array_equal(out[i], orc.poseidon(s[i])), i  # the oracle canonicalises its input


@pytest.mark.parametrize("length", [1, 4, 7, 8, 9, 10, 12, 16, 17, 135, 188])
def test_hash_no_pad(ctx, length):
    rng = np.random.default_rng(length)
    x = rand_felts(rng, (33, length))
    out = ctx.hash_no_pad(x)
    for i in range(x.shape[0]):
        assert np.array_equal(out[i], orc.hash_no_pad(x[i]))


def test_reference_kats_on_gpu(ctx):
    """P1/P2 of SURVEY §8(c) through the CUDA path (unspendable_account_tests.rs:12-27, prover_tests.rs:29-43)."""
    from test_oracle_poseidon import (ADDRESSES, DEFAULT_SECRET, NULLIFIER_BYTES, SECRETS, bytes4_to_felts,
                                      felts_to_digest, string_to_felts, u64_to_felts)
    for secret, address in zip(SECRETS, ADDRESSES):
        pre = np.array(string_to_felts("wormhole") + bytes4_to_felts(bytes.fromhex(secret)), np.uint64)
        assert felts_to_digest(ctx.hash_no_pad(ctx.hash_no_pad(pre))).hex() == address
    pre = np.array(string_to_felts("~nullif~") + bytes4_to_felts(bytes.fromhex(DEFAULT_SECRET)) + u64_to_felts(4),
                   np.uint64)
    assert felts_to_digest(ctx.hash_no_pad(ctx.hash_no_pad(pre))) == NULLIFIER_BYTES


def test_two_to_one(ctx):
    rng = np.random.default_rng(12)
    l, r = rand_felts(rng, (50, 4)), rand_felts(rng, (50, 4))
    out = ctx.two_to_one(l, r)
    for i in range(50):
        assert np.array_equal(out[i], orc.two_to_one(l[i], r[i]))


@pytest.mark.parametrize("log_n,cap_h,width", [(0, 0, 5), (3, 3, 9), (4, 0, 7), (5, 2, 3), (6, 4, 4), (9, 4, 32),
                                               (13, 4, 32), (10, 0, 1)])
def test_merkle_tree(ctx, log_n, cap_h, width):
    """MerkleTree::new: cap, plonky2 digest layout and prove(), incl. hash_or_noop rows (width <= 4)
    and the all-leaves-in-cap case."""
    import qpzk
    rng = np.random.default_rng(100 + log_n)
    n = 1 << log_n
    leaves = rand_felts(rng, (n, width))
    t = qpzk.MerkleTree(ctx, leaves, cap_h)
    dig, cap = orc.merkle_new(leaves, cap_h, threads=4)
    assert np.array_equal(t.cap, cap)
    assert np.array_equal(t.digests, dig)
    for i in sorted(set([0, n - 1, n // 2, int(rng.integers(0, n))])):
        sib = t.prove(i)
        assert np.array_equal(sib, orc.merkle_prove(dig, n, cap_h, i))
        assert orc.merkle_verify(leaves[i], i, cap, sib)
    t.free()


def test_merkle_tree_rejects_bad_arguments(ctx):
    import qpzk
    with pytest.raises(qpzk.QpzkError):
        qpzk.MerkleTree(ctx, np.zeros((6, 3), np.uint64), 1)      # not a power of two
    with pytest.raises(qpzk.QpzkError):
        qpzk.MerkleTree(ctx, np.zeros((4, 3), np.uint64), 3)      # cap higher than the tree


COMMIT_CASES = [
    # (degree_bits, ncols, rate_bits, cap_height, salted)
    (0, 3, 3, 2, False), (1, 2, 3, 4, False), (3, 5, 3, 4, True), (5, 20, 3, 4, True), (8, 9, 3, 4, False),
    (9, 4, 1, 0, False), (12, 6, 3, 4, True),           # single-CTA transform path
    (13, 5, 3, 4, True), (14, 3, 3, 4, False), (15, 2, 2, 3, False), (16, 2, 3, 4, False),  # two-pass path
    (10, 1, 3, 4, False), (7, 2, 3, 10, False),         # width <= 4 (hash_or_noop), cap == all leaves
]


@pytest.mark.parametrize("k,ncols,r,cap_h,salted", COMMIT_CASES)
def test_from_values_matches_oracle(ctx, k, ncols, r, cap_h, salted):
    import qpzk
    rng = np.random.default_rng(1000 + 17 * k + ncols)
    n, N = 1 << k, 1 << (k + r)
    vals = rand_felts(rng, (ncols, n))
    salts = rand_felts(rng, (qpzk.SALT_SIZE, N)) if salted else None
    want = orc.batch_commit(vals, r, cap_h, salts=salts, threads=8)
    b = qpzk.PolynomialBatch.from_values(ctx, vals, r, cap_h, salts=salts)
    assert np.array_equal(b.cap, want["cap"])
    assert np.array_equal(b.polynomials, want["coeffs"])
    leaves, digests = b.export()
    assert np.array_equal(leaves, want["leaves"])
    assert np.array_equal(digests, want["digests"])
    # get_lde_values(i, step) and merkle openings
    step = 1 << r
    idx = np.array(sorted(set([0, n - 1] + [int(x) for x in rng.integers(0, n, 5)])), np.uint32)
    rows = b.get_lde_values(idx, step)
    for j, i in enumerate(idx):
        assert np.array_equal(rows[j], want["leaves"][bitrev(int(i) * step, k + r), :ncols])
    for leaf in sorted(set([0, N - 1] + [int(x) for x in rng.integers(0, N, 3)])):
        row, sib = b.open(leaf)
        assert np.array_equal(row, want["leaves"][leaf])
        assert orc.merkle_verify(row, leaf, want["cap"], sib)
    b.free()


@pytest.mark.parametrize("k,ncols,r,cap_h", [(4, 16, 3, 4), (12, 3, 3, 4), (14, 2, 3, 4)])
def test_from_coeffs_matches_oracle(ctx, k, ncols, r, cap_h):
    import qpzk
    rng = np.random.default_rng(2000 + k)
    coeffs = rand_felts(rng, (ncols, 1 << k))
    want = orc.batch_commit(coeffs, r, cap_h, is_coeffs=True, threads=8)
    b = qpzk.PolynomialBatch.from_coeffs(ctx, coeffs, r, cap_h)
    assert np.array_equal(b.cap, want["cap"])
    leaves, digests = b.export()
    assert np.array_equal(leaves, want["leaves"]) and np.array_equal(digests, want["digests"])
    b.free()


def test_noncanonical_inputs_are_reduced(ctx):
    import qpzk
    rng = np.random.default_rng(5)
    k, r, cap_h = 6, 3, 4
    vals = rand_felts(rng, (3, 1 << k))
    shifted = vals.copy()
    small = shifted < np.uint64(2**32 - 1)
    shifted[small] += np.uint64(P)                    # same field elements, non-canonical words
    a = qpzk.PolynomialBatch.from_values(ctx, vals, r, cap_h)
    b = qpzk.PolynomialBatch.from_values(ctx, shifted, r, cap_h)
    assert np.array_equal(a.cap, b.cap)
    a.free(); b.free()


def test_wormhole_shapes_match_oracle(ctx):
    """The four oracles of one wormhole proof at the non-ZK 2^13 shape: 84, 135, 20, 16 columns."""
    import qpzk
    rng = np.random.default_rng(13)
    for ncols in (84, 135, 20, 16):
        vals = rand_felts(rng, (ncols, 1 << 13))
        want = orc.batch_commit(vals, 3, 4, threads=8, want_leaves=False, want_digests=False)
        b = qpzk.PolynomialBatch.from_values(ctx, vals, 3, 4)
        assert np.array_equal(b.cap, want["cap"]), ncols
        b.free()


def test_microbench_full_size(ctx):
    """BASELINE config 3 at full size: 2^16 x 135, rate 3, cap 4 (SURVEY §8(d) inputs: SplitMix64 seed
    0x5eed0001). Cap compared with the oracle; plus size-independent properties."""
    import qpzk
    k, ncols, r, cap_h = 16, 135, 3, 4
    n, N = 1 << k, 1 << (k + r)
    vals = splitmix64(0x5EED0001, ncols * n).reshape(ncols, n)
    b = qpzk.PolynomialBatch.from_values(ctx, vals, r, cap_h)
    cap = b.cap
    want = orc.batch_commit(vals, r, cap_h, threads=16, want_leaves=False, want_digests=False)
    assert np.array_equal(cap, want["cap"])
    coeffs = b.polynomials
    assert np.array_equal(coeffs, want["coeffs"])
    rng = np.random.default_rng(99)
    # (1) opened rows are evaluations of the committed polynomials at g*w_N^i, and verify against the cap
    wN = orc.root_of_unity(k + r)
    for leaf in [0, N - 1] + [int(x) for x in rng.integers(0, N, 6)]:
        row, sib = b.open(leaf)
        assert orc.merkle_verify(row, leaf, cap, sib)
        x = orc.mul(GEN, orc.fpow(wN, bitrev(leaf, k + r)))
        for c in (0, 67, 134):
            acc = 0
            for coef in coeffs[c][::-1]:
                acc = (acc * x + int(coef)) % P
            assert int(row[c]) == acc
    # (2) values -> coeffs -> values round trip on the subgroup (rows i*8 of the LDE are NOT the trace: coset),
    #     so check interpolation directly on two columns
    for c in (1, 133):
        assert np.array_equal(orc.fft(coeffs[c]), vals[c])
    # (3) linearity of the LDE: commit(a) + commit(b) rows == commit(a+b) rows
    a2 = vals[:4]
    b2 = splitmix64(0x5EED0003, 4 * n).reshape(4, n)
    s2 = ((a2.astype(object) + b2.astype(object)) % P).astype(np.uint64)
    ba, bb, bs = (qpzk.PolynomialBatch.from_values(ctx, v, r, cap_h) for v in (a2, b2, s2))
    idx = rng.integers(0, N, 16).astype(np.uint32)
    ra, rb, rs = ba.get_lde_values(idx), bb.get_lde_values(idx), bs.get_lde_values(idx)
    assert np.array_equal(((ra.astype(object) + rb.astype(object)) % P).astype(np.uint64), rs)
    for x in (b, ba, bb, bs):
        x.free()


@pytest.mark.parametrize("k", [17, 19, 20])
def test_large_degrees_roundtrip_and_paths(ctx, k):
    """degree_bits above the microbench (up to 2^20 rows): interpolation round trip against the oracle's FFT
    on a few columns, LDE rows are evaluations of the committed polynomials, paths verify against the cap."""
    import qpzk
    ncols, r, cap_h = 3, 3, 4
    n, N = 1 << k, 1 << (k + r)
    vals = splitmix64(0x5EED0100 + k, ncols * n).reshape(ncols, n)
    b = qpzk.PolynomialBatch.from_values(ctx, vals, r, cap_h)
    coeffs = b.polynomials
    for c in range(ncols):
        assert np.array_equal(orc.fft(coeffs[c]), vals[c])
    cap = b.cap
    rng = np.random.default_rng(k)
    wN = orc.root_of_unity(k + r)
    lde0 = orc.coset_fft(np.concatenate([coeffs[0], np.zeros(N - n, np.uint64)]))   # natural order on g*<w_N>
    for leaf in [0, N - 1] + [int(x) for x in rng.integers(0, N, 6)]:
        row, sib = b.open(leaf)
        assert orc.merkle_verify(row, leaf, cap, sib)
        assert int(row[0]) == int(lde0[bitrev(leaf, k + r)])
    b.free()


@pytest.mark.parametrize("k", list(range(0, 19)))
def test_every_degree_matches_oracle(ctx, k):
    """Each transform length exercises a different stage decomposition of the radix-16 NTT (16 | 8 | 4 | 2
    remainders, one or two passes); compare coefficients and cap for every degree_bits 0..18."""
    import qpzk
    rng = np.random.default_rng(5000 + k)
    r = 3
    cap_h = min(4, k + r)
    vals = rand_felts(rng, (2, 1 << k))
    want = orc.batch_commit(vals, r, cap_h, threads=8, want_leaves=False, want_digests=False)
    b = qpzk.PolynomialBatch.from_values(ctx, vals, r, cap_h)
    assert np.array_equal(b.polynomials, want["coeffs"])
    assert np.array_equal(b.cap, want["cap"])
    b.free()


def test_batch_eval_ext_is_horner_in_the_extension(ctx):
    """`OpeningSet::new` for one oracle: p(zeta) for every committed polynomial, zeta in F_p^2 (X^2 = 7)."""
    import qpzk
    rng = np.random.default_rng(31)
    k, ncols = 9, 7
    vals = rand_felts(rng, (ncols, 1 << k))
    b = qpzk.PolynomialBatch.from_values(ctx, vals, 3, 4)
    coeffs = b.polynomials
    z = (int(rng.integers(0, P, dtype=np.uint64)), int(rng.integers(0, P, dtype=np.uint64)))
    got = b.eval_ext(np.array(z, np.uint64))
    for c in range(ncols):
        a0, a1 = 0, 0
        for coef in coeffs[c][::-1]:      # (a0 + a1 X) * (z0 + z1 X) + coef
            a0, a1 = (a0 * z[0] + 7 * a1 * z[1] + int(coef)) % P, (a0 * z[1] + a1 * z[0]) % P
        assert (int(got[c, 0]), int(got[c, 1])) == (a0, a1)
    b.free()


@pytest.mark.parametrize("bits,pos", [(8, 3), (16, 0), (16, 5)])
def test_fri_pow_returns_the_smallest_valid_witness(ctx, bits, pos):
    """`fri_proof_of_work`: permuting the state with the witness at `pos` gives >= `bits` leading zeros in
    output word 7; no smaller witness does (the reference takes whichever a rayon worker finds first)."""
    rng = np.random.default_rng(40 + bits + pos)
    state = rand_felts(rng, 12)
    w = ctx.fri_pow(state, pos, bits)

    def lz(cand):
        s = state.copy()
        s[pos] = np.uint64(cand)
        v = int(orc.poseidon(s)[7])
        return 64 - v.bit_length()

    assert lz(w) >= bits
    if bits <= 8:                          # exhaustive minimality check is cheap only for easy targets
        assert all(lz(c) < bits for c in range(w))
    else:
        # minimality through the GPU itself: every smaller candidate, permuted in one batch
        cands = np.tile(state, (w, 1))
        cands[:, pos] = np.arange(w, dtype=np.uint64)
        out7 = ctx.poseidon_permute(cands)[:, 7]
        assert not (out7 < np.uint64(1 << (64 - bits))).any()
