"""`write_polynomial_batch` / `read_polynomial_batch`: the serialized
`constants_sigmas_commitment` a prover restored from files carries
(/root/reference/wormhole/prover/src/lib.rs:105-187) goes to the device as stored, and a circuit created from it
proves the same bytes as one created from the value columns. The byte layout is restated from upstream plonky2
(no serialized ProverOnlyCircuitData ships with the reference): GPU bytes == the oracle's restatement, round
trips are exact, and bytes that are not a commitment are refused when verification is asked for."""
import numpy as np
import pytest

from helpers import rand_felts
from oracle import oracle as orc
from qpzk import synth


def _layout_len(ncols, width, k, r, h):
    n, N = 1 << k, 1 << (k + r)
    return 8 + ncols * (8 + 8 * n) + 8 + N * (8 + 8 * width) + 8 + 32 * (2 * N - 2 * (1 << h)) + 8 + (32 << h) + 17


@pytest.mark.parametrize("k,ncols,r,h,salted", [(3, 2, 1, 0, False), (5, 7, 3, 4, True), (4, 3, 2, 6, False)])
def test_oracle_layout(k, ncols, r, h, salted):
    rng = np.random.default_rng(k)
    vals = rand_felts(rng, (ncols, 1 << k))
    salts = rand_felts(rng, (4, 1 << (k + r))) if salted else None
    want = orc.batch_commit(vals, r, h, salts=salts)
    raw = orc.batch_to_bytes(want, r, salted)
    assert len(raw) == _layout_len(ncols, ncols + (4 if salted else 0), k, r, h)
    # the fields sit where `read_polynomial_batch` looks for them
    w = np.frombuffer(raw[:-1], "<u8")
    n, N, width = 1 << k, 1 << (k + r), ncols + (4 if salted else 0)
    assert w[0] == ncols and w[1] == n and np.array_equal(w[2:2 + n], want["coeffs"][0])
    off = 1 + ncols * (n + 1)
    assert w[off] == N and w[off + 1] == width and np.array_equal(w[off + 2:off + 2 + width], want["leaves"][0])
    off += 1 + N * (width + 1)
    nd = 2 * (N - (1 << h))
    assert w[off] == nd
    off += 1 + 4 * nd
    assert w[off] == h and np.array_equal(w[off + 1:off + 1 + (4 << h)].reshape(-1, 4), want["cap"])
    assert list(w[-2:]) == [k, r] and raw[-1] == (1 if salted else 0)


@pytest.fixture(scope="module")
def ctx():
    import qpzk
    c = qpzk.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k,ncols,r,h,salted", [(3, 2, 1, 0, False), (6, 9, 3, 4, True), (10, 84, 3, 4, False),
                                                (5, 3, 2, 7, False), (13, 5, 3, 4, True)])
def test_gpu_bytes_equal_oracle_and_round_trip(ctx, k, ncols, r, h, salted):
    import qpzk
    rng = np.random.default_rng(100 + k)
    vals = rand_felts(rng, (ncols, 1 << k))
    salts = rand_felts(rng, (4, 1 << (k + r))) if salted else None
    b = qpzk.PolynomialBatch.from_values(ctx, vals, r, h, salts=salts)
    raw = b.to_bytes()
    want = orc.batch_commit(vals, r, h, salts=salts, threads=4)
    assert raw == orc.batch_to_bytes(want, r, salted)
    # nothing recomputed on the way back in; the handle serves the same commitment
    for verify in (False, True):
        b2, used = qpzk.PolynomialBatch.from_bytes(ctx, raw + b"trailing", verify=verify)
        assert used == len(raw)
        assert (b2.ncols, b2.salt_cols, b2.degree_bits, b2.rate_bits, b2.cap_height) == (ncols, 4 if salted else 0, k, r, h)
        assert np.array_equal(b2.cap, want["cap"]) and np.array_equal(b2.polynomials, want["coeffs"])
        for leaf in (0, (1 << (k + r)) - 1, 5 % (1 << (k + r))):
            row, sib = b2.open(leaf)
            assert np.array_equal(row, want["leaves"][leaf]) and orc.merkle_verify(row, leaf, want["cap"], sib)
        idx = np.arange(0, 1 << k, max(1, (1 << k) // 4), dtype=np.uint32)
        assert np.array_equal(b2.get_lde_values(idx, 1 << r), b.get_lde_values(idx, 1 << r))
        assert b2.to_bytes() == raw
        b2.free()
    b.free()


@pytest.mark.gpu
def test_gpu_refuses_malformed_bytes(ctx):
    import qpzk
    rng = np.random.default_rng(7)
    k, ncols, r, h = 6, 5, 3, 4
    b = qpzk.PolynomialBatch.from_values(ctx, rand_felts(rng, (ncols, 1 << k)), r, h)
    raw = bytearray(b.to_bytes())
    b.free()
    n, N = 1 << k, 1 << (k + r)

    def refused(data, verify=False):
        with pytest.raises(qpzk.QpzkError) as e:
            qpzk.PolynomialBatch.from_bytes(ctx, bytes(data), verify=verify)
        assert e.value.code == -1
    for cut in (0, 7, 8, 100, 8 + ncols * 8 * (n + 1) + 3, len(raw) - 1, len(raw) - 17, len(raw) - 18 - (32 << h)):
        refused(raw[:cut])
    bad = bytearray(raw); bad[0:8] = (0).to_bytes(8, "little"); refused(bad)                      # no polynomials
    bad = bytearray(raw); bad[8:16] = (n + 1).to_bytes(8, "little"); refused(bad)                 # length not a power of two
    bad = bytearray(raw); bad[8:16] = (1 << 40).to_bytes(8, "little"); refused(bad)               # absurd length
    off = 8 + 8 * (n + 1)
    bad = bytearray(raw); bad[off:off + 8] = (n // 2).to_bytes(8, "little"); refused(bad)         # ragged polynomials
    off = 8 + ncols * 8 * (n + 1)
    bad = bytearray(raw); bad[off:off + 8] = (N * 2).to_bytes(8, "little"); refused(bad)          # more leaves than bytes
    bad = bytearray(raw); bad[off + 8:off + 16] = (ncols + 1).to_bytes(8, "little"); refused(bad)  # leaf width
    off2 = off + 8 + 8 * (ncols + 1) * 3
    bad = bytearray(raw); bad[off2:off2 + 8] = (ncols - 1).to_bytes(8, "little"); refused(bad)    # one short leaf
    bad = bytearray(raw); bad[-1] = 1; refused(bad)                                               # blinding without salts
    bad = bytearray(raw); bad[-9:-1] = (r + 1).to_bytes(8, "little"); refused(bad)                # rate_bits contradicts N
    offd = off + 8 + N * 8 * (ncols + 1)
    bad = bytearray(raw); bad[offd:offd + 8] = (2 * N - 4).to_bytes(8, "little"); refused(bad)    # digest count vs cap height
    # well-formed, but not a commitment: accepted on trust (as plonky2 does), refused when verification is on
    for where in (20, off + 8 + 8 * (ncols + 1) * 9 + 16, offd + 8 + 40, len(raw) - 17 - 8):
        bad = bytearray(raw)
        bad[where] ^= 1
        b2, _ = qpzk.PolynomialBatch.from_bytes(ctx, bytes(bad))
        b2.free()
        refused(bad, verify=True)
    bad = bytearray(raw); bad[16:24] = (0xFFFFFFFFFFFFFFFF).to_bytes(8, "little"); refused(bad, verify=True)  # non-canonical


@pytest.mark.gpu
@pytest.mark.parametrize("k,zk", [(6, False), (9, True)])
def test_circuit_from_commitment_proves_the_same_bytes(ctx, k, zk):
    import qpzk

    class _Prov:
        poseidon_tables = staticmethod(orc.poseidon_tables)
        hash_no_pad = staticmethod(orc.hash_no_pad)

    circ = synth.build(k, zk=zk, seed=40 + k, provider=_Prov())
    gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
    blob = gc.commitment_bytes()
    r, h = gc.info["rate_bits"], gc.info["cap_height"]
    assert blob == orc.batch_to_bytes(orc.batch_commit(circ["constants_sigmas"], r, h, threads=4), r, False)
    launches = ctx.launch_count()
    gf = qpzk.Circuit(ctx, circ["common"], circ["digest"], commitment=blob)
    # two copies, the inverse transform and its read-back, plus per-circuit tables: no LDE, no hashing
    assert ctx.launch_count() - launches <= 8
    assert np.array_equal(gf.constants_sigmas_cap, gc.constants_sigmas_cap)
    assert gf.verifier_only_bytes() == gc.verifier_only_bytes()
    salts = circ["salts"] if zk else None
    p1 = gc.prove(circ["wires"], circ["public_inputs"], salts)
    p2 = gf.prove(circ["wires"], circ["public_inputs"], salts)
    assert p1 == p2
    rc, _ = orc.verify(circ["common"], gf.verifier_only_bytes(), p2)
    assert rc == 0
    assert gf.commitment_bytes() == blob
    # a commitment of another shape is refused
    other = qpzk.PolynomialBatch.from_values(ctx, circ["constants_sigmas"][:5], r, h)
    with pytest.raises(qpzk.QpzkError):
        qpzk.Circuit(ctx, circ["common"], circ["digest"], commitment=other.to_bytes())
    other.free()
    gc.free()
    gf.free()
