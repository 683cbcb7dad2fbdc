"""Pins the oracle's Merkle / transcript / vanishing-identity / FRI conventions against the proofs
the reference ships (SURVEY.md §8(c) pins P4-P8): the restated verifier must ACCEPT
wormhole/bench-data/proof.bin exactly as `verifier_verify_proof`
(/root/reference/wormhole/verifier/benches/verifier.rs:22-30) does, and reject tampered copies the way
/root/reference/wormhole/tests/src/verifier/verifier_tests.rs:48-91 expects."""
import numpy as np
import pytest

from helpers import common_bytes
from oracle import oracle as orc

BD = "wormhole/bench-data/"


def test_common_bytes_layout_is_byte_exact(ref_fixture):
    common = ref_fixture(BD + "common.bin")
    assert common_bytes(14, True, [4, 4, 4]) == common
    # verifier.bin = VerifierOnly (552 B) || common (verifier.rs:22-25)
    ver = ref_fixture(BD + "verifier.bin")
    assert ver[552:] == common


def test_bench_proof_is_accepted(ref_fixture):
    common, ver, proof = ref_fixture(BD + "common.bin"), ref_fixture(BD + "verifier.bin"), ref_fixture(BD + "proof.bin")
    assert len(proof) == 148932
    assert orc.proof_roundtrip(common, proof) == 1
    rc, ch = orc.verify(common, ver, proof)
    assert rc == 0
    # P5: first query index, also found by independent brute force in the survey
    assert int(ch["query_indices"][0]) == 34707
    assert ch["pow_response"] >> 48 == 0  # >= 16 leading zero bits
    assert 64 - int(ch["pow_response"]).bit_length() == 19


def test_tampered_bench_proof_is_rejected(ref_fixture):
    common, ver, proof = ref_fixture(BD + "common.bin"), ref_fixture(BD + "verifier.bin"), ref_fixture(BD + "proof.bin")
    rng = np.random.default_rng(7)
    seen = set()
    # a byte in each region: caps, openings, FRI caps / queries, final poly, pow witness, public inputs
    offsets = [3, 600, 1600, 4000, 6000, 20000, 80000, len(proof) - 300, len(proof) - 140, len(proof) - 100]
    offsets += [int(x) for x in rng.integers(0, len(proof), 20)]
    for off in offsets:
        bad = bytearray(proof)
        bad[off] ^= 1
        try:
            rc, _ = orc.verify(common, ver, bytes(bad))
        except RuntimeError:  # non-canonical element / bad shape
            rc = -1
        assert rc != 0, "tampering byte %d was not detected" % off
        seen.add(rc)
    assert len(seen) >= 3  # different checks fire for different regions


def _derive_index(leaf, path, cap):
    """Find the leaf index whose direction bits make `path` end in one of `cap`'s entries."""
    cands = {0: orc.hash_or_noop(leaf)}
    for lvl, sib in enumerate(path):
        nxt = {}
        for idx, h in cands.items():
            nxt[idx] = orc.two_to_one(h, sib)
            nxt[idx | (1 << lvl)] = orc.two_to_one(sib, h)
        cands = nxt
    hits = []
    for idx, h in cands.items():
        for c in range(cap.shape[0]):
            if np.array_equal(h, cap[c]):
                hits.append(idx | (c << len(path)))
    return hits


@pytest.mark.parametrize("name", ["dummy_proof.bin", "dummy_proof_zk.bin"])
def test_dummy_proofs_merkle_paths(ref_fixture, name):
    """P8: the two non-ZK 2^13-row proofs. Their verifier data is not shipped, so only what is
    self-contained is checked: every opened row / FRI coset verifies against the caps inside the
    proof, at one consistent query index per round."""
    proof = ref_fixture("wormhole/aggregator/data/" + name)
    assert len(proof) == 132712
    common = common_bytes(13, False, [4, 4])
    assert orc.proof_roundtrip(common, proof) == 1
    # parse the bits we need by hand: caps at the front, then openings, FRI caps, queries
    a = np.frombuffer(proof, dtype=np.uint8)
    u64 = lambda off, n: np.frombuffer(a[off:off + 8 * n].tobytes(), dtype="<u8").astype(np.uint64)
    wires_cap = u64(0, 64).reshape(16, 4)
    off = 3 * 512 + (4 + 80 + 135 + 2 + 2 + 18 + 16) * 16 + 2 * 512
    indices = []
    for q in range(28):
        o = off
        o += 84 * 8 + 1 + 12 * 32                      # constants_sigmas row + path
        row = u64(o, 135)
        o += 135 * 8
        assert a[o] == 12
        path = u64(o + 1, 48).reshape(12, 4)
        hits = _derive_index(row, path, wires_cap)
        assert len(hits) == 1
        indices.append(hits[0])
        off += (84 + 135 + 20 + 16) * 8 + 4 * (1 + 12 * 32) + 2 * 32 * 8 + (1 + 8 * 32) + (1 + 4 * 32)
    assert indices[0] == {"dummy_proof.bin": 34670, "dummy_proof_zk.bin": 9643}[name]
    assert orc.check_proof_paths(common, proof, indices) == 28 * 5
