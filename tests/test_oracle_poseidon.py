"""Pins the CPU oracle's Poseidon / sponge against every known answer the reference's tests hold
for this path (SURVEY.md §8(c) pins P1-P3) and the published permutation KATs (App. A.2)."""
import hashlib

import numpy as np

from oracle import oracle as orc

P = orc.P


def bytes4_to_felts(b):  # injective_bytes_to_felts, /root/reference/common/src/utils.rs:163-176
    out = []
    for i in range(0, len(b), 4):
        out.append(int.from_bytes(b[i:i + 4].ljust(4, b"\0"), "little"))
    return out


def digest_to_felts(b):  # digest_bytes_to_felts, /root/reference/common/src/utils.rs:192-204
    return [int.from_bytes(b[i:i + 8], "little") for i in range(0, 32, 8)]


def felts_to_digest(f):
    return b"".join(int(x).to_bytes(8, "little") for x in f)


def string_to_felts(s):  # injective_string_to_felt, /root/reference/common/src/utils.rs:144-160
    b = s.encode()
    assert len(b) == 8
    return [int.from_bytes(b[0:4], "little"), int.from_bytes(b[4:8], "little")]


def u64_to_felts(v):  # /root/reference/common/src/utils.rs:126-131
    return [(v >> 32) & 0xFFFFFFFF, v & 0xFFFFFFFF]


def u128_to_felts(v):  # /root/reference/common/src/utils.rs:104-115
    return [(v >> (96 - 32 * i)) & 0xFFFFFFFF for i in range(4)]


def test_selfcheck_and_constant_digest():
    assert orc.selfcheck() == 0
    t = orc.poseidon_tables()
    # SHA-256 of the 360 round constants as little-endian u64
    assert hashlib.sha256(t["rc"].astype("<u8").tobytes()).hexdigest() == \
        "d2fcbb5be293c50ab4b1ddcd9c81005b12d689816a54c91a054f97f6588a20a8"
    assert all(int(x) < 0xfffeeac900011537 for x in t["rc"])


def test_permutation_kats():
    z = orc.poseidon(np.zeros(12, np.uint64))
    assert [int(x) for x in z[:4]] == [0x3c18a9786cb0b359, 0xc4055e3364a246c3, 0x7953db0ab48808f4,
                                       0xc71603f33a1144ca]
    assert int(z[11]) == 0x1792b1c4342109d7
    r = orc.poseidon(np.arange(12, dtype=np.uint64))
    assert int(r[0]) == 0xd64e1e3efc5b8e9e and int(r[11]) == 0x5c0a27fcb0e1459b
    m = orc.poseidon(np.full(12, P - 1, np.uint64))
    assert int(m[0]) == 0xbe0085cfc57a8357
    h = orc.hash_no_pad(np.arange(1, 10, dtype=np.uint64))
    assert [int(x) for x in h] == [0x5a90f7c562413c2b, 0xa1874b91e26076d4, 0x37b5cd4fe1fb94da,
                                   0x3db54acf2fa3b131]


def test_fast_equals_naive_random():
    rng = np.random.default_rng(1)
    for _ in range(200):
        s = rng.integers(0, P, 12, dtype=np.uint64)
        assert np.array_equal(orc.poseidon(s), orc.poseidon(s, naive=True))


# P1: /root/reference/wormhole/tests/src/circuit/unspendable_account_tests.rs:12-27
SECRETS = [
    "cd94df2e3c38a87f3e429b62af022dbe4363143811219d80037e8798b2ec9229",
    "8b680b2421968a0c1d3cff6f3408e9d780157ae725724a78c3bc0998d1ac8194",
    "87f5fc11df0d12f332ccfeb92ddd8995e6c11709501a8b59c2aaf9eefee63ec1",
    "ef69da4e3aa2a6f15b3a9eec5e481f17260ac812faf1e685e450713327c3ab1c",
    "9aa84f99ef2de22e3070394176868df41d6a148117a36132d010529e19b018b7",
]
ADDRESSES = [
    "582d3b97e9b09c7776921d3ead2d8186e3aa199cf8d63f5d014e65d04ac80f26",
    "b0807446c24263def407aa8328400fef981ec30fc8453d7adbcc57bcf8af3bbf",
    "ac081f035cc995574fef749f33b455c31cb02759932d01b6367ab852bb5599ac",
    "a5073c13573f10552c37f35080dc0118bda22f1217381611cf4644909377ce05",
    "73378f4b54f48a38b17073e08440531594f2b771ceefc5c3cd621e1309fbe927",
]


def test_unspendable_account_kats():
    # UnspendableAccount::from_secret, /root/reference/wormhole/circuit/src/unspendable_account.rs:40-64
    for secret, address in zip(SECRETS, ADDRESSES):
        pre = string_to_felts("wormhole") + bytes4_to_felts(bytes.fromhex(secret))
        assert len(pre) == 10
        inner = orc.hash_no_pad(pre)
        outer = orc.hash_no_pad(inner)
        assert felts_to_digest(outer).hex() == address


# P2: /root/reference/wormhole/tests/src/prover/prover_tests.rs:29-43 with inputs from
# /root/reference/wormhole/tests/test-helpers/src/lib.rs:10-23
DEFAULT_SECRET = "4c8587bd422e01d961acdc75e7d66f6761b7af7c9b1864a492f369c9d6724f05"
NULLIFIER_BYTES = bytes([169, 76, 150, 35, 66, 248, 76, 193, 57, 204, 106, 33, 169, 160, 248, 113, 235,
                         144, 212, 48, 9, 232, 146, 7, 105, 125, 170, 24, 33, 54, 135, 28])


def test_nullifier_kat():
    # Nullifier::from_preimage, /root/reference/wormhole/circuit/src/nullifier.rs:53-75
    pre = string_to_felts("~nullif~") + bytes4_to_felts(bytes.fromhex(DEFAULT_SECRET)) + u64_to_felts(4)
    assert len(pre) == 12
    h = orc.hash_no_pad(orc.hash_no_pad(pre))
    assert felts_to_digest(h) == NULLIFIER_BYTES


# P3: /root/reference/wormhole/tests/test-helpers/src/lib.rs:68-80
ROOT_HASH = "5ffa2ab5b0db9883b22b1e5810932ea9d9eab1840730fd39ace71c26bb8d082d"
STORAGE_PROOF = [
    "0000000000000020bfb500000000000020000000000000005d7c4eb0b2a8bb01872f88950f8c736fc72a250c32b4bdad9a50e7b5163a27aa20000000000000008f6440ed6cd23d75bfdd64b70ec7b0c969bd03e53f9fc1df688f8538dad89f402000000000000000545576a55a3f69e109b776d252064d3c9bf2fd3a0cd0447c8d82ec12b0343f3a20000000000000000f3ed746dd90e0e2a0d3f8faf0b8a41d5fafd9edcbc88630e389f2db76dd44b7200000000000000091c3eead5530405e48b8df6453a60be878eb1fa46c2a95638cdec8c8d722b46020000000000000008475575039b5b19da2901935792d5b1d5f9a09e08065e4d27a438329710120002000000000000000e6f538f42cbc6e72d6a302a648da34c475bcfa104e7cb80625fcf3219bd12172200000000000000056c6d22ef15fbb6005782db4c357b38cb53f5d39e5d8abdb3efffaec0537381420000000000000007f7b9a72037f9305f49bb2c25aa2f2c0108753ae606e1f094e887071e2596cfb2000000000000000805a0b660043743ecac1396810e2c3664e5f6bd54890cfc4eb04d914a38a32ba2000000000000000a22c86fb54dbd5c704fc4d849c715109d7cb3167b0eb2ed270ca658bd9dcca2a20000000000000003687179c5ce1cb12b50e50d421bcbdceb82ec583de7585fb7898e167108168b5",
    "000000000000002004100000000000002000000000000000508b02bea5f6ec0560cb2cbfda44d44ee4ea671f5f3cbb5d27b90e6afcafa1f32000000000000000b7361080961b2d3b348d96affbf10c7ee2d6416efa14b524289e264863a270b6",
    "1e00000000000020261276cc9d1f8598ea4b6a74b15c2f003280000000000000200000000000000036eed7029a2181549ea0a84a554dd682b0184a06f1c56a53ebf70c127123252920000000000000001961560d112cfd667e09610793793d3fc2ee32eb87171773c2e4c6e1473f400b2000000000000000b5e25bb2727a369c7a991e657eb15e8a578a30b89088ba5cf5c588deaee3a9f5200000000000000016b14e363d6ed03d0f13adc683dab364d051a8394db2f605adfe69d0ef5dd78a",
    "000000000000002084000000000000002000000000000000c58635f106880ea6ac74b554a030a74e08587a15fe9cca1117415c1f086613e62000000000000000abf9dfa05f2adc8c6b9447a6dae41d898ac8d77d683c8fe8c9a563a0cd05e0d7",
    "1e00000000000020857e7ea49e785c4e3e1f77a710cfc20085eb00000000000020000000000000007f6a20004a9e9c8534de8e4a017e3795c9d8a30e036108eb593d2ac31f6a34e42000000000000000baf5a768ed92d1ac1cead4bcee891151641cfb6b109c9b6075952a36e5808dfc20000000000000006e19211b4ff0a3feb43b34373129676d22378dfe1303191a96b34012713b65832000000000000000f6885f81a0d9ee08a3a67c4f2ef71a2ec725c8a9c79599eb975c2319e4aae5e920000000000000008d4b3c32ff1324fe3b7a05467e88e9f69b0df523bc3b6fbfdc888f06401bc9e72000000000000000ea72cebf4e99ec5a02713c47fa3198ea718fabce8eaf27707c3ec03eafa34174200000000000000077c5198a04b75c9795fe20a45d68df141ef53182a243c6102607da94ee03a9a82000000000000000ee55785e535fe32542b8b7f8537d8f921df34012c8f8dfd97087159ac05b99d1200000000000000013da88523a40420379a2776f484740dd9e78e858b11c7f43d5db16dc923b5e71",
    "0000000000000020a0000000000000002000000000000000439f73a9fe5a17162de32efd7abca06f0c880dc966613afdcf1ab350e1619c4a2000000000000000797b157cc18a8d60054cf9e008630ef8642b335fe0869a9796b5feb0f464ff4b",
    "3e0000000000003000e339aa4f999f6414fef6d1a1eae663e1cbc7ba7fe5fd365ea504b46241cddf0000000000000000",
]
STORAGE_INDICES = [768, 48, 240, 48, 160, 128, 16]
FUNDING_ACCOUNT = bytes([226, 124, 203, 9, 80, 60, 124, 205, 165, 5, 178, 216, 195, 15, 149, 38, 116, 1, 238,
                         133, 181, 154, 106, 17, 41, 228, 118, 179, 82, 141, 225, 76])
TO_ACCOUNT = bytes([162, 77, 187, 9, 249, 178, 185, 87, 194, 50, 198, 98, 179, 134, 179, 126, 123, 21, 247,
                    44, 50, 216, 140, 243, 97, 177, 13, 94, 26, 255, 19, 170])
FUNDING_AMOUNT = int.from_bytes(bytes([0, 16, 165, 212, 232, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]), "little")


def test_storage_proof_hash_chain():
    # What the StorageProof fragment constrains, /root/reference/wormhole/circuit/src/storage_proof/mod.rs:169-243:
    # every node (188 felts, zero padded) hashes to the digest embedded in its parent.
    prev = digest_to_felts(bytes.fromhex(ROOT_HASH))
    for node_hex, idx in zip(STORAGE_PROOF, STORAGE_INDICES):
        felts = bytes4_to_felts(bytes.fromhex(node_hex))
        assert len(felts) <= 188
        padded = felts + [0] * (188 - len(felts))
        h = [int(x) for x in orc.hash_no_pad(padded)]
        assert h == prev
        j = idx // 8
        prev = [padded[j + 2 * k] + (padded[j + 2 * k + 1] << 32) for k in range(4)]
    # leaf-inputs hash: last three felts must match (first nibble is not stored), mod.rs:226-231
    leaf = u64_to_felts(4) + digest_to_felts(FUNDING_ACCOUNT) + digest_to_felts(TO_ACCOUNT) + \
        u128_to_felts(FUNDING_AMOUNT)
    assert len(leaf) == 14  # 2 + 4 + 4 + 4 (leaf.rs:40-47; NUM_LEAF_INPUT_FELTS = 11 at leaf.rs:15 is stale)
    lh = [int(x) for x in orc.hash_no_pad(leaf)]
    assert lh[1:] == prev[1:]
