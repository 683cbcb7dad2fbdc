"""Test shim: the synthetic circuit builder (qpzk.synth) backed by the ORACLE for its few hashes, so
CPU-only tests never touch the GPU library."""
from oracle import oracle as orc
from qpzk import synth
from qpzk.synth import ARITHMETIC, BASE_SUM, CONSTANT, NOOP, POSEIDON, PUBLIC_INPUT  # noqa: F401


class OracleProvider:
    poseidon_tables = staticmethod(orc.poseidon_tables)
    hash_no_pad = staticmethod(orc.hash_no_pad)


def build(degree_bits, zk=False, seed=1, **kw):
    return synth.build(degree_bits, zk=zk, seed=seed, provider=OracleProvider(), **kw)


def build_recursion(degree_bits, zk=False, seed=1, **kw):
    return synth.build_recursion(degree_bits, zk=zk, seed=seed, provider=OracleProvider(), **kw)
