#!/usr/bin/env python3
"""Copies the reference's shipped binary fixtures for the hot path into tests/golden/ (data, not source)
and records where each came from. They are the golden vectors that pin the oracle (SURVEY.md §8(c)
P4-P8) and, through tests/test_gpu_reference_fixtures.py, the CUDA Poseidon / Merkle path directly.
Run in the build container, where /root/reference exists."""
import hashlib
import json
import os
import shutil

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
FILES = {
    "bench_common.bin": "wormhole/bench-data/common.bin",
    "bench_verifier.bin": "wormhole/bench-data/verifier.bin",
    "bench_proof.bin": "wormhole/bench-data/proof.bin",
    "dummy_proof.bin": "wormhole/aggregator/data/dummy_proof.bin",
    "dummy_proof_zk.bin": "wormhole/aggregator/data/dummy_proof_zk.bin",
}
USED_BY = {
    "bench_common.bin": "wormhole/verifier/benches/verifier.rs:22-25 (CommonCircuitData)",
    "bench_verifier.bin": "wormhole/verifier/benches/verifier.rs:22-25 (VerifierOnlyCircuitData || common)",
    "bench_proof.bin": "wormhole/verifier/benches/verifier.rs:26-30 (the proof `verifier_verify_proof` accepts)",
    "dummy_proof.bin": "wormhole/aggregator/src/util.rs:6-9 (include_bytes!, feature no_zk)",
    "dummy_proof_zk.bin": "wormhole/aggregator/src/util.rs:6-9 (include_bytes!)",
}


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for name, rel in FILES.items():
        src = os.path.join(REF, rel)
        dst = os.path.join(OUT, name)
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        data = open(dst, "rb").read()
        manifest[name] = {"reference_path": rel, "bytes": len(data), "sha256": hashlib.sha256(data).hexdigest(),
                          "read_by_reference_at": USED_BY[name]}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", len(manifest), "fixtures to", OUT)


if __name__ == "__main__":
    main()
