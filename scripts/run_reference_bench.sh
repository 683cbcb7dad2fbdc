#!/usr/bin/env bash
# The reference's OWN benchmarks, to be run where a Rust toolchain and the crates.io
# dependencies exist - neither is in this repository's build image, so this script is shipped unrun.
#
#   scripts/run_reference_bench.sh /path/to/qp-zk-circuits-rm [threads]
#
# It times, with criterion, exactly what the reference ships:
#   wormhole/prover/benches/prover.rs:11-30        prover_create_proof  (WormholeProver::new + commit + prove:
#                                                  the circuit build is INSIDE the timed closure, lines 15-18)
#   wormhole/verifier/benches/verifier.rs:13-33    verifier_verify_proof
#   wormhole/aggregator/benches/aggregator.rs:23-141  aggregate_proofs / verify_aggregate_proof
# and prints the numbers next to which bench.py's lines belong:
#   bench.py `value` / `e2e`             <->  1 / (prover_create_proof time - circuit build time): bench.py times
#                                             prove() AFTER witness generation, on a prebuilt circuit
#   bench.py `aggregation_tree`          <->  aggregate_proofs (8 leaves, branching factor 2)
#   bench.py --impl reference            <->  this script's prover number (the oracle port stands in for it here)
set -euo pipefail
REF=${1:?path to a checkout of aletheia-labs/qp-zk-circuits-rm}
THREADS=${2:-$(nproc)}
export RAYON_NUM_THREADS=$THREADS
export RUSTFLAGS="${RUSTFLAGS:-} -C target-cpu=native"
echo "# host: $(nproc) cores, RAYON_NUM_THREADS=$RAYON_NUM_THREADS, RUSTFLAGS=$RUSTFLAGS"
grep -m1 'model name' /proc/cpuinfo || true
cd "$REF"
cargo --version
# Cargo.lock pins qp-plonky2 1.1.1 / qp-plonky2-field 1.1.1 (Cargo.lock:489-490, 514-515)
cargo bench --locked -p qp-wormhole-prover -- --noplot
cargo bench --locked -p qp-wormhole-verifier -- --noplot
cargo bench --locked -p qp-wormhole-aggregator -- --noplot
# per-stage breakdown of one proof (qp-plonky2's TimingTree), to set against bench.py's proof_stage_ms
RUST_LOG=debug cargo run --locked --release -p qp-wormhole-example 2>&1 | grep -E "prove|commit|quotient|FRI|fri|Merkle|total" || true
echo "# criterion reports are under $REF/target/criterion/{prover_create_proof,verifier_verify_proof,aggregate_proofs}"
