/* qpzk — B200-native (sm_100a) backend for the Plonky2 polynomial-commitment / quotient / FRI hot
 * path of the qp-wormhole prover, voting circuit and aggregator.
 *
 * This is the drop-in boundary: a flat C ABI (plain pointers and sizes) that a `cc`-built Rust
 * `-sys` crate binds, so that a `[patch.crates-io]` copy of qp-plonky2 1.1.1 can forward the
 * bodies of the functions below to the GPU while every signature the reference touches stays as it
 * is. The reference contains no FFI of its own (it calls qp-plonky2 as plain Rust, pinned at
 * /root/reference/Cargo.lock:489-490), so each entry point cites the qp-plonky2 function whose
 * body it replaces and the reference call site that reaches it. INTEGRATION.md shows the Rust
 * bindings.
 *
 * Conventions
 *  - All field elements are u64. Inputs may be any representative (< 2^64); outputs are canonical
 *    (< p = 2^64 - 2^32 + 1). Extension elements are two consecutive u64 (a + bX, X^2 = 7).
 *    Hashes / digests are four consecutive u64.
 *  - Host pointers are borrowed for the duration of the call and never retained; `_dev` entry
 *    points take device pointers valid on the context's device.
 *  - Handles own device memory and are released with their *_free function.
 *  - Every function returns QPZK_OK (0) or a negative qpzk_status; nothing throws or aborts across
 *    the ABI (every entry point that allocates runs inside an exception guard). `qpzk_last_error` returns a
 *    human-readable message for the calling thread.
 *  - Re-entrant: one qpzk_ctx per calling thread (the aggregator proves chunks concurrently from
 *    rayon workers, /root/reference/wormhole/aggregator/src/circuits/tree.rs:92-103). A context
 *    owns one CUDA stream; calls on different contexts overlap on the device. Handles (batches, trees,
 *    circuits, FRI provers) belong to the context they were created on and must not be used from two
 *    threads at once; a qpzk_circuit carries ONE proof at a time (its calls are serialised by a mutex).
 *  - Device memory comes from the library's own stream-ordered pool (one per device); the device's default
 *    pool and its attributes are not touched.
 *  - There is NO CPU fallback: if no CUDA device is usable, qpzk_ctx_create fails.
 */
#ifndef QPZK_H
#define QPZK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum qpzk_status {
  QPZK_OK = 0,
  QPZK_ERR_BAD_ARG = -1,
  QPZK_ERR_CUDA = -2,
  QPZK_ERR_OOM = -3,
  QPZK_ERR_NOT_DIVISIBLE = -4, /* reserved: the prover does not test divisibility by Z_H (an unsatisfied
                                  witness yields a proof the verifier rejects, as in plonky2 release builds) */
  QPZK_ERR_UNSUPPORTED = -5
} qpzk_status;

typedef struct qpzk_ctx qpzk_ctx;
typedef struct qpzk_batch qpzk_batch; /* PolynomialBatch: coefficients + LDE + Merkle tree, on device */
typedef struct qpzk_tree qpzk_tree;   /* MerkleTree over caller-provided leaves, on device */

#define QPZK_SALT_SIZE 4 /* plonky2 SALT_SIZE: blinding columns appended to each hiding oracle */
/* qpzk_prove flags */
#define QPZK_PROVE_TRACE 1u        /* keep intermediates for qpzk_prove_trace */
#define QPZK_PROVE_DEVICE_INPUTS 2u /* wires / salts are device pointers */
#define QPZK_PROVE_SEEDED_SALTS 4u  /* salts drawn on the device from a 32-byte seed */

/* Stage indices for qpzk_ctx_stage_ms (CUDA-event time of the last commit on this context). */
enum {
  QPZK_STAGE_H2D = 0,
  QPZK_STAGE_IFFT = 1,
  QPZK_STAGE_LDE = 2,
  QPZK_STAGE_LEAF_HASH = 3,
  QPZK_STAGE_MERKLE_LEVELS = 4,
  QPZK_STAGE_D2H = 5,
  QPZK_NUM_STAGES = 6
};

/* ---- context ---- */
/* flags: how a call waits for its stream (a proof is one stream-ordered enqueue - the Fiat-Shamir transcript
 * runs on the device - so qpzk_prove waits once, at the end).
 * Default: spin on a host core - lowest latency, right when proving threads <= cores.
 * QPZK_CTX_BLOCKING_SYNC - sleep on a blocking-sync event (interrupt wake-up, tens of microseconds each).
 * QPZK_CTX_YIELD_SYNC    - poll the event and sched_yield() between polls: spinning latency while cores are
 *                          free, and proving threads x processes may exceed the host cores (many rayon
 *                          workers, 8 ranks x 6 proofs in flight on 32 cores) without starving each other. */
#define QPZK_CTX_BLOCKING_SYNC 1u
#define QPZK_CTX_YIELD_SYNC 2u
int qpzk_ctx_create(int device, uint32_t flags, qpzk_ctx** out);
void qpzk_ctx_destroy(qpzk_ctx* ctx);
const char* qpzk_last_error(void);
int qpzk_ctx_sync(qpzk_ctx* ctx);
/* CUDA stream owned by the context (a cudaStream_t), for callers that time with their own events. */
void* qpzk_ctx_stream(qpzk_ctx* ctx);
/* Per-stage device time in milliseconds of the most recent commit on this context. */
int qpzk_ctx_stage_ms(qpzk_ctx* ctx, float* out_ms /* [QPZK_NUM_STAGES] */);
/* Number of kernel launches issued by this context since creation. */
uint64_t qpzk_ctx_launch_count(const qpzk_ctx* ctx);

/* Pinned host memory for callers that want full-speed transfers (optional: 55 GB/s on the B200 boxes).
 * Host buffers handed to any entry point may also be ordinary pageable memory (a Rust Vec, a numpy array). From
 * 32 MB up such a buffer is uploaded through pinned chunks owned by the context and filled by four host threads
 * (28-35 GB/s instead of the driver's 11-12 GB/s; environment: QPZK_H2D_THREADS, 0 or 1 = always the driver's
 * path; QPZK_H2D_MIN_MB). The buffer has been read completely when the call returns. */
int qpzk_host_alloc(size_t bytes, void** out);
void qpzk_host_free(void* p);
/* Plain device memory helpers (used by the bench harness and tests for the `_dev` entry points). */
int qpzk_dev_alloc(qpzk_ctx* ctx, size_t bytes, void** out);
void qpzk_dev_free(qpzk_ctx* ctx, void* p);
int qpzk_memcpy_h2d(qpzk_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int qpzk_memcpy_d2h(qpzk_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* ---- hashing: PoseidonPermutation / PoseidonHash (qp-plonky2 hash/poseidon.rs, hashing.rs) ----
 * Reference call sites: /root/reference/wormhole/circuit/src/nullifier.rs:64-65,
 * /root/reference/wormhole/circuit/src/unspendable_account.rs:54-56. */
/* n independent permutations of 12-element states, in place. */
int qpzk_poseidon_permute(qpzk_ctx* ctx, uint64_t* states /* host [n][12] */, uint64_t n);
/* n independent `hash_no_pad` of `len` elements each (overwrite-mode sponge, rate 8). */
int qpzk_hash_no_pad(qpzk_ctx* ctx, const uint64_t* inputs /* host [n][len] */, uint64_t n, uint32_t len,
                     uint64_t* out /* host [n][4] */);
/* n independent `two_to_one(left, right)` compressions. */
int qpzk_two_to_one(qpzk_ctx* ctx, const uint64_t* pairs /* host [n][8] */, uint64_t n,
                    uint64_t* out /* host [n][4] */);

/* ---- MerkleTree::new / prove / cap (qp-plonky2 hash/merkle_tree.rs) ----
 * Reached from every commit under /root/reference/wormhole/prover/src/lib.rs:233-237 and the FRI
 * commit phase. Leaves: row-major [nleaves][leaf_len]; nleaves a power of two >= 2^cap_height. */
int qpzk_merkle_new(qpzk_ctx* ctx, const uint64_t* leaves, uint64_t nleaves, uint32_t leaf_len,
                    uint32_t cap_height, qpzk_tree** out);
int qpzk_tree_cap(const qpzk_tree* t, uint64_t* out /* [2^cap_height][4] */);
/* MerkleTree::prove(leaf_index): siblings bottom-up, log2(nleaves) - cap_height of them. */
int qpzk_tree_prove(const qpzk_tree* t, uint64_t leaf_index, uint64_t* siblings /* [.][4] */);
/* The `digests` vector in plonky2's interleaved layout, 2*(nleaves - 2^cap_height) hashes. */
int qpzk_tree_digests(const qpzk_tree* t, uint64_t* out);
void qpzk_tree_free(qpzk_tree* t);

/* ---- PolynomialBatch (qp-plonky2 fri/oracle.rs) ----
 * from_values: `PolynomialBatch::from_values(values, rate_bits, blinding, cap_height, ..)`
 *   = per column IFFT -> LDE on the coset g*<w_N> -> (+ salt columns) -> bit-reversed row-major
 *   leaves -> MerkleTree::new.  Reached from ProverCircuitData::prove
 *   (/root/reference/wormhole/prover/src/lib.rs:233-237, wires and Z/partial-product batches) and
 *   CircuitBuilder::build (/root/reference/wormhole/circuit/src/circuit.rs:98-108, constants|sigmas;
 *   also /root/reference/wormhole/aggregator/src/circuits/tree.rs:127 and /root/reference/voting/src/lib.rs:355).
 * from_coeffs: same from coefficient form (quotient chunks, /root/reference/wormhole/prover/src/lib.rs:233-237).
 * values / coeffs: column-major [ncols][2^degree_bits].
 * salts: NULL (blinding = false), or column-major [salt_cols][2^(degree_bits+rate_bits)] values in
 *   natural LDE-domain order — the caller draws them from its RNG (the reference uses an OS RNG;
 *   injecting them keeps proofs reproducible). */
int qpzk_batch_from_values(qpzk_ctx* ctx, const uint64_t* values, uint32_t ncols, uint32_t degree_bits,
                           uint32_t rate_bits, uint32_t cap_height, const uint64_t* salts,
                           uint32_t salt_cols, qpzk_batch** out);
int qpzk_batch_from_coeffs(qpzk_ctx* ctx, const uint64_t* coeffs, uint32_t ncols, uint32_t degree_bits,
                           uint32_t rate_bits, uint32_t cap_height, const uint64_t* salts,
                           uint32_t salt_cols, qpzk_batch** out);
/* Same, inputs already resident on the context's device (used by the prover pipeline, the
 * multi-GPU shards and the HBM-resident bench leg). The input buffer is not modified. */
int qpzk_batch_from_values_dev(qpzk_ctx* ctx, const uint64_t* values_dev, uint32_t ncols,
                               uint32_t degree_bits, uint32_t rate_bits, uint32_t cap_height,
                               const uint64_t* salts_dev, uint32_t salt_cols, qpzk_batch** out);
int qpzk_batch_from_coeffs_dev(qpzk_ctx* ctx, const uint64_t* coeffs_dev, uint32_t ncols,
                               uint32_t degree_bits, uint32_t rate_bits, uint32_t cap_height,
                               const uint64_t* salts_dev, uint32_t salt_cols, qpzk_batch** out);
/* Shard of a commit for multi-GPU: every rank holds the value columns and computes
 * the coefficients, but only the cap subtrees [subtree_begin, subtree_end) of the 2^cap_height -
 * i.e. a contiguous range of bit-reversed leaves = whole cosets of the LDE domain - are evaluated,
 * hashed and reduced. Cap entries (and digests) of the other subtrees are zero; the caller
 * all-gathers the 2^cap_height x 32-byte subtree roots (NCCL on `qpzk_batch_cap_dev`, or on the host
 * followed by `qpzk_batch_set_cap`). The range must be a multiple of 2^(cap_height - rate_bits)
 * subtrees when cap_height > rate_bits (QPZK_ERR_UNSUPPORTED otherwise). Rows, Merkle paths and
 * `get_lde_values` are served for the owned leaves ONLY: a shard allocates just its own rows, and
 * qpzk_batch_open / _get_lde_rows return QPZK_ERR_BAD_ARG for a leaf of another rank (qpzk_batch_export
 * refuses shards). salts_dev stays [salt_cols][N] in natural order on every rank. */
int qpzk_batch_from_values_shard_dev(qpzk_ctx* ctx, const uint64_t* values_dev, uint32_t ncols,
                                     uint32_t degree_bits, uint32_t rate_bits, uint32_t cap_height,
                                     const uint64_t* salts_dev, uint32_t salt_cols,
                                     uint32_t subtree_begin, uint32_t subtree_end, qpzk_batch** out);
/* The same, ENQUEUED only: the call returns without waiting for the context's stream, the handle is valid for
 * stream-ordered use on the same context at once (an NCCL all-gather on qpzk_batch_cap_dev issued on
 * qpzk_ctx_stream follows the commit without a host round trip); synchronise (qpzk_ctx_sync, or any blocking
 * accessor) before reading results on the host. Per-stage times are not collected. */
int qpzk_batch_from_values_shard_dev_async(qpzk_ctx* ctx, const uint64_t* values_dev, uint32_t ncols,
                                           uint32_t degree_bits, uint32_t rate_bits, uint32_t cap_height,
                                           const uint64_t* salts_dev, uint32_t salt_cols,
                                           uint32_t subtree_begin, uint32_t subtree_end, qpzk_batch** out);
int qpzk_batch_from_coeffs_shard_dev(qpzk_ctx* ctx, const uint64_t* coeffs_dev, uint32_t ncols,
                                     uint32_t degree_bits, uint32_t rate_bits, uint32_t cap_height,
                                     const uint64_t* salts_dev, uint32_t salt_cols,
                                     uint32_t subtree_begin, uint32_t subtree_end, qpzk_batch** out);
/* `merkle_tree.cap`: [2^cap_height][4]. */
int qpzk_batch_cap(const qpzk_batch* b, uint64_t* out);
/* Device pointer to the cap (for NCCL all-gather of subtree roots); 2^cap_height * 4 u64. */
uint64_t* qpzk_batch_cap_dev(qpzk_batch* b);
/* Overwrite the cap with the gathered one (host pointer, [2^cap_height][4]). */
int qpzk_batch_set_cap(qpzk_batch* b, const uint64_t* cap);
/* `polynomials`: coefficient form, column-major [ncols][n]. */
int qpzk_batch_coeffs(const qpzk_batch* b, uint64_t* out);
/* `get_lde_values(index, step)` for a list of indices: out[i] = leaves[rev(idx[i]*step)][..ncols]
 * (salt columns stripped), row-major [nidx][ncols]. */
int qpzk_batch_get_lde_rows(const qpzk_batch* b, const uint32_t* idx, uint32_t nidx, uint32_t step,
                            uint64_t* out);
/* `merkle_tree.leaves[leaf_index]` (salted row, ncols + salt_cols wide) and
 * `merkle_tree.prove(leaf_index)`. Either output may be NULL. */
int qpzk_batch_open(const qpzk_batch* b, uint64_t leaf_index, uint64_t* leaf_out, uint64_t* siblings_out);
/* Full materialisation (compatibility / debugging): `merkle_tree.leaves` row-major
 * [N][ncols+salt_cols] and `merkle_tree.digests` in plonky2's layout. Either may be NULL. */
int qpzk_batch_export(const qpzk_batch* b, uint64_t* leaves, uint64_t* digests);
/* PolynomialBatch <-> bytes in the layout of plonky2's `Write::write_polynomial_batch` /
 * `Read::read_polynomial_batch` (qp-plonky2 util/serialization): polynomials (count, then length + coefficients
 * each) | merkle_tree (leaf count, then length + elements each; digest count + digests in plonky2's layout; cap
 * height + cap) | degree_log | rate_bits | blinding (1 byte). This is the `constants_sigmas_commitment` field of
 * the serialized ProverOnlyCircuitData that `WormholeProver::new_from_files`
 * (/root/reference/wormhole/prover/src/lib.rs:105-187) reads, so a prover restored from files puts the
 * commitment on the device WITHOUT recomputing it. from_bytes checks every length against the
 * others (QPZK_ERR_BAD_ARG on a mismatch or a truncated buffer) and copies; with QPZK_IMPORT_VERIFY it also
 * recomputes the LDE and the tree on the device and refuses bytes that are not the commitment of the
 * polynomials they carry. `consumed` (may be NULL) receives the number of bytes read. */
#define QPZK_IMPORT_VERIFY 1u
int qpzk_batch_serialized_size(const qpzk_batch* b, uint64_t* nbytes);
int qpzk_batch_to_bytes(const qpzk_batch* b, uint8_t* out, uint64_t capacity);
int qpzk_batch_from_bytes(qpzk_ctx* ctx, const uint8_t* bytes, uint64_t nbytes, uint32_t flags, qpzk_batch** out,
                          uint64_t* consumed);
/* `OpeningSet::new` for one oracle (qp-plonky2 plonk/proof.rs, reached from prove() at
 * /root/reference/wormhole/prover/src/lib.rs:233-237): every committed polynomial evaluated at an
 * extension-field point; out = [ncols][2]. */
int qpzk_batch_eval_ext(const qpzk_batch* b, const uint64_t* point /* [2] */, uint64_t* out);
uint32_t qpzk_batch_ncols(const qpzk_batch* b);
uint32_t qpzk_batch_width(const qpzk_batch* b); /* ncols + salt_cols */
uint32_t qpzk_batch_degree_bits(const qpzk_batch* b);
void qpzk_batch_free(qpzk_batch* b);

/* ---- circuit + prove() (qp-plonky2 plonk/prover.rs `prove`, plonk/circuit_builder.rs `build`) ----
 * qpzk_circuit_create is the prover-side residue of `CircuitBuilder::build()`
 * (/root/reference/wormhole/circuit/src/circuit.rs:98-108): it takes the CommonCircuitData bytes
 * (`common.to_bytes()`, the layout of /root/reference/wormhole/bench-data/common.bin), the circuit
 * digest and the constants|sigmas value columns ([num_constants + num_routed_wires][n], values on
 * the subgroup), performs the constants|sigmas commit ONCE and keeps everything device-resident for
 * every later proof of the same circuit. Supported gate set: Noop, Constant, PublicInput,
 * BaseSum<2>, Arithmetic, Poseidon (the wormhole and voting circuits) and the recursion set
 * ArithmeticExtension, MulExtension, PoseidonMds, RandomAccess, Reducing, ReducingExtension,
 * Exponentiation, CosetInterpolation (aggregation nodes); anything else returns QPZK_ERR_UNSUPPORTED
 * at creation. The common data is untrusted input: every count, index and size in it is range-checked
 * before anything is launched, and constants_sigmas_words must equal
 * (num_constants + num_routed_wires) << degree_bits (QPZK_ERR_BAD_ARG otherwise). */
typedef struct qpzk_circuit qpzk_circuit;
int qpzk_circuit_create(qpzk_ctx* ctx, const uint8_t* common_bytes, size_t common_len,
                        const uint64_t* circuit_digest /* [4] */, const uint64_t* constants_sigmas,
                        size_t constants_sigmas_words, qpzk_circuit** out);
/* The same circuit from a prover restored from files (`WormholeProver::new_from_files`,
 * /root/reference/wormhole/prover/src/lib.rs:105-187): instead of the value columns the call takes
 * the serialized `constants_sigmas_commitment` (qpzk_batch_from_bytes layout; flags as there) and puts it on the
 * device as it is - no transform of 84 columns x 8 cosets, no hashing. Only the value columns the permutation
 * argument reads are rebuilt, by one forward transform of the stored coefficients. The commitment's shape must
 * be the one the common data states. qpzk_circuit_commitment_size / _to_bytes write that field for a circuit
 * created either way (what a patched `ProverOnlyCircuitData::to_bytes` stores). */
int qpzk_circuit_create_from_commitment(qpzk_ctx* ctx, const uint8_t* common_bytes, size_t common_len,
                                        const uint64_t* circuit_digest /* [4] */, const uint8_t* commitment_bytes,
                                        uint64_t commitment_len, uint32_t flags, qpzk_circuit** out);
int qpzk_circuit_commitment_size(const qpzk_circuit* c, uint64_t* nbytes);
int qpzk_circuit_commitment_to_bytes(const qpzk_circuit* c, uint8_t* out, uint64_t capacity);
/* `verifier_only.constants_sigmas_cap`: [2^cap_height][4]; cap_words = capacity of `out` in u64. */
int qpzk_circuit_cap(const qpzk_circuit* c, uint64_t* out, size_t cap_words);
/* Shape of the circuit as parsed from the common data: out[0..8) = degree_bits, rate_bits, cap_height,
 * num_wires, num_routed_wires, num_challenges, salt columns per blinded oracle (0 or 4), num_public_inputs. */
int qpzk_circuit_info(const qpzk_circuit* c, uint32_t* out /* [8] */);
/* `VerifierOnlyCircuitData::to_bytes()`; returns the length (writes only if it fits in cap). */
size_t qpzk_circuit_verifier_only(const qpzk_circuit* c, uint8_t* out, size_t cap);
void qpzk_circuit_free(qpzk_circuit* c);
/* `ProverCircuitData::prove` after witness generation
 * (/root/reference/wormhole/prover/src/lib.rs:233-237, /root/reference/wormhole/aggregator/src/circuits/tree.rs:136,
 * /root/reference/voting/src/lib.rs:356): wires = the full witness matrix, column-major
 * [num_wires][n]. Runs: commit wires -> Z / partial products -> commit -> quotient -> commit ->
 * openings -> FRI (combine, fold + commit, proof of work, queries) and writes
 * `ProofWithPublicInputs::to_bytes()`. salts_*: NULL for non-hiding circuits, else host
 * [4][n << rate_bits] per blinded oracle. The proof-of-work witness is the
 * smallest valid one (the reference returns whichever a rayon worker finds first; any valid witness
 * verifies). flags bit 0: keep intermediates for qpzk_prove_trace; bit 1: `wires` and the salt
 * pointers are DEVICE pointers on the context's device (HBM-resident witness).
 * bit 2 (QPZK_PROVE_SEEDED_SALTS): the blinding salts are DRAWN ON THE DEVICE: salts_wires points to a
 * 32-byte seed on the host (salt_words = 4), salts_zs / salts_quotient are ignored, and the salts of oracle o
 * (0 wires, 1 Z|partial products, 2 quotient) are the ChaCha8 stream keyed by the seed with nonce o (eight
 * salts per 64-byte block, element j = column j / N, natural row j % N, values >= p reduced by p) - what
 * plonky2 takes from the OS RNG, reproducible, and 12.6 MB less host-to-device traffic per wormhole proof.
 * wires_words / salt_words: element counts of `wires` and of EACH salt array, checked against the circuit
 * (num_wires << degree_bits, 4 << (degree_bits + rate_bits)).
 * The whole proof is enqueued on the context's stream without a host round trip (the transcript is a
 * device kernel, the challenges never leave HBM); the call waits once, at the end. */
int qpzk_prove(qpzk_circuit* c, const uint64_t* wires, size_t wires_words, const uint64_t* public_inputs,
               uint32_t num_public_inputs, const uint64_t* salts_wires, const uint64_t* salts_zs,
               const uint64_t* salts_quotient, size_t salt_words, uint32_t flags, uint8_t* proof_out,
               size_t proof_cap, size_t* proof_len);
/* The same in two halves, so that ONE host thread can keep several contexts busy (the aggregator's
 * concurrent chunk proofs, /root/reference/wormhole/aggregator/src/circuits/tree.rs:92-103, without a
 * thread per proof): qpzk_prove_begin enqueues the proof and returns; qpzk_prove_end waits for it and
 * writes the bytes. One proof in flight per circuit handle; host input buffers must stay valid (and, for
 * an asynchronous upload, pinned) until qpzk_prove_end returns. */
int qpzk_prove_begin(qpzk_circuit* c, const uint64_t* wires, size_t wires_words, const uint64_t* public_inputs,
                     uint32_t num_public_inputs, const uint64_t* salts_wires, const uint64_t* salts_zs,
                     const uint64_t* salts_quotient, size_t salt_words, uint32_t flags);
int qpzk_prove_end(qpzk_circuit* c, uint8_t* proof_out, size_t proof_cap, size_t* proof_len);
/* ONE proof over several GPUs (BASELINE configs[4]: an aggregation proof of 2^17-2^18 rows on 8 B200;
 * SURVEY.md 8(e)). Every rank holds the circuit and the whole witness and calls the same sequence on its own
 * context with its own range of cap subtrees - whole LDE cosets, as for qpzk_batch_from_values_shard_dev:
 *
 *   qpzk_sprove_begin(.., subtree_begin, subtree_end, &s)        commit the wires shard
 *   loop:  for i = 0, 1, ..: qpzk_sprove_exchange(s, i, ..)       what to exchange now (kind 0: nothing more)
 *          qpzk_sprove_next(s)                                    next phase
 *   until qpzk_sprove_phase(s) == 6;  qpzk_sprove_end(s, ..)      wait once, ProofWithPublicInputs bytes
 *
 * Phases and what follows them (all stream-ordered on the context's stream, no host synchronisation):
 *   1 wires commit, 2 Z/partial products + commit, 4 quotient commit: QPZK_EXCHANGE_ALLGATHER of the
 *     2^cap_height subtree roots, in place on the device cap (each rank owns words [own_begin, own_end));
 *   3 quotient values on the rank's own points: all-gather of one contiguous block per challenge;
 *   5 openings, FRI, proof of work, queries (replicated; rows of the sharded oracles are served by their
 *     owner, zeros elsewhere): QPZK_EXCHANGE_SUM over the ranks (u64 wrap-around add);
 *   6 the proof pieces are on their way to the host.
 * The Fiat-Shamir transcript runs replicated on every device, so every rank ends with the same bytes, equal
 * to the single-GPU proof. With the full range [0, 2^cap_height) nothing is exchanged and the sequence is
 * qpzk_prove in steps. Needs quotient_degree_factor == 2^rate_bits (every standard configuration). */
typedef struct qpzk_sprove qpzk_sprove;
#define QPZK_EXCHANGE_ALLGATHER 1u
#define QPZK_EXCHANGE_SUM 2u
int qpzk_sprove_begin(qpzk_circuit* c, const uint64_t* wires, size_t wires_words, const uint64_t* public_inputs,
                      uint32_t num_public_inputs, const uint64_t* salts_wires, const uint64_t* salts_zs,
                      const uint64_t* salts_quotient, size_t salt_words, uint32_t flags, uint32_t subtree_begin,
                      uint32_t subtree_end, qpzk_sprove** out);
int qpzk_sprove_next(qpzk_sprove* s);
uint32_t qpzk_sprove_phase(const qpzk_sprove* s);
int qpzk_sprove_exchange(const qpzk_sprove* s, uint32_t index, uint64_t** dev_ptr, uint64_t* words,
                         uint64_t* own_begin, uint64_t* own_end, uint32_t* kind);
/* Waits for the proof, writes the bytes and releases s (also when it fails). */
int qpzk_sprove_end(qpzk_sprove* s, uint8_t* proof_out, size_t proof_cap, size_t* proof_len);

/* Per-stage hooks for a fork that keeps plonky2's own `prove()` loop and transcript (integration depth (b),
 * INTEGRATION.md): the same device code qpzk_prove runs.
 * qpzk_zs_partial_products = `all_wires_permutation_partial_products` + the running product (qp-plonky2
 *   plonk/prover.rs): wires host [num_wires][n]; out [nch*(1+npp)][n] = Z_0..Z_{nch-1} followed by each
 *   challenge's num_partial_products columns (the layout of the second committed oracle).
 * qpzk_quotient = `compute_quotient_polys` (plonk/prover.rs + vanishing_poly.rs + gates/*): from the wires
 *   and Z|partial-product batches committed on the circuit's context; out_chunks [nch*qdf][n] quotient
 *   chunk coefficients, ready for `from_coeffs`. pi_hash = hash_no_pad(public_inputs). */
int qpzk_zs_partial_products(qpzk_circuit* c, const uint64_t* wires, const uint64_t* betas, const uint64_t* gammas,
                             uint64_t* out);
int qpzk_quotient(qpzk_circuit* c, const qpzk_batch* wires_batch, const qpzk_batch* zs_batch,
                  const uint64_t* pi_hash /* [4] */, const uint64_t* betas, const uint64_t* gammas,
                  const uint64_t* alphas, uint64_t* out_chunks);
/* The FRI prover as a resumable object (`PolynomialBatch::prove_openings`, `fri_committed_trees`,
 * `fri_prover_query_rounds`; qp-plonky2 fri/oracle.rs, fri/prover.rs); the transcript stays with the caller:
 *   qpzk_fri_begin(circuit, wires, zs|pp, quotient batches, zeta, alpha)      batch-combine at zeta / g*zeta
 *   repeat qpzk_fri_num_rounds times:
 *     qpzk_fri_commit_round -> cap [2^cap_height][4]   (caller observes it and squeezes beta)
 *     qpzk_fri_fold(beta)
 *   qpzk_fri_final_poly -> [len][2] extension coefficients   (caller observes, grinds with qpzk_fri_pow)
 *   qpzk_fri_query(x_index) -> for each of the 4 oracles (constants|sigmas, wires, zs|pp, quotient): the
 *     salted row then its Merkle path ([L][4]); then for each reduction round: 2^arity_bits extension
 *     evaluations then the path. Lengths follow from the circuit; *len_words reports the total. */
typedef struct qpzk_fri qpzk_fri;
int qpzk_fri_begin(qpzk_circuit* c, const qpzk_batch* wires_batch, const qpzk_batch* zs_batch,
                   const qpzk_batch* quotient_batch, const uint64_t* zeta /* [2] */, const uint64_t* alpha /* [2] */,
                   qpzk_fri** out);
uint32_t qpzk_fri_num_rounds(const qpzk_fri* f);
int qpzk_fri_commit_round(qpzk_fri* f, uint64_t* cap_out);
int qpzk_fri_fold(qpzk_fri* f, const uint64_t* beta /* [2] */);
int qpzk_fri_final_poly(qpzk_fri* f, uint64_t* out, size_t cap_words, size_t* len_words);
int qpzk_fri_query(qpzk_fri* f, uint64_t x_index, uint64_t* out, size_t cap_words, size_t* len_words);
void qpzk_fri_free(qpzk_fri* f);
/* Parity hook (flags bit 0): which = 0 challenges, 1 zs|partial-product values [.][n],
 * 2 quotient chunk coefficients [.][n], 3 FRI input polynomial [n][2]. Returns u64 count. */
size_t qpzk_prove_trace(const qpzk_circuit* c, int which, uint64_t* out);
/* Device time of the stages of the last proof, ms: [0] wires commit, [1] Z/partial products +
 * commit, [2] quotient + commit, [3] openings, [4] FRI combine, [5] FRI commit phase,
 * [6] proof of work, [7] queries. */
int qpzk_prove_stage_ms(const qpzk_circuit* c, float* out16);

/* `fri_proof_of_work` (qp-plonky2 fri/prover.rs): the SMALLEST witness w such that permuting the
 * duplex-sponge state with w written at `input_pos` (the challenger's input buffer length) yields an
 * output word 7 with at least `min_leading_zeros` leading zero bits. The reference returns whichever
 * witness a rayon worker finds first; any valid witness verifies. */
int qpzk_fri_pow(qpzk_ctx* ctx, const uint64_t* sponge_state /* [12] */, uint32_t input_pos,
                 uint32_t min_leading_zeros, uint64_t* witness_out);

/* ---- measurement helper: dependency-free integer multiply-add throughput (the Poseidon
 * roofline denominator). kind 0: 32-bit mad.lo.u32, kind 1: mad.wide.u32.
 * Returns multiply-adds per second over the whole device. */
int qpzk_measure_imad_peak(qpzk_ctx* ctx, int kind, double* out_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* QPZK_H */
