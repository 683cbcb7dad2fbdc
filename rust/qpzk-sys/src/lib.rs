//! One declaration per symbol of `include/qpzk.h` (generated from the header by the snippet in
//! INTEGRATION.md; `tests/test_abi.py` fails if the two drift apart). Compile-unverified here: the build
//! image has no Rust toolchain.
//!
//! Conventions (see the header): field elements are `u64` (any representative in, canonical out),
//! extension elements two consecutive `u64`, hashes four. Host pointers are borrowed for the call.
//! Every function returns `QPZK_OK` (0) or a negative status; `qpzk_last_error` has the message.
#![allow(non_camel_case_types)]
#![no_std]
use core::ffi::{c_char, c_int, c_void};

pub const QPZK_OK: c_int = 0;
pub const QPZK_ERR_BAD_ARG: c_int = -1;
pub const QPZK_ERR_CUDA: c_int = -2;
pub const QPZK_ERR_OOM: c_int = -3;
pub const QPZK_ERR_NOT_DIVISIBLE: c_int = -4;
pub const QPZK_ERR_UNSUPPORTED: c_int = -5;
pub const QPZK_SALT_SIZE: usize = 4;
pub const QPZK_CTX_BLOCKING_SYNC: u32 = 1;
pub const QPZK_CTX_YIELD_SYNC: u32 = 2;
pub const QPZK_EXCHANGE_ALLGATHER: u32 = 1;
pub const QPZK_EXCHANGE_SUM: u32 = 2;
pub const QPZK_IMPORT_VERIFY: u32 = 1;
pub const QPZK_PROVE_TRACE: u32 = 1;
pub const QPZK_PROVE_DEVICE_INPUTS: u32 = 2;
pub const QPZK_PROVE_SEEDED_SALTS: u32 = 4;

#[repr(C)] pub struct qpzk_ctx { _p: [u8; 0] }
#[repr(C)] pub struct qpzk_batch { _p: [u8; 0] }
#[repr(C)] pub struct qpzk_tree { _p: [u8; 0] }
#[repr(C)] pub struct qpzk_circuit { _p: [u8; 0] }
#[repr(C)] pub struct qpzk_fri { _p: [u8; 0] }
#[repr(C)] pub struct qpzk_sprove { _p: [u8; 0] }

extern "C" {
    pub fn qpzk_ctx_create(device: c_int, flags: u32, out_: *mut *mut qpzk_ctx) -> c_int;
    pub fn qpzk_ctx_destroy(ctx: *mut qpzk_ctx);
    pub fn qpzk_last_error() -> *const c_char;
    pub fn qpzk_ctx_sync(ctx: *mut qpzk_ctx) -> c_int;
    pub fn qpzk_ctx_stream(ctx: *mut qpzk_ctx) -> *mut c_void;
    pub fn qpzk_ctx_stage_ms(ctx: *mut qpzk_ctx, out_ms: *mut f32) -> c_int;
    pub fn qpzk_ctx_launch_count(ctx: *const qpzk_ctx) -> u64;
    pub fn qpzk_host_alloc(bytes: usize, out_: *mut *mut c_void) -> c_int;
    pub fn qpzk_host_free(p: *mut c_void);
    pub fn qpzk_dev_alloc(ctx: *mut qpzk_ctx, bytes: usize, out_: *mut *mut c_void) -> c_int;
    pub fn qpzk_dev_free(ctx: *mut qpzk_ctx, p: *mut c_void);
    pub fn qpzk_memcpy_h2d(ctx: *mut qpzk_ctx, dst_dev: *mut c_void, src_host: *const c_void, bytes: usize) -> c_int;
    pub fn qpzk_memcpy_d2h(ctx: *mut qpzk_ctx, dst_host: *mut c_void, src_dev: *const c_void, bytes: usize) -> c_int;
    pub fn qpzk_poseidon_permute(ctx: *mut qpzk_ctx, states: *mut u64, n: u64) -> c_int;
    pub fn qpzk_hash_no_pad(ctx: *mut qpzk_ctx, inputs: *const u64, n: u64, len: u32, out_: *mut u64) -> c_int;
    pub fn qpzk_two_to_one(ctx: *mut qpzk_ctx, pairs: *const u64, n: u64, out_: *mut u64) -> c_int;
    pub fn qpzk_merkle_new(ctx: *mut qpzk_ctx, leaves: *const u64, nleaves: u64, leaf_len: u32, cap_height: u32, out_: *mut *mut qpzk_tree) -> c_int;
    pub fn qpzk_tree_cap(t: *const qpzk_tree, out_: *mut u64) -> c_int;
    pub fn qpzk_tree_prove(t: *const qpzk_tree, leaf_index: u64, siblings: *mut u64) -> c_int;
    pub fn qpzk_tree_digests(t: *const qpzk_tree, out_: *mut u64) -> c_int;
    pub fn qpzk_tree_free(t: *mut qpzk_tree);
    pub fn qpzk_batch_from_values(ctx: *mut qpzk_ctx, values: *const u64, ncols: u32, degree_bits: u32, rate_bits: u32, cap_height: u32, salts: *const u64, salt_cols: u32, out_: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_batch_from_coeffs(ctx: *mut qpzk_ctx, coeffs: *const u64, ncols: u32, degree_bits: u32, rate_bits: u32, cap_height: u32, salts: *const u64, salt_cols: u32, out_: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_batch_from_values_dev(ctx: *mut qpzk_ctx, values_dev: *const u64, ncols: u32, degree_bits: u32, rate_bits: u32, cap_height: u32, salts_dev: *const u64, salt_cols: u32, out_: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_batch_from_coeffs_dev(ctx: *mut qpzk_ctx, coeffs_dev: *const u64, ncols: u32, degree_bits: u32, rate_bits: u32, cap_height: u32, salts_dev: *const u64, salt_cols: u32, out_: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_batch_from_values_shard_dev(ctx: *mut qpzk_ctx, values_dev: *const u64, ncols: u32, degree_bits: u32, rate_bits: u32, cap_height: u32, salts_dev: *const u64, salt_cols: u32, subtree_begin: u32, subtree_end: u32, out_: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_batch_from_values_shard_dev_async(ctx: *mut qpzk_ctx, values_dev: *const u64, ncols: u32, degree_bits: u32, rate_bits: u32, cap_height: u32, salts_dev: *const u64, salt_cols: u32, subtree_begin: u32, subtree_end: u32, out_: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_batch_from_coeffs_shard_dev(ctx: *mut qpzk_ctx, coeffs_dev: *const u64, ncols: u32, degree_bits: u32, rate_bits: u32, cap_height: u32, salts_dev: *const u64, salt_cols: u32, subtree_begin: u32, subtree_end: u32, out_: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_batch_cap(b: *const qpzk_batch, out_: *mut u64) -> c_int;
    pub fn qpzk_batch_cap_dev(b: *mut qpzk_batch) -> *mut u64;
    pub fn qpzk_batch_set_cap(b: *mut qpzk_batch, cap: *const u64) -> c_int;
    pub fn qpzk_batch_coeffs(b: *const qpzk_batch, out_: *mut u64) -> c_int;
    pub fn qpzk_batch_get_lde_rows(b: *const qpzk_batch, idx: *const u32, nidx: u32, step: u32, out_: *mut u64) -> c_int;
    pub fn qpzk_batch_open(b: *const qpzk_batch, leaf_index: u64, leaf_out: *mut u64, siblings_out: *mut u64) -> c_int;
    pub fn qpzk_batch_export(b: *const qpzk_batch, leaves: *mut u64, digests: *mut u64) -> c_int;
    pub fn qpzk_batch_serialized_size(b: *const qpzk_batch, nbytes: *mut u64) -> c_int;
    pub fn qpzk_batch_to_bytes(b: *const qpzk_batch, out_: *mut u8, capacity: u64) -> c_int;
    pub fn qpzk_batch_from_bytes(ctx: *mut qpzk_ctx, bytes: *const u8, nbytes: u64, flags: u32, out_: *mut *mut qpzk_batch, consumed: *mut u64) -> c_int;
    pub fn qpzk_batch_eval_ext(b: *const qpzk_batch, point: *const u64, out_: *mut u64) -> c_int;
    pub fn qpzk_batch_ncols(b: *const qpzk_batch) -> u32;
    pub fn qpzk_batch_width(b: *const qpzk_batch) -> u32;
    pub fn qpzk_batch_degree_bits(b: *const qpzk_batch) -> u32;
    pub fn qpzk_batch_free(b: *mut qpzk_batch);
    pub fn qpzk_circuit_create(ctx: *mut qpzk_ctx, common_bytes: *const u8, common_len: usize, circuit_digest: *const u64, constants_sigmas: *const u64, constants_sigmas_words: usize, out_: *mut *mut qpzk_circuit) -> c_int;
    pub fn qpzk_circuit_create_from_commitment(ctx: *mut qpzk_ctx, common_bytes: *const u8, common_len: usize, circuit_digest: *const u64, commitment_bytes: *const u8, commitment_len: u64, flags: u32, out_: *mut *mut qpzk_circuit) -> c_int;
    pub fn qpzk_circuit_commitment_size(c: *const qpzk_circuit, nbytes: *mut u64) -> c_int;
    pub fn qpzk_circuit_commitment_to_bytes(c: *const qpzk_circuit, out_: *mut u8, capacity: u64) -> c_int;
    pub fn qpzk_circuit_cap(c: *const qpzk_circuit, out_: *mut u64, cap_words: usize) -> c_int;
    pub fn qpzk_circuit_info(c: *const qpzk_circuit, out_: *mut u32) -> c_int;
    pub fn qpzk_circuit_verifier_only(c: *const qpzk_circuit, out_: *mut u8, cap: usize) -> usize;
    pub fn qpzk_circuit_free(c: *mut qpzk_circuit);
    pub fn qpzk_prove(c: *mut qpzk_circuit, wires: *const u64, wires_words: usize, public_inputs: *const u64, num_public_inputs: u32, salts_wires: *const u64, salts_zs: *const u64, salts_quotient: *const u64, salt_words: usize, flags: u32, proof_out: *mut u8, proof_cap: usize, proof_len: *mut usize) -> c_int;
    pub fn qpzk_prove_begin(c: *mut qpzk_circuit, wires: *const u64, wires_words: usize, public_inputs: *const u64, num_public_inputs: u32, salts_wires: *const u64, salts_zs: *const u64, salts_quotient: *const u64, salt_words: usize, flags: u32) -> c_int;
    pub fn qpzk_prove_end(c: *mut qpzk_circuit, proof_out: *mut u8, proof_cap: usize, proof_len: *mut usize) -> c_int;
    pub fn qpzk_sprove_begin(c: *mut qpzk_circuit, wires: *const u64, wires_words: usize, public_inputs: *const u64, num_public_inputs: u32, salts_wires: *const u64, salts_zs: *const u64, salts_quotient: *const u64, salt_words: usize, flags: u32, subtree_begin: u32, subtree_end: u32, out_: *mut *mut qpzk_sprove) -> c_int;
    pub fn qpzk_sprove_next(s: *mut qpzk_sprove) -> c_int;
    pub fn qpzk_sprove_phase(s: *const qpzk_sprove) -> u32;
    pub fn qpzk_sprove_exchange(s: *const qpzk_sprove, index: u32, dev_ptr: *mut *mut u64, words: *mut u64, own_begin: *mut u64, own_end: *mut u64, kind: *mut u32) -> c_int;
    pub fn qpzk_sprove_end(s: *mut qpzk_sprove, proof_out: *mut u8, proof_cap: usize, proof_len: *mut usize) -> c_int;
    pub fn qpzk_zs_partial_products(c: *mut qpzk_circuit, wires: *const u64, betas: *const u64, gammas: *const u64, out_: *mut u64) -> c_int;
    pub fn qpzk_quotient(c: *mut qpzk_circuit, wires_batch: *const qpzk_batch, zs_batch: *const qpzk_batch, pi_hash: *const u64, betas: *const u64, gammas: *const u64, alphas: *const u64, out_chunks: *mut u64) -> c_int;
    pub fn qpzk_fri_begin(c: *mut qpzk_circuit, wires_batch: *const qpzk_batch, zs_batch: *const qpzk_batch, quotient_batch: *const qpzk_batch, zeta: *const u64, alpha: *const u64, out_: *mut *mut qpzk_fri) -> c_int;
    pub fn qpzk_fri_num_rounds(f: *const qpzk_fri) -> u32;
    pub fn qpzk_fri_commit_round(f: *mut qpzk_fri, cap_out: *mut u64) -> c_int;
    pub fn qpzk_fri_fold(f: *mut qpzk_fri, beta: *const u64) -> c_int;
    pub fn qpzk_fri_final_poly(f: *mut qpzk_fri, out_: *mut u64, cap_words: usize, len_words: *mut usize) -> c_int;
    pub fn qpzk_fri_query(f: *mut qpzk_fri, x_index: u64, out_: *mut u64, cap_words: usize, len_words: *mut usize) -> c_int;
    pub fn qpzk_fri_free(f: *mut qpzk_fri);
    pub fn qpzk_prove_trace(c: *const qpzk_circuit, which: c_int, out_: *mut u64) -> usize;
    pub fn qpzk_prove_stage_ms(c: *const qpzk_circuit, out16: *mut f32) -> c_int;
    pub fn qpzk_fri_pow(ctx: *mut qpzk_ctx, sponge_state: *const u64, input_pos: u32, min_leading_zeros: u32, witness_out: *mut u64) -> c_int;
    pub fn qpzk_measure_imad_peak(ctx: *mut qpzk_ctx, kind: c_int, out_ops_per_s: *mut f64) -> c_int;
}
