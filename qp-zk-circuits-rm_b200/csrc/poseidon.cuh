// Poseidon-Goldilocks permutation (width 12, rate 8, x^7, 4 + 22 + 4 rounds) for sm_100a:
// one permutation per thread, the 12-word state held in registers.
//
// Replaces `PoseidonPermutation` / `PoseidonHash::{hash_no_pad, hash_or_noop, two_to_one}` of
// qp-plonky2 1.1.1 (un-vendored; reference call sites
// /root/reference/wormhole/circuit/src/nullifier.rs:64-65,
// /root/reference/wormhole/circuit/src/unspendable_account.rs:54-56 and every `MerkleTree::new`
// under /root/reference/wormhole/prover/src/lib.rs:233-237).
//
// Integer-pipe bound: ~1.1k 64-bit modular multiplies per permutation.
//  * full rounds: s-box = 4 modmuls per lane; the circulant MDS has entries < 2^6, so it runs on
//    the 32-bit halves of the state with mad.wide.u32 accumulation and ONE fold per output lane.
//  * partial rounds use the sparse factorisation (tables derived in poseidon_tables.hpp): one
//    s-box, one 12-term dot product accumulated in 160 bits with a single fold, 11 fused
//    multiply-adds.
#pragma once
#include "gl.cuh"

namespace qpzk {

__constant__ u64 c_rc[360];
__constant__ u64 c_fast_first[12];
__constant__ u64 c_fast_rc[22];
__constant__ u64 c_fast_init[121];
__constant__ u64 c_fast_w_hat[242];
__constant__ u64 c_fast_v[242];

GL_DEV u64 sbox7(u64 x) {
  u64 x2 = gl_sqr(x);
  u64 x4 = gl_sqr(x2);
  u64 x3 = gl_mul(x, x2);
  return gl_mul(x3, x4);
}

// lo + 2^64*hi with hi < 2^32 (a 96-bit value): lo + hi*EPS, single correction.
GL_DEV u64 gl_reduce96(u64 lo, u32 hi) {
  u64 t1 = (u64)hi * (u64)0xFFFFFFFFu;
  u64 r = lo + t1;
  if (r < t1) r += GL_EPS;
  return r;
}

// 160-bit accumulator for sums of up to 2^32 128-bit products.
struct Acc160 {
  u64 lo, hi;
  u32 top;
};
GL_DEV void acc_init(Acc160& a) { a.lo = 0; a.hi = 0; a.top = 0; }
GL_DEV void acc_mac(Acc160& a, u64 x, u64 y) {
  u64 plo = x * y, phi = __umul64hi(x, y);
  asm("add.cc.u64 %0, %0, %3;\n\t"
      "addc.cc.u64 %1, %1, %4;\n\t"
      "addc.u32 %2, %2, 0;"
      : "+l"(a.lo), "+l"(a.hi), "+r"(a.top)
      : "l"(plo), "l"(phi));
}
// value = lo + 2^64*hi + 2^128*top, and 2^128 == -2^32 (mod p)
GL_DEV u64 acc_reduce(const Acc160& a) {
  u64 r = gl_reduce128(a.lo, a.hi);
  return gl_sub(r, (u64)a.top << 32);
}

GL_DEV void mds_layer(u64 (&s)[12]) {
  const u32 circ[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  u32 lo[12], hi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    lo[i] = (u32)s[i];
    hi[i] = (u32)(s[i] >> 32);
  }
#pragma unroll
  for (int r = 0; r < 12; r++) {
    u64 al = 0, ah = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      al += (u64)lo[(i + r) % 12] * circ[i];
      ah += (u64)hi[(i + r) % 12] * circ[i];
    }
    if (r == 0) {
      al += (u64)lo[0] * 8u;
      ah += (u64)hi[0] * 8u;
    }
    // al, ah < 2^41.  value = al + ah*2^32
    u64 l = al + (ah << 32);
    u32 h = (u32)(ah >> 32) + (u32)(l < al);
    s[r] = gl_reduce96(l, h);
  }
}

// Full round: add constants, x^7 on every lane, MDS.
GL_DEV void full_round(u64 (&s)[12], const u64* __restrict__ rc) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = sbox7(gl_add_c(s[i], rc[i]));
  mds_layer(s);
}

GL_DEV void partial_rounds(u64 (&s)[12]) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_add_c(s[i], c_fast_first[i]);
  {  // mds_partial_layer_init: out[c] = sum_r in[r] * init[r-1][c-1]
    u64 o[12];
    o[0] = s[0];
#pragma unroll
    for (int c = 1; c < 12; c++) {
      Acc160 a;
      acc_init(a);
#pragma unroll
      for (int r = 1; r < 12; r++) acc_mac(a, s[r], c_fast_init[(r - 1) * 11 + (c - 1)]);
      o[c] = acc_reduce(a);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = o[i];
  }
#pragma unroll 1
  for (int r = 0; r < 22; r++) {
    u64 s0 = sbox7(s[0]);
    s0 = gl_add_c(s0, c_fast_rc[r]);  // entry 21 is zero
    Acc160 a;
    acc_init(a);
    acc_mac(a, s0, 25);  // MDS[0][0] = 17 + 8
#pragma unroll
    for (int i = 1; i < 12; i++) acc_mac(a, s[i], c_fast_w_hat[r * 11 + i - 1]);
#pragma unroll
    for (int i = 1; i < 12; i++) s[i] = gl_mad(s0, c_fast_v[r * 11 + i - 1], s[i]);
    s[0] = acc_reduce(a);
  }
}

// In-place permutation; inputs may be any u64 representatives, outputs likewise (not canonical).
GL_DEV void poseidon_permute(u64 (&s)[12]) {
#pragma unroll 1
  for (int r = 0; r < 4; r++) full_round(s, c_rc + 12 * r);
  partial_rounds(s);
#pragma unroll 1
  for (int r = 0; r < 4; r++) full_round(s, c_rc + 12 * (26 + r));
}

}  // namespace qpzk
