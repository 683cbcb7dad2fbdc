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
// The sparse tables of the permutation kernels: the same derivation for the partial rounds that remain sparse
// when the first PV_DENSE_PARTIAL of them run in the textbook form (equal to c_fast_* when that is 0; the
// quotient kernel's PoseidonGate always evaluates the all-sparse form, whose s-box inputs are its wires).
__constant__ u64 c_h_rc[22];
__constant__ u64 c_h_init[121];
__constant__ u64 c_h_w_hat[242];
__constant__ u64 c_h_v[242];
// Kept in constant memory (not literals) on purpose: with literal multipliers nvcc strength-reduces
// every c*x into shift/add chains on the already saturated ALU pipe; as constant-bank operands they
// stay single IMAD.WIDE instructions.
__constant__ u32 c_mds_circ[12];
__constant__ u32 c_mds_diag0;

// PV_SBOX_ALU / PV_COMBINE_ALU / PV_ADDC_ALU (experiments, off): the s-box products, the MDS recombination
// and the round-constant additions spelled so that every addition stays on the ALU pipe (gl.cuh,
// gl_mul_wide_alu). The SASS model said this should win - the FMA-heavy pipe is the busiest unit (76 %,
// profiles/r1_leaf_hash_v5_twoplane_ncu_full.txt) and the spelling takes 17-20 % of its cycles away for 1-2 %
// more instructions - but the permutation got SLOWER: 1106.9 (off) / 1065.0 (s-box) / 1050.3 (+ combine) /
// 1049.6 (+ add) M perm/s, and the dependent 16-lane permutation 9.11 -> 9.28 us. The ALU pipe issues one
// warp instruction every two cycles and is the one that matters once the two are this close.
#ifndef PV_SBOX_ALU
#define PV_SBOX_ALU 0
#endif
#ifndef PV_SQR_ALU
#define PV_SQR_ALU 0
#endif
#ifndef PV_COMBINE_ALU
#define PV_COMBINE_ALU 0
#endif
#ifndef PV_ADDC_ALU
#define PV_ADDC_ALU 0
#endif
GL_DEV u64 sbox7(u64 x) {
#if PV_SBOX_ALU
  u64 x2 = gl_sqr_alu(x);
  u64 x4 = gl_sqr_alu(x2);
  u64 x3 = gl_mul_alu(x, x2);
  return gl_mul_alu(x3, x4);
#elif PV_SQR_ALU  // only the squarings (18 instructions against 20, one partial product fewer)
  u64 x2 = gl_sqr_alu(x);
  u64 x4 = gl_sqr_alu(x2);
  u64 x3 = gl_mul(x, x2);
  return gl_mul(x3, x4);
#else
  u64 x2 = gl_sqr(x);
  u64 x4 = gl_sqr(x2);
  u64 x3 = gl_mul(x, x2);
  return gl_mul(x3, x4);
#endif
}
GL_DEV u64 pv_add_c(u64 a, u64 b_canon) {
#if PV_ADDC_ALU
  return gl_add_c_alu(a, b_canon);
#else
  return gl_add_c(a, b_canon);
#endif
}

// MDS lane recombination: al + ah*2^32 with al = l1:l0, ah = h1:h0 < 2^42, folded to 64 bits.
//   value = l0 + (l1 + h0)*2^32 + (h1 + carry)*2^64,  2^64 == EPS; the multiply-add by EPS can carry once.
GL_DEV u64 mds_combine(u32 l0, u32 l1, u32 h0, u32 h1) {
#if PV_COMBINE_ALU
  // r2 = h1 + carry < 2^22: add r2 * (2^32 - 1) = (r2 - [r2 != 0] : -r2) as a 64-bit number, then one
  // correction by 2^32 - 1 if that wrapped (the wrapped sum is below 2^54, so it cannot wrap again)
  asm("{\n\t.reg .u32 c, e0, e1;\n\t"
      "add.cc.u32 %1, %1, %2;\n\t"
      "addc.u32 %3, %3, 0;\n\t"
      "sub.cc.u32 e0, 0, %3;\n\t"
      "subc.u32 e1, %3, 0;\n\t"
      "add.cc.u32 %0, %0, e0;\n\t"
      "addc.cc.u32 %1, %1, e1;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, c;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "add.u32 %1, %1, c;\n\t"
      "}"
      : "+r"(l0), "+r"(l1), "+r"(h0), "+r"(h1));
  return ((u64)l1 << 32) | l0;
#endif
  asm("{\n\t.reg .u32 c;\n\t"
      "add.cc.u32 %1, %1, %2;\n\t"
      "addc.u32 %3, %3, 0;\n\t"
      "mad.lo.cc.u32 %0, %3, 0xffffffff, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, 0xffffffff, %1;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "mad.lo.cc.u32 %0, c, 0xffffffff, %0;\n\t"
      "madc.hi.u32 %1, c, 0xffffffff, %1;\n\t"
      "}"
      : "+r"(l0), "+r"(l1), "+r"(h0), "+r"(h1));
  return ((u64)l1 << 32) | l0;
}

// Dot-product accumulator: three 96-bit column sums A + B*2^32 + C*2^64 of the 32x32 partial
// products (a0b0 | a0b1 + a1b0 | a1b1). Every partial product is ONE IMAD.WIDE with carry-out plus one
// carry add into the column's top limb; carries never ripple across columns, so only one carry
// predicate is live at a time (the previous 160-bit ripple accumulator made ptxas spill predicates
// through LOP3, 135 of them per partial round). Good for up to 2^31 terms.
struct Acc160 {
  u32 a0, a1, a2, b0, b1, b2, c0, c1, c2;
};
GL_DEV void acc_init(Acc160& a) { a.a0 = a.a1 = a.a2 = a.b0 = a.b1 = a.b2 = a.c0 = a.c1 = a.c2 = 0; }
GL_DEV void acc_mac(Acc160& a, u64 x, u64 y) {
  u32 x0 = (u32)x, x1 = (u32)(x >> 32), y0 = (u32)y, y1 = (u32)(y >> 32);
  asm("mad.lo.cc.u32 %0, %9, %11, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %11, %1;\n\t"
      "addc.u32 %2, %2, 0;\n\t"
      "mad.lo.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.hi.cc.u32 %4, %9, %12, %4;\n\t"
      "addc.u32 %5, %5, 0;\n\t"
      "mad.lo.cc.u32 %3, %10, %11, %3;\n\t"
      "madc.hi.cc.u32 %4, %10, %11, %4;\n\t"
      "addc.u32 %5, %5, 0;\n\t"
      "mad.lo.cc.u32 %6, %10, %12, %6;\n\t"
      "madc.hi.cc.u32 %7, %10, %12, %7;\n\t"
      "addc.u32 %8, %8, 0;"
      : "+r"(a.a0), "+r"(a.a1), "+r"(a.a2), "+r"(a.b0), "+r"(a.b1), "+r"(a.b2), "+r"(a.c0), "+r"(a.c1),
        "+r"(a.c2)
      : "r"(x0), "r"(x1), "r"(y0), "r"(y1));
}
// Column sums -> A + B*2^32 + C*2^64 = lo + 2^64*r2 + 2^96*h, then one fold.
GL_DEV u64 acc_reduce(const Acc160& a) {
  typedef unsigned __int128 u128;
  u128 A = ((u128)a.a2 << 64) | ((u64)a.a1 << 32) | a.a0;
  u128 B = ((u128)a.b2 << 64) | ((u64)a.b1 << 32) | a.b0;
  u128 C = ((u128)a.c2 << 64) | ((u64)a.c1 << 32) | a.c0;
  u128 lowsum = A + (B << 32);          // < 2^129 never happens: A < 2^96, B << 32 < 2^128 - top limbs are tiny
  u128 hi = (lowsum >> 64) + C;         // 2^64 units
  return gl_fold((u64)lowsum, (u32)(u64)hi, (u64)(hi >> 32));
}

// (hi:lo) += a*b as ONE IMAD.WIDE.U32 with accumulate. Written as a mad.lo.cc / madc.hi pair on
// purpose: from `acc += (u64)a * b` nvcc added a zero high word per term, and from a single
// `mad.wide.u32` ptxas split every term into a product plus 3-input IADD3 trees - both double the
// load on the ALU pipe, which is the saturated one.
GL_DEV void mad_wide(u32& lo, u32& hi, u32 a, u32 b) {
  asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\t"
      "madc.hi.u32 %1, %2, %3, %1;"
      : "+r"(lo), "+r"(hi)
      : "r"(a), "r"(b));
}

#ifndef PV_MDS_F64
#define PV_MDS_F64 1
#endif
// MDS on the FP64 pipe. The integer multiplier (FMA-heavy) pipe is the one Poseidon saturates, while
// the DFMA pipe of the B200 (64 per clock per SM) idles. The circulant has entries < 2^6 (row sum 264), so
// the 12x12 product runs on the two 32-bit HALVES of the state as exact integer arithmetic in doubles:
// every sum is < 264 * 2^32 < 2^41 << 2^53. int -> double through the 2^52 bias (the half is the low
// mantissa word, one DADD removes the bias), double -> int by adding the bias back and reading the
// mantissa: low word + 20 bits of the high word. No conversion instructions, no limb shifting or masking
// on the way in, 288 DFMA per layer. (The first version cut the state into three 22-bit limbs so that each
// sum fitted one 32-bit word: 432 DFMA, 72 conversions and a three-way stitch per layer.)
__constant__ double c_mds_circ_d[12];
__constant__ double c_mds_half_d[12];  // (c[i] + c[i+6]) / 2 for i < 6, then (c[i] - c[i+6]) / 2
#ifndef PV_RC_FOLD
#define PV_RC_FOLD 1
#endif
// With PV_RC_FOLD the layer also adds the constants of the round that follows it (poseidon_next_rc_f64):
// they enter as the accumulators' initial values, which costs nothing, and the 64-bit modular additions at
// the head of the next round disappear.
__constant__ double c_mds_next_rc_d[30][2][12];
// PV_CVT (experiment): 1 = int -> double with the conversion instruction (I2F.F64.U32, its own pipe) instead of
// the bias trick (a register-pair move plus a DADD); 2 = also double -> int with F2I.U64.F64.
#ifndef PV_CVT
#define PV_CVT 0
#endif
GL_DEV double u32_to_f64(u32 v) {
#if PV_CVT >= 1
  return __uint2double_rn(v);
#else
  return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
#endif
}
GL_DEV void f64_to_u52(double d, u32& lo, u32& hi) {
#if PV_CVT >= 2
  const unsigned long long t = __double2ull_rz(d);  // the sums are non-negative integers below 2^42
  lo = (u32)t;
  hi = (u32)(t >> 32);
#else
  const double t = d + 4503599627370496.0;
  lo = (u32)__double2loint(t);
  hi = (u32)__double2hiint(t) & 0xFFFFFu;
#endif
}
// The 12x12 circulant product itself is split once by z^12 - 1 = (z^6 - 1)(z^6 + 1): with s = x_lo + x_hi and
// d = x_lo - x_hi (x_lo = x[0..5], x_hi = x[6..11]), U = cyclic_6(s, (c_lo + c_hi)/2) and V = negacyclic_6(d,
// (c_lo - c_hi)/2) give y[r] = U[r] + V[r] and y[r+6] = U[r] - V[r]: 12 + 72 + 12 FP64 operations per plane
// instead of 144. The halved constants are integers for this matrix ({15,14,40,17,18,24}, {2,1,1,-1,-16,4}),
// intermediate values are signed integers below 2^41 in magnitude (constants: half-integers below 2^33), so
// everything stays exact. The folded round constants enter through U and V: U starts from
// (rc[r] + rc[r+6])/2, V from (rc[r] - rc[r+6])/2.
#ifndef PV_MDS_SPLIT
#define PV_MDS_SPLIT 1
#endif
GL_DEV void mds_layer_f64(u64 (&s)[12], int layer) {
  u32 S[2][2][12];  // [half][word][lane]: sum over the low / high halves, each < 2^41 (+ 2^32 of constants)
#pragma unroll
  for (int k = 0; k < 2; k++) {
    double d[12];
#pragma unroll
    for (int i = 0; i < 12; i++) d[i] = u32_to_f64(k == 0 ? (u32)s[i] : (u32)(s[i] >> 32));
#if PV_MDS_SPLIT
    double sm[6], df[6];
#pragma unroll
    for (int j = 0; j < 6; j++) {
      sm[j] = d[j] + d[j + 6];
      df[j] = d[j] - d[j + 6];
    }
#pragma unroll
    for (int r = 0; r < 6; r++) {
#if PV_RC_FOLD
      const double c0 = c_mds_next_rc_d[layer][k][r], c6 = c_mds_next_rc_d[layer][k][r + 6];  // (rc[r] +- rc[r+6]) / 2
      double u = c0, v = c6;
#else
      double u = 0.0, v = 0.0;
#endif
#pragma unroll
      for (int i = 0; i < 6; i++) {
        u = fma(sm[(i + r) % 6], c_mds_half_d[i], u);
        v = fma(df[(i + r) % 6], i + r < 6 ? c_mds_half_d[6 + i] : -c_mds_half_d[6 + i], v);
      }
      double y0 = u + v, y6 = u - v;
      if (r == 0) y0 = fma(d[0], 8.0, y0);
      f64_to_u52(y0, S[k][0][r], S[k][1][r]);
      f64_to_u52(y6, S[k][0][r + 6], S[k][1][r + 6]);
    }
#else
#pragma unroll
    for (int r = 0; r < 12; r++) {
#if PV_RC_FOLD
      double acc = c_mds_next_rc_d[layer][k][r];
      if (r == 0) acc = fma(d[0], 8.0, acc);
#else
      double acc = r == 0 ? d[0] * 8.0 : 0.0;
#endif
#pragma unroll
      for (int i = 0; i < 12; i++) acc = fma(d[(i + r) % 12], c_mds_circ_d[i], acc);
      f64_to_u52(acc, S[k][0][r], S[k][1][r]);
    }
#endif
  }
#pragma unroll
  for (int r = 0; r < 12; r++) s[r] = mds_combine(S[0][0][r], S[0][1][r], S[1][0][r], S[1][1][r]);
}
// MDS on the integer multiplier: 32-bit halves, IMAD.WIDE accumulate, one fold per lane. Used where the
// FP64 variant's extra registers hurt (the quotient kernel's PoseidonGate evaluation).
GL_DEV void mds_layer_int(u64 (&s)[12]) {
  u32 lo[12], hi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    lo[i] = (u32)s[i];
    hi[i] = (u32)(s[i] >> 32);
  }
#pragma unroll
  for (int r = 0; r < 12; r++) {
    u32 al0 = 0, al1 = 0, ah0 = 0, ah1 = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      mad_wide(al0, al1, lo[(i + r) % 12], c_mds_circ[i]);
      mad_wide(ah0, ah1, hi[(i + r) % 12], c_mds_circ[i]);
    }
    if (r == 0) {
      mad_wide(al0, al1, lo[0], c_mds_diag0);
      mad_wide(ah0, ah1, hi[0], c_mds_diag0);
    }
    s[r] = mds_combine(al0, al1, ah0, ah1);
  }
}
// Full round: add constants (unless the previous MDS layer already did), x^7 on every lane, MDS.
// Code size matters more than the last instruction here: ncu showed the fully unrolled permutation
// (90 KB of SASS) stalled ~50% on instruction fetch, the 32 KB L1.5 I-cache thrashing with 24 warps
// per SM spread over the kernel. The 12 s-boxes are therefore issued as PV_SBOX_TRIPS trips over
// 12 / PV_SBOX_TRIPS lanes with the state ROTATED per trip (indices stay compile-time, 24 MOVs per trip).
#ifndef PV_SBOX_TRIPS
#define PV_SBOX_TRIPS 1
#endif
template <bool ADD_RC>
GL_DEV void sbox_layer(u64 (&s)[12], const u64* __restrict__ rc) {
  constexpr int LANES = 12 / PV_SBOX_TRIPS;
#pragma unroll 1
  for (int g = 0; g < PV_SBOX_TRIPS; g++) {
    u64 t[LANES];
#pragma unroll
    for (int j = 0; j < LANES; j++) t[j] = sbox7(ADD_RC ? pv_add_c(s[j], rc[LANES * g + j]) : s[j]);
#pragma unroll
    for (int i = 0; i < 12 - LANES; i++) s[i] = s[i + LANES];
#pragma unroll
    for (int j = 0; j < LANES; j++) s[12 - LANES + j] = t[j];
  }
}
#ifndef PV_INIT_UNROLL
#define PV_INIT_UNROLL 1
#endif
#ifndef PV_INIT_SMEM
#define PV_INIT_SMEM 1
#endif
#ifndef PV_PARTIAL_UNROLL
#define PV_PARTIAL_UNROLL 1
#endif
#define PV_PRAGMA_(x) _Pragma(#x)
#define PV_UNROLL(n) PV_PRAGMA_(unroll n)
// mds_partial_layer_init of the sparse partial-round form: out[c] = sum_r in[r] * init[r-1][c-1] for
// c = 1..11 (11x11, 64-bit entries), one trip per output lane. Also used by the quotient kernel's
// PoseidonGate. Every caller runs 128-thread blocks.
template <bool HYBRID_TABLES>
GL_DEV void partial_init_layer(u64 (&s)[12]) {
  const u64* init = HYBRID_TABLES ? c_h_init : c_fast_init;
#if PV_INIT_SMEM
  // Outputs are parked in shared memory (one column per thread, conflict-free) until all 11 are done: a
  // store per trip and 11 loads at the end, instead of shifting an 11-word register file every trip
  // (220 moves per permutation plus the spills they caused).
  __shared__ u64 sh_o[11][128];
#pragma unroll 1
  for (int c = 0; c < 11; c++) {
    Acc160 a;
    acc_init(a);
#pragma unroll
    for (int r = 1; r < 12; r++) acc_mac(a, s[r], init[(r - 1) * 11 + c]);
    sh_o[c][threadIdx.x] = acc_reduce(a);
  }
#pragma unroll
  for (int i = 1; i < 12; i++) s[i] = sh_o[i - 1][threadIdx.x];
#else
  // Results are shifted through o[] so every index is static.
  u64 o[11];
  PV_UNROLL(PV_INIT_UNROLL)
  for (int c = 0; c < 11; c++) {
    Acc160 a;
    acc_init(a);
#pragma unroll
    for (int r = 1; r < 12; r++) acc_mac(a, s[r], init[(r - 1) * 11 + c]);
#pragma unroll
    for (int i = 0; i < 10; i++) o[i] = o[i + 1];
    o[10] = acc_reduce(a);
  }
#pragma unroll
  for (int i = 1; i < 12; i++) s[i] = o[i - 1];
#endif
}

// The partial rounds that run in the sparse form (the last 22 - PV_DENSE_PARTIAL of them): initial layer, then
// per round one s-box, one 12-term dot product with a single fold, 11 fused multiply-adds.
#ifndef PV_DENSE_PARTIAL
#define PV_DENSE_PARTIAL 0
#endif
GL_DEV void sparse_partial_rounds(u64 (&s)[12]) {
#if !(PV_MDS_F64 && PV_RC_FOLD)  // otherwise the MDS layer before them added the first constants
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = pv_add_c(s[i], c_fast_first[i]);
#endif
  partial_init_layer<true>(s);
PV_UNROLL(PV_PARTIAL_UNROLL)
  for (int r = 0; r < 22 - PV_DENSE_PARTIAL; r++) {
    u64 s0 = sbox7(s[0]);
    s0 = pv_add_c(s0, c_h_rc[r]);  // the last entry is zero
    Acc160 a;
    acc_init(a);
    acc_mac(a, s0, 25);  // MDS[0][0] = 17 + 8
#pragma unroll
    for (int i = 1; i < 12; i++) acc_mac(a, s[i], c_h_w_hat[r * 11 + i - 1]);
#pragma unroll
    for (int i = 1; i < 12; i++) s[i] = gl_mad(s0, c_h_v[r * 11 + i - 1], s[i]);
    s[0] = acc_reduce(a);
  }
}

// In-place permutation; inputs may be any u64 representatives, outputs likewise (not canonical).
// One copy of the round body (s-boxes + MDS) serves the first four full rounds, the PV_DENSE_PARTIAL dense
// partial rounds (s-box on lane 0 only) and the last four full rounds (see the I-cache note above).
//
// Why dense partial rounds at all: a sparse partial round is 23 multiplications by 64-bit constants = 92
// IMAD.WIDE on the multiplier pipe, the pipe the whole permutation is bound by; the textbook round is one
// s-box and the small-entry MDS, which runs as 288 DFMA on the otherwise idle FP64 pipe (and adds the next
// round's constants for free). Moving some of the 22 rounds over balances the two pipes.
GL_DEV void poseidon_permute(u64 (&s)[12]) {
#if PV_MDS_F64 && PV_RC_FOLD
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
    const u64* rc = c_rc + 12 * 26 * half;  // the rounds not preceded by an FP64 MDS layer add their constants
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = pv_add_c(s[i], rc[i]);
    const int nrounds = half ? 4 : 4 + PV_DENSE_PARTIAL, layer0 = half ? 4 + PV_DENSE_PARTIAL : 0;
#pragma unroll 1
    for (int r = 0; r < nrounds; r++) {
      if (r < 4) sbox_layer<false>(s, rc);
      else s[0] = sbox7(s[0]);
      mds_layer_f64(s, layer0 + r);
    }
    if (half == 0) sparse_partial_rounds(s);
  }
#else
  static_assert(PV_DENSE_PARTIAL == 0, "dense partial rounds need the FP64 MDS with folded constants");
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
      sbox_layer<true>(s, c_rc + 12 * (26 * half + r));
#if PV_MDS_F64
      mds_layer_f64(s, 0);
#else
      mds_layer_int(s);
#endif
    }
    if (half == 0) sparse_partial_rounds(s);
  }
#endif
}


// ---- low-latency variant: 16 lanes per permutation (12 active), one state word per lane ----------
// The thread-per-permutation kernel above is the throughput path; one permutation there is ~20k
// dependent-ish instructions = ~50 us. Small Merkle levels, FRI trees and voting-sized commits have too
// few permutations to fill the machine, so their time is that latency times the number of dependent
// steps. Here the 12 s-boxes of a full round run in parallel lanes and the MDS row of each lane reads
// the other 11 words from a shared-memory exchange. The eight full rounds are executed in the textbook
// form (constants, s-boxes, MDS); the 22 partial rounds as one linear recurrence driven by the s-box
// outputs (coop_partial_rounds below; PV_COOP_LINEAR=0 restores the textbook partial rounds): ~4.2k serial
// instructions per permutation, 6.5 us. The same permutation as the sparse partial-round form of the
// throughput kernel, so results are bit-identical.
__device__ u64 g_rc[372];  // lane-indexed reads: global/L1, not the constant bank (divergent index); 12 zeros appended
#define COOP_XCH_WORDS 48

// s: this lane's state word (lanes 12..15 of the group carry garbage and must be ignored by the caller).
// xch: COOP_XCH_WORDS u64 of shared memory private to the 16-lane group: two buffers of 24 words. Every
// lane stores its word TWICE, at [lane] and [lane + 12], so that the circulant row of lane l is the 12
// consecutive words [l, l + 12): one base address and immediate offsets instead of a compare-and-wrap per
// load (25 of the ~185 instructions of a round, all of them on the dependent chain of a single warp).
// mask: the lanes that execute this call together (__syncwarp mask) - the whole warp when both 16-lane
// groups of a warp run in lockstep, one half when they may diverge (tree climbing, merkle.cuh).
// The constants of round r + 1 enter as the initial value of round r's MDS accumulators (their 32-bit halves
// join the low / high column sums: free), and they are requested one round ahead, so neither the modular
// addition nor the load sits on the dependent chain.
#ifndef PV_COOP_LINEAR
#define PV_COOP_LINEAR 1
#endif
// One full round: s-boxes in parallel lanes, exchange, circulant row; `rc` = the NEXT round's constant of this lane.
GL_DEV u64 coop_round(u64 s, bool sbox_here, u64 rc, u32 lane, u32 li, bool act, u64* buf, u32 mask) {
  if (sbox_here) s = sbox7(s);
  if (act) {
    buf[lane] = s;
    buf[lane + 12] = s;
  }
  __syncwarp(mask);
  const u64* row = buf + li;
  u32 al0 = (u32)rc, al1 = 0, ah0 = (u32)(rc >> 32), ah1 = 0, bl0 = 0, bl1 = 0, bh0 = 0, bh1 = 0;
#pragma unroll
  for (int i = 0; i < 12; i += 2) {
    const u64 v0 = row[i], v1 = row[i + 1];
    mad_wide(al0, al1, (u32)v0, c_mds_circ[i]);
    mad_wide(ah0, ah1, (u32)(v0 >> 32), c_mds_circ[i]);
    mad_wide(bl0, bl1, (u32)v1, c_mds_circ[i + 1]);
    mad_wide(bh0, bh1, (u32)(v1 >> 32), c_mds_circ[i + 1]);
  }
  if (lane == 0) {
    mad_wide(al0, al1, (u32)s, c_mds_diag0);
    mad_wide(ah0, ah1, (u32)(s >> 32), c_mds_diag0);
  }
  u64 lo = (((u64)al1 << 32) | al0) + (((u64)bl1 << 32) | bl0);   // < 2^43: no overflow
  u64 hi = (((u64)ah1 << 32) | ah0) + (((u64)bh1 << 32) | bh0);
  return mds_combine((u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32));
}

#if PV_COOP_LINEAR
// The 22 partial rounds without an MDS layer (poseidon_tables.hpp, build_linear_tables). In the textbook form a
// partial round costs one warp as much as a full round - the s-box runs on lane 0 while 15 lanes wait, then the
// whole exchange + circulant row - and a single warp's time is the number of instructions it issues
// (profiles/r2_coop16_latency_vs_load.txt). Here the partial rounds are one linear recurrence driven by the s-box
// outputs y_0..y_21: each lane keeps two unreduced dot-product accumulators (the 20 later s-box INPUTS x_2..x_21
// and the 12 words of the state after round 25), a round is broadcast x_k -> y_k = x_k^7 in every lane -> two
// multiply-adds -> fold of the accumulator that holds x_(k+1): ~115 instructions against ~185, plus 22
// multiply-adds per lane to set the accumulators up from the state after round 3.
__device__ u64 g_lin_p[2 * 11 * 16];
__device__ u64 g_lin_c[2 * 16];
__device__ u64 g_lin_coef[22 * 2 * 16];

GL_DEV void acc_set(Acc160& a, u64 c) {
  acc_init(a);
  a.a0 = (u32)c;
  a.a1 = (u32)(c >> 32);
}
GL_DEV u64 shfl16(u32 mask, u64 v, u32 src) {
  u32 lo = __shfl_sync(mask, (u32)v, src, 16), hi = __shfl_sync(mask, (u32)(v >> 32), src, 16);
  return ((u64)hi << 32) | lo;
}

// s = this lane's word of t_0 (state after round 3 with rc[4] added); returns its word of t_22 (rc[26] added).
GL_DEV u64 coop_partial_rounds(u64 s, u32 lane, bool act, u64* buf, u32 mask) {
  if (act) buf[lane] = s;
  u64 c0 = __ldg(&g_lin_coef[lane]), c1 = __ldg(&g_lin_coef[16 + lane]);
  Acc160 a0, a1;
  acc_set(a0, __ldg(&g_lin_c[lane]));
  acc_set(a1, __ldg(&g_lin_c[16 + lane]));
  const u64 rcx = __ldg(&g_rc[60]);
  u32 xl0 = (u32)rcx, xl1 = 0, xh0 = (u32)(rcx >> 32), xh1 = 0;  // x_1 = rc[5][0] + row 0 of the MDS matrix
  __syncwarp(mask);
#pragma unroll
  for (int i = 1; i < 12; i++) {
    const u64 v = buf[i];
    acc_mac(a0, v, __ldg(&g_lin_p[(i - 1) * 16 + lane]));
    acc_mac(a1, v, __ldg(&g_lin_p[(11 + i - 1) * 16 + lane]));
    mad_wide(xl0, xl1, (u32)v, c_mds_circ[i]);
    mad_wide(xh0, xh1, (u32)(v >> 32), c_mds_circ[i]);
  }
  // k = 0: x_0 is lane 0's word. From here on slot 0 is kept reduced (one multiply-add with a single fold per round:
  // every round hands one of these accumulators on as the next s-box input, and all lanes execute that fold anyway),
  // slot 1 stays an unreduced column accumulator until its first word is due (k = 17).
  u64 y = sbox7(buf[0]);
  u64 e0 = gl_mad(y, c0, acc_reduce(a0));
  acc_mac(a1, y, c1);
  c0 = __ldg(&g_lin_coef[32 + lane]);
  c1 = __ldg(&g_lin_coef[48 + lane]);
  mad_wide(xl0, xl1, (u32)y, c_mds_circ[0] + c_mds_diag0);
  mad_wide(xh0, xh1, (u32)(y >> 32), c_mds_circ[0] + c_mds_diag0);
  u64 x = mds_combine(xl0, xl1, xh0, xh1);
  // k = 1..16: x_(k+1) is slot 0 of lane k - 1
#pragma unroll 1
  for (u32 k = 1; k <= 16; k++) {
    y = sbox7(x);
    e0 = gl_mad(y, c0, e0);
    acc_mac(a1, y, c1);
    c0 = __ldg(&g_lin_coef[(k + 1) * 32 + lane]);
    c1 = __ldg(&g_lin_coef[(k + 1) * 32 + 16 + lane]);
    x = shfl16(mask, e0, k - 1);
  }
  // k = 17..20: x_(k+1) is slot 1 of lane k - 5; slot 0 is spent
  u64 e1 = acc_reduce(a1);
#pragma unroll 1
  for (u32 k = 17; k <= 20; k++) {
    y = sbox7(x);
    e1 = gl_mad(y, c1, e1);
    c1 = __ldg(&g_lin_coef[(k + 1) * 32 + 16 + lane]);
    x = shfl16(mask, e1, k - 5);
  }
  return gl_mad(sbox7(x), c1, e1);
}
#endif

GL_DEV u64 poseidon_permute_coop(u64 s, u32 lane, u64* xch, u32 mask = 0xffffffffu) {
  const bool act = lane < 12;
  const u32 li = act ? lane : 0;
  s = pv_add_c(s, __ldg(&g_rc[li]));
  u64 rc_next = __ldg(&g_rc[12 + li]);
#if PV_COOP_LINEAR
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
    const u64 rc = rc_next;
    rc_next = __ldg(&g_rc[(r < 3 ? r + 2 : 27) * 12 + li]);  // the last trip fetches round 26's successor
    s = coop_round(s, true, rc, lane, li, act, xch + (r & 1) * 24, mask);
  }
  s = coop_partial_rounds(s, lane, act, xch, mask);
#pragma unroll 1
  for (int r = 26; r < 30; r++) {
    const u64 rc = rc_next;
    rc_next = __ldg(&g_rc[(r < 28 ? r + 2 : 30) * 12 + li]);  // the last two trips read the zero padding
    s = coop_round(s, true, rc, lane, li, act, xch + (r & 1) * 24, mask);
  }
#else
#pragma unroll 1
  for (int r = 0; r < 30; r++) {
    const u64 rc = rc_next;
    rc_next = __ldg(&g_rc[(r < 28 ? r + 2 : 30) * 12 + li]);  // the last two trips read the zero padding
    const bool full = r < 4 || r >= 26;
    s = coop_round(s, full || lane == 0, rc, lane, li, act, xch + (r & 1) * 24, mask);
  }
#endif
  return s;
}

}  // namespace qpzk
