// Prover-stage kernels for sm_100a: Z / partial products (H8), quotient evaluation over the LDE
// coset (H9), openings (H10), FRI batch-combine / divide-by-linear (H11), fold (H12), proof-of-work
// grinding (H13) and query gathers (H14) - SURVEY.md §8(a).
//
// Replaces the bodies of `all_wires_permutation_partial_products`, `compute_quotient_polys` +
// `eval_vanishing_poly_base_batch` + `Gate::eval_unfiltered_base_batch`, `OpeningSet::new`,
// `PolynomialBatch::prove_openings`, `fri_committed_trees`, `fri_proof_of_work` and
// `fri_prover_query_rounds` of qp-plonky2 1.1.1 (plonk/prover.rs, plonk/vanishing_poly.rs,
// gates/*.rs, fri/oracle.rs, fri/prover.rs; un-vendored), reached from
// /root/reference/wormhole/prover/src/lib.rs:233-237,
// /root/reference/wormhole/aggregator/src/circuits/tree.rs:136 and /root/reference/voting/src/lib.rs:356.
//
// Data layout: every oracle's LDE is column-major [width][N] in bit-reversed row order (see
// merkle.cuh), so a thread that owns leaf position L reads column c at lde[c*N + L]: all loads in
// the quotient kernel are unit-stride across the warp. Natural LDE index i = bitrev(L).
#pragma once
#include "merkle.cuh"
#include "ntt.cuh"
#include "transcript.cuh"

namespace qpzk {

// ids = position in plonky2's default gate serializer list
enum : u32 {
  G_ARITHMETIC = 0,
  G_ARITHMETIC_EXT = 1,
  G_BASE_SUM_2 = 2,
  G_CONSTANT = 3,
  G_COSET_INTERPOLATION = 4,
  G_EXPONENTIATION = 5,
  G_MUL_EXT = 8,
  G_NOOP = 9,
  G_POSEIDON_MDS = 10,
  G_POSEIDON = 11,
  G_PUBLIC_INPUT = 12,
  G_RANDOM_ACCESS = 13,
  G_REDUCING_EXT = 14,
  G_REDUCING = 15,
};
GL_HD bool gate_is_recursion_only(u32 id) {
  return id == G_ARITHMETIC_EXT || id == G_COSET_INTERPOLATION || id == G_EXPONENTIATION || id == G_MUL_EXT ||
         id == G_POSEIDON_MDS || id == G_RANDOM_ACCESS || id == G_REDUCING_EXT || id == G_REDUCING;
}

#define QPZK_MAX_GATES 16

// Everything the quotient / Z kernels need to know about the circuit (passed by value).
struct CircuitDesc {
  u32 degree_bits, rate_bits, quotient_degree_bits;
  u32 num_wires, num_routed, num_constants, num_challenges, num_partial_products, qdf;
  u32 num_selectors, num_gates, num_gate_constraints;
  u32 gate_id[QPZK_MAX_GATES], gate_param[QPZK_MAX_GATES], gate_selector[QPZK_MAX_GATES];
  u32 gate_param2[QPZK_MAX_GATES], gate_param3[QPZK_MAX_GATES];  // RandomAccess copies / extra constants; CosetInterpolation degree
  u32 group_lo[QPZK_MAX_GATES], group_hi[QPZK_MAX_GATES];  // indexed by selector index
  const u64* coset_aux;  // CosetInterpolation: [2^bits] subgroup points then [2^bits] barycentric weights (device)
};

// Challenges (transcript.cuh) reach the kernels as a pointer into the device-resident transcript: they are
// produced on the device and never visit the host on the proving path.

// ---------------------------------------------------------------------------------------------
// H8  Z and partial products
// ---------------------------------------------------------------------------------------------
// One thread per (row, challenge): the nchunks chunk quotients prod(num)/prod(den) and their product.
// wires / cs are value columns on the subgroup, [.][n]. k_is in constant-like global memory.
__global__ void __launch_bounds__(128)
k_zs_chunk_quotients(const u64* __restrict__ wires, const u64* __restrict__ cs, const u64* __restrict__ k_is,
                     CircuitDesc d, const Challenges* __restrict__ chp, RootTab tab, u64* __restrict__ chunk_q /*[nch][nchunks][n]*/,
                     u64* __restrict__ row_prod /*[nch][n]*/) {
  const u64 n = (u64)1 << d.degree_bits;
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32 c = blockIdx.y;
  const u32 nchunks = d.num_partial_products + 1;
  const u64 x = root_pow(tab, i);
  const u64 beta = __ldg(&chp->beta[c]), gamma = __ldg(&chp->gamma[c]);
  const u64 bx = gl_mul(beta, x);
  u64 nums[12], dens[12];  // nchunks <= 12
  u64 dprod = 1;
  for (u32 k = 0; k < nchunks; k++) {
    u64 num = 1, den = 1;
    for (u32 j = k * d.qdf; j < (k + 1) * d.qdf && j < d.num_routed; j++) {
      u64 w = wires[(u64)j * n + i];
      u64 sg = cs[(u64)(d.num_constants + j) * n + i];
      u64 nu = gl_add(gl_mad(bx, k_is[j], w), gamma);
      u64 de = gl_add(gl_mad(beta, sg, w), gamma);
      num = gl_mul(num, nu);
      den = gl_mul(den, de);
    }
    nums[k] = num;
    dens[k] = den;
    dprod = gl_mul(dprod, den);
  }
  // Montgomery batch inversion of the nchunks denominators: one field inversion per row
  u64 pre[12];
  {
    u64 acc = 1;
    for (u32 k = 0; k < nchunks; k++) {
      pre[k] = acc;
      acc = gl_mul(acc, dens[k]);
    }
  }
  u64 inv_all = gl_inv(dprod);
  u64 total = 1;
  for (int k = (int)nchunks - 1; k >= 0; k--) {
    u64 dinv = gl_mul(inv_all, pre[k]);    // 1 / dens[k]
    inv_all = gl_mul(inv_all, dens[k]);    // 1 / prod(dens[0..k))
    nums[k] = gl_mul(nums[k], dinv);
  }
  for (u32 k = 0; k < nchunks; k++) {
    chunk_q[((u64)c * nchunks + k) * n + i] = nums[k];
    total = gl_mul(total, nums[k]);
  }
  row_prod[(u64)c * n + i] = total;
}

// Exclusive prefix product along rows: z[0] = 1, z[i+1] = z[i] * row_prod[i]. One CTA per challenge.
__global__ void __launch_bounds__(1024) k_prefix_product(const u64* __restrict__ row_prod, u64* __restrict__ z, u64 n) {
  __shared__ u64 part[1024];
  const u64* in = row_prod + (u64)blockIdx.x * n;
  u64* out = z + (u64)blockIdx.x * n;
  const u32 t = threadIdx.x, T = blockDim.x;
  const u64 per = (n + T - 1) / T;
  const u64 lo = (u64)t * per, hi = lo + per < n ? lo + per : n;
  u64 p = 1;
  for (u64 i = lo; i < hi; i++) p = gl_mul(p, in[i]);
  part[t] = p;
  __syncthreads();
  for (u32 off = 1; off < T; off <<= 1) {  // inclusive Hillis-Steele scan
    u64 v = t >= off ? part[t - off] : 1;
    __syncthreads();
    part[t] = gl_mul(part[t], v);
    __syncthreads();
  }
  u64 acc = t ? part[t - 1] : 1;
  for (u64 i = lo; i < hi; i++) {
    out[i] = gl_canon(acc);
    acc = gl_mul(acc, in[i]);
  }
}

// zs_pp[ch] (Z) is already in place; fill pp[ch][k][i] = Z[i] * prod_{j<=k} chunk_q[ch][j][i].
__global__ void k_partial_products(const u64* __restrict__ chunk_q, const u64* __restrict__ z, u32 nch, u32 npp,
                                   u64 n, u64* __restrict__ pp /*[nch][npp][n]*/) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 c = blockIdx.y;
  u64 acc = z[(u64)c * n + i];
  for (u32 k = 0; k < npp; k++) {
    acc = gl_mul(acc, chunk_q[((u64)c * (npp + 1) + k) * n + i]);
    pp[((u64)c * npp + k) * n + i] = gl_canon(acc);
  }
}

// ---------------------------------------------------------------------------------------------
// H9  quotient evaluation
// ---------------------------------------------------------------------------------------------
// Running sum_t alpha^t * term_t for every challenge, 160-bit lazily reduced accumulators.
// Reduction of the constraint terms with powers of alpha (`reduce_with_powers_multi`): term t of the list
// contributes alpha_c^t * term to challenge c. The powers come from a per-proof table apw[c][t]
// (uniform address across the warp), and terms are multiply-accumulated into column accumulators
// with ONE fold at the end; a gate's constraints go to a gate-local accumulator that is folded, scaled
// by the gate's filter and added to the total when the gate is done - one multiplication by the filter
// per gate instead of one per constraint, and no running alpha-power products.
#define QPZK_APW_STRIDE 256
struct AlphaAcc {
  Acc160 acc[2];   // current target: the total outside gates, the gate-local sum inside
  Acc160 tot[2];
  const u64* apw;  // [2][QPZK_APW_STRIDE]
  u32 t;           // index of the next term
  u32 nch;
};
GL_DEV void aa_emit(AlphaAcc& a, u64 term) {
  QPZK_CHECK(a.t < QPZK_APW_STRIDE);
#pragma unroll
  for (int c = 0; c < 2; c++)
    if (c < (int)a.nch) acc_mac(a.acc[c], __ldg(a.apw + c * QPZK_APW_STRIDE + a.t), term);
  a.t++;
}
GL_DEV void aa_gate_begin(AlphaAcc& a, u32 base) {
#pragma unroll
  for (int c = 0; c < 2; c++) {
    a.tot[c] = a.acc[c];
    acc_init(a.acc[c]);
  }
  a.t = base;
}
GL_DEV void aa_gate_end(AlphaAcc& a, u64 filter) {
#pragma unroll
  for (int c = 0; c < 2; c++) {
    if (c < (int)a.nch) acc_mac(a.tot[c], acc_reduce(a.acc[c]), filter);
    a.acc[c] = a.tot[c];
  }
}

struct WireRow {  // column-major LDE accessor for one leaf position
  const u64* base;
  u64 N;
#ifdef QPZK_CHECKED
  u32 width;
#endif
  GL_DEV u64 operator[](u32 c) const {
    QPZK_CHECK(c < width);
    return __ldg(base + (u64)c * N);
  }
};
#ifdef QPZK_CHECKED
#define QPZK_WIRE_ROW(base, stride, width) WireRow{base, stride, width}
#else
#define QPZK_WIRE_ROW(base, stride, width) WireRow{base, stride}
#endif

// PoseidonGate::eval_unfiltered (123 constraints) with the fast partial rounds; emits in order.
// The s-box input wires (29..134) are consumed in index order, a few per round, each right before a long
// dependent computation: fetched just in time they stalled the kernel on memory latency (ncu: 41 % of the
// warp stall samples were long_scoreboard at 16 warps per SM). Every group of wires is therefore loaded
// one group AHEAD - the next trip's four while the current four go through their s-boxes, the next partial
// round's one during the current round.
GL_DEV void poseidon_gate_eval(const WireRow& w, AlphaAcc& a) {
  u64 swap = w[24];
  aa_emit(a, gl_mul(swap, gl_sub(swap, 1)));
  u64 s[12];
  u64 nx[4];  // the next four s-box inputs, in flight
#pragma unroll
  for (int q = 0; q < 4; q++) nx[q] = w[29 + q];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    u64 lhs = w[i], rhs = w[i + 4], delta = w[25 + i];
    aa_emit(a, gl_sub(gl_mul(swap, gl_sub(rhs, lhs)), delta));
    s[i] = gl_add(lhs, delta);
    s[i + 4] = gl_sub(rhs, delta);
  }
#pragma unroll
  for (int i = 8; i < 12; i++) s[i] = w[i];
  u32 next = 33;  // index of the first wire not yet requested
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll 1
    for (int g = 0; g < 3; g++) {  // rotate-by-4 so indices stay static (see poseidon.cuh)
      u64 t[4], in[4];
      if (r != 0) {
#pragma unroll
        for (int q = 0; q < 4; q++) in[q] = nx[q];
#pragma unroll
        for (int q = 0; q < 4; q++) nx[q] = w[next + q];  // wires 33..64, then 65..68 (only 65 is used)
        next += 4;
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        u64 v = gl_add_c(s[q], c_rc[12 * r + 4 * g + q]);
        if (r != 0) {
          aa_emit(a, gl_sub(v, in[q]));
          v = in[q];
        }
        t[q] = sbox7(v);
      }
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = s[i + 4];
      s[8] = t[0]; s[9] = t[1]; s[10] = t[2]; s[11] = t[3];
    }
    mds_layer_f64(s, 29);  // layer 29: no constants folded in
  }
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_add_c(s[i], c_fast_first[i]);
  partial_init_layer<false>(s);
  u64 in_next = nx[0];  // wire 65
#pragma unroll 1
  for (int r = 0; r < 22; r++) {
    u64 in = in_next;
    in_next = w[66 + r];  // the last trip requests wire 87, the first input of the second half
    aa_emit(a, gl_sub(s[0], in));
    u64 s0 = gl_add_c(sbox7(in), c_fast_rc[r]);
    Acc160 acc;
    acc_init(acc);
    acc_mac(acc, s0, 25);
#pragma unroll
    for (int i = 1; i < 12; i++) acc_mac(acc, s[i], c_fast_w_hat[r * 11 + i - 1]);
#pragma unroll
    for (int i = 1; i < 12; i++) s[i] = gl_mad(s0, c_fast_v[r * 11 + i - 1], s[i]);
    s[0] = acc_reduce(acc);
  }
  nx[0] = in_next;
#pragma unroll
  for (int q = 1; q < 4; q++) nx[q] = w[87 + q];
  next = 91;
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll 1
    for (int g = 0; g < 3; g++) {
      u64 t[4], in[4];
#pragma unroll
      for (int q = 0; q < 4; q++) in[q] = nx[q];
      if (next < 135) {
#pragma unroll
        for (int q = 0; q < 4; q++) nx[q] = w[next + q];
        next += 4;
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        u64 v = gl_add_c(s[q], c_rc[12 * (26 + r) + 4 * g + q]);
        aa_emit(a, gl_sub(v, in[q]));
        t[q] = sbox7(in[q]);
      }
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = s[i + 4];
      s[8] = t[0]; s[9] = t[1]; s[10] = t[2]; s[11] = t[3];
    }
    mds_layer_f64(s, 29);
  }
#pragma unroll 1
  for (int g = 0; g < 3; g++) {
#pragma unroll
    for (int q = 0; q < 4; q++) aa_emit(a, gl_sub(s[q], w[12 + 4 * g + q]));
    u64 t0 = s[0], t1 = s[1], t2 = s[2], t3 = s[3];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = s[i + 4];
    s[8] = t0; s[9] = t1; s[10] = t2; s[11] = t3;
  }
}

// ---- the recursion gate set: constraint order as in oracle/plonk.hpp ----
GL_DEV gl2 ext_at(const WireRow& w, u32 i) { return gl2_make(w[i], w[i + 1]); }
GL_DEV void aa_emit_ext(AlphaAcc& a, gl2 v) {
  aa_emit(a, v.a);
  aa_emit(a, v.b);
}
GL_DEV void recursion_gate_eval(u32 id, u32 p1, u32 p2, u32 p3, const u64* __restrict__ aux, const WireRow& w,
                                u64 c0, u64 c1, AlphaAcc& a) {
  switch (id) {
    case G_ARITHMETIC_EXT:
      for (u32 t = 0; t < p1; t++) {
        gl2 m0 = ext_at(w, 8 * t), m1 = ext_at(w, 8 * t + 2), ad = ext_at(w, 8 * t + 4), o = ext_at(w, 8 * t + 6);
        aa_emit_ext(a, gl2_sub(o, gl2_add(gl2_scale(gl2_mul(m0, m1), c0), gl2_scale(ad, c1))));
      }
      break;
    case G_MUL_EXT:
      for (u32 t = 0; t < p1; t++) {
        gl2 m0 = ext_at(w, 6 * t), m1 = ext_at(w, 6 * t + 2), o = ext_at(w, 6 * t + 4);
        aa_emit_ext(a, gl2_sub(o, gl2_scale(gl2_mul(m0, m1), c0)));
      }
      break;
    case G_POSEIDON_MDS:
      for (u32 r = 0; r < 12; r++) {
        Acc160 sa, sb;
        acc_init(sa);
        acc_init(sb);
        for (u32 i = 0; i < 12; i++) {
          u32 j = i + r >= 12 ? i + r - 12 : i + r;
          acc_mac(sa, w[2 * j], c_mds_circ[i]);
          acc_mac(sb, w[2 * j + 1], c_mds_circ[i]);
        }
        if (r == 0) {
          acc_mac(sa, w[0], c_mds_diag0);
          acc_mac(sb, w[1], c_mds_diag0);
        }
        aa_emit_ext(a, gl2_sub(ext_at(w, 2 * (12 + r)), gl2_make(acc_reduce(sa), acc_reduce(sb))));
      }
      break;
    case G_RANDOM_ACCESS: {
      const u32 bits = p1, copies = p2, extra = p3, vec = 1u << bits;
      const u32 routed = (2 + vec) * copies + extra;
      for (u32 cp = 0; cp < copies; cp++) {
        const u32 w0 = (2 + vec) * cp, b0 = routed + cp * bits;
        u64 rec = 0;
        for (u32 i = 0; i < bits; i++) {
          u64 b = w[b0 + i];
          aa_emit(a, gl_mul(b, gl_sub(b, 1)));
        }
        for (u32 i = bits; i-- > 0;) rec = gl_add(gl_add(rec, rec), w[b0 + i]);
        aa_emit(a, gl_sub(rec, w[w0]));
        u64 items[64];  // vec <= 64 (checked at circuit creation)
        for (u32 i = 0; i < vec; i++) items[i] = w[w0 + 2 + i];
        u32 len = vec;
        for (u32 i = 0; i < bits; i++) {
          u64 b = w[b0 + i];
          len >>= 1;
          for (u32 j = 0; j < len; j++) items[j] = gl_add(items[2 * j], gl_mul(b, gl_sub(items[2 * j + 1], items[2 * j])));
        }
        aa_emit(a, gl_sub(items[0], w[w0 + 1]));
      }
      for (u32 i = 0; i < extra; i++) aa_emit(a, gl_sub(i == 0 ? c0 : c1, w[(2 + vec) * copies + i]));
      break;
    }
    case G_REDUCING:
    case G_REDUCING_EXT: {
      const bool ext = id == G_REDUCING_EXT;
      const u32 ncf = p1, start_accs = 6 + ncf * (ext ? 2 : 1);
      gl2 alpha = ext_at(w, 2), acc = ext_at(w, 4);
      for (u32 i = 0; i < ncf; i++) {
        gl2 cf = ext ? ext_at(w, 6 + 2 * i) : gl2_make(w[6 + i], 0);
        gl2 ai = i == ncf - 1 ? ext_at(w, 0) : ext_at(w, start_accs + 2 * i);
        aa_emit_ext(a, gl2_sub(gl2_add(gl2_mul(acc, alpha), cf), ai));
        acc = ai;
      }
      break;
    }
    case G_EXPONENTIATION: {
      const u32 nb = p1;
      const u64 base = w[0];
      u64 prev_iv = 1;
      for (u32 i = 0; i < nb; i++) {
        u64 prev = i == 0 ? 1 : gl_sqr(prev_iv);
        u64 bit = w[1 + (nb - 1 - i)];
        u64 iv = w[2 + nb + i];
        aa_emit(a, gl_sub(gl_mul(prev, gl_add(gl_mul(bit, base), gl_sub(1, bit))), iv));
        prev_iv = iv;
      }
      aa_emit(a, gl_sub(w[1 + nb], prev_iv));
      break;
    }
    case G_COSET_INTERPOLATION: {
      const u32 npoints = 1u << p1, degree = p2, nint = (npoints - 2) / (degree - 1);
      const u32 sp = 1 + 2 * npoints, sv = sp + 2, si = sv + 2, ss = si + 4 * nint;
      const u64* dom = aux;
      const u64* wt = aux + npoints;
      const u64 shift = w[0];
      const gl2 x = ext_at(w, sp), xs = ext_at(w, ss);
      aa_emit_ext(a, gl2_sub(x, gl2_scale(xs, shift)));
      gl2 ev = gl2_make(0, 0), pr = gl2_make(1, 0);
      u32 lo = 0, hi = degree;
      for (u32 seg = 0; seg <= nint; seg++) {
        for (u32 i = lo; i < hi; i++) {
          gl2 term = gl2_sub(xs, gl2_make(dom[i], 0));
          gl2 val = ext_at(w, 1 + 2 * i);
          ev = gl2_add(gl2_mul(ev, term), gl2_scale(gl2_mul(val, pr), wt[i]));
          pr = gl2_mul(pr, term);
        }
        if (seg == nint) break;
        gl2 ie = ext_at(w, si + 2 * seg), ip = ext_at(w, si + 2 * (nint + seg));
        aa_emit_ext(a, gl2_sub(ie, ev));
        aa_emit_ext(a, gl2_sub(ip, pr));
        ev = ie;
        pr = ip;
        lo = 1 + (degree - 1) * (seg + 1);
        hi = lo + degree - 1 < npoints ? lo + degree - 1 : npoints;
      }
      aa_emit_ext(a, gl2_sub(ext_at(w, sv), ev));
      break;
    }
    default:
      break;
  }
}

// Per-circuit table for L_0 on the quotient domain: out[i] = 1 / (n * (x_i - 1)), x_i = g * w^i.
__global__ void k_build_l0_den_inv(u64* __restrict__ out, RootTab tab, u32 degree_bits, u32 lb) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ((u64)1 << lb)) return;
  u64 x = gl_mul(GL_GEN, root_pow(tab, i));
  out[i] = gl_canon(gl_inv(gl_mul(((u64)1 << degree_bits) % GL_P, gl_sub(x, 1))));
}

// One thread per LDE leaf position. out[ch][i] (natural index i) = vanishing(x_i) / Z_H(x_i).
// RECURSION = false is the wormhole / voting gate set; the recursion gates are compiled only into the
// <true> instantiation so that they cost the common case neither registers nor instruction cache.
template <bool RECURSION>
__global__ void __launch_bounds__(128, RECURSION ? 3 : 4)
k_quotient(const u64* __restrict__ cs_lde, const u64* __restrict__ wires_lde, const u64* __restrict__ zs_lde,
           u64 cs_stride, u64 wires_stride, u64 zs_stride, u32 step_bits, const u64* __restrict__ k_is,
           CircuitDesc d, const Challenges* __restrict__ chp, const u64* __restrict__ pi_hash, const u64* __restrict__ zh /*[2^qdb]*/,
           const u64* __restrict__ zh_inv, const u64* __restrict__ apw /* alpha powers [2][QPZK_APW_STRIDE] */,
           const u64* __restrict__ l0_den_inv /* [2^lb]: 1 / (n (x_i - 1)) */, RootTab tab /* size degree_bits + qdb */,
           u64 q_begin, u64 q_end /* positions to evaluate */, u32 out_bitrev, u64* __restrict__ out) {
  const u32 lb = d.degree_bits + d.quotient_degree_bits;  // quotient domain bits
  const u64 lde = (u64)1 << lb;
  u64 q = q_begin + (u64)blockIdx.x * blockDim.x + threadIdx.x;  // position in the bit-reversed quotient domain
  if (q >= q_end) return;
  const u64 i = __brevll(q) >> (64 - lb);                  // natural index on g*<w_lde>
  // position inside the committed LDE (rate_bits >= qdb): natural index i*step -> leaf rev(i*step)
  const u32 full_bits = d.degree_bits + d.rate_bits;
  const u64 leaf = __brevll(i << step_bits) >> (64 - full_bits);
  const u64 inext = (i + ((u64)1 << d.quotient_degree_bits)) & (lde - 1);
  const u64 leaf_next = __brevll(inext << step_bits) >> (64 - full_bits);
  const u64 x = gl_mul(GL_GEN, root_pow(tab, i));
  const u32 nch = d.num_challenges, npp = d.num_partial_products;
  QPZK_CHECK(leaf < ((u64)1 << full_bits) && leaf_next < ((u64)1 << full_bits));
  QPZK_CHECK(d.num_gates <= QPZK_MAX_GATES && d.num_selectors <= QPZK_MAX_GATES && nch <= 2);
  const u32 zs_width = nch * (1 + npp);
  WireRow cs = QPZK_WIRE_ROW(cs_lde + leaf, cs_stride, d.num_constants + d.num_routed);
  WireRow w = QPZK_WIRE_ROW(wires_lde + leaf, wires_stride, d.num_wires);
  WireRow zs = QPZK_WIRE_ROW(zs_lde + leaf, zs_stride, zs_width);
  WireRow zsn = QPZK_WIRE_ROW(zs_lde + leaf_next, zs_stride, zs_width);
  (void)zs_width;
  const u64 zhx = zh[i & (((u64)1 << d.quotient_degree_bits) - 1)];

  AlphaAcc a;
  a.nch = nch;
  a.apw = apw;
  a.t = 0;
  for (int c = 0; c < 2; c++) acc_init(a.acc[c]);
  // L_0(x) (Z_i(x) - 1),  L_0(x) = Z_H(x) / (n (x - 1)); the denominators are a per-circuit table
  {
    u64 l0 = gl_mul(zhx, l0_den_inv[i]);
    for (u32 c = 0; c < nch; c++) aa_emit(a, gl_mul(l0, gl_sub(zs[c], 1)));
  }
  // partial-product checks: term (c, k) = prev * prod_j num_j - next * prod_j den_j over chunk k of qdf routed
  // wires. Chunks outside, challenges inside: the chunk's wires, sigmas and k_is are requested as ONE batch
  // before any arithmetic and serve both challenges (fetched one by one inside the product loop, every
  // factor waited for its own loads: 38 % of the kernel's stall samples, 97 % of them long_scoreboard).
  // The alpha-power table makes the emission order free, so each term goes straight to its slot.
  const u32 nchunks = npp + 1;
  const u32 pp_base = a.t;
  {
    u64 bx[2], prev[2], betas[2], gammas[2];
    for (u32 c = 0; c < nch; c++) {
      betas[c] = __ldg(&chp->beta[c]);
      gammas[c] = __ldg(&chp->gamma[c]);
      bx[c] = gl_mul(betas[c], x);
      prev[c] = zs[c];
    }
    for (u32 k = 0; k < nchunks; k++) {
      const u32 j0 = k * d.qdf;
      u64 next[2];
      for (u32 c = 0; c < nch; c++) next[c] = k + 1 < nchunks ? zs[nch + c * npp + k] : zsn[c];
      if (d.qdf == 8 && j0 + 8 <= d.num_routed) {
        u64 wv[8], sg[8], kk[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
          wv[e] = w[j0 + e];
          sg[e] = cs[d.num_constants + j0 + e];
          kk[e] = __ldg(k_is + j0 + e);
        }
        for (u32 c = 0; c < nch; c++) {
          const u64 beta = betas[c], gamma = gammas[c];
          u64 num = gl_add(gl_mad(bx[c], kk[0], wv[0]), gamma);
          u64 den = gl_add(gl_mad(beta, sg[0], wv[0]), gamma);
#pragma unroll
          for (int e = 1; e < 8; e++) {
            num = gl_mul(num, gl_add(gl_mad(bx[c], kk[e], wv[e]), gamma));
            den = gl_mul(den, gl_add(gl_mad(beta, sg[e], wv[e]), gamma));
          }
          a.t = pp_base + c * nchunks + k;
          aa_emit(a, gl_sub(gl_mul(prev[c], num), gl_mul(next[c], den)));
        }
      } else {  // other chunk sizes, ragged last chunk
        for (u32 c = 0; c < nch; c++) {
          const u64 beta = betas[c], gamma = gammas[c];
          u64 num = 1, den = 1;
          for (u32 j = j0; j < j0 + d.qdf && j < d.num_routed; j++) {
            u64 wv = w[j];
            num = gl_mul(num, gl_add(gl_mad(bx[c], k_is[j], wv), gamma));
            den = gl_mul(den, gl_add(gl_mad(beta, cs[d.num_constants + j], wv), gamma));
          }
          a.t = pp_base + c * nchunks + k;
          aa_emit(a, gl_sub(gl_mul(prev[c], num), gl_mul(next[c], den)));
        }
      }
      for (u32 c = 0; c < nch; c++) prev[c] = next[c];
    }
    a.t = pp_base + nch * nchunks;
  }
  // gate constraints: slot j gets sum_g filter_g * c_{g,j}; each gate restarts at alpha^base
  const u32 base_t = a.t;
  for (u32 g = 0; g < d.num_gates; g++) {
    if (d.gate_id[g] == G_NOOP) continue;
    const u32 si = d.gate_selector[g];
    QPZK_CHECK(si < d.num_selectors && d.group_hi[si] <= d.num_gates);
    const u64 s = cs[si];
    u64 filter = 1;
    for (u32 j = d.group_lo[si]; j < d.group_hi[si]; j++)
      if (j != g) filter = gl_mul(filter, gl_sub((u64)j, s));
    if (d.num_selectors > 1) filter = gl_mul(filter, gl_sub((u64)0xFFFFFFFFu, s));
    aa_gate_begin(a, base_t);
    switch (d.gate_id[g]) {
      case G_NOOP:
        break;
      case G_CONSTANT:
        for (u32 t = 0; t < d.gate_param[g]; t++) aa_emit(a, gl_sub(cs[d.num_selectors + t], w[t]));
        break;
      case G_PUBLIC_INPUT:
        for (u32 t = 0; t < 4; t++) aa_emit(a, gl_sub(w[t], pi_hash[t]));
        break;
      case G_BASE_SUM_2: {
        // One pass from the top limb down, eight limbs requested at a time (fetched one by one the Horner
        // chain waited for every load). Term 0 is the sum check, term 1 + t the range check of limb t: the
        // alpha-power table lets each go straight to its slot.
        const u32 nl = d.gate_param[g];
        const u32 t0 = a.t;
        const u64 w0 = w[0];
        u64 sum = 0;
        int t = (int)nl - 1;
        for (; t >= 7; t -= 8) {
          u64 l[8];
#pragma unroll
          for (int e = 0; e < 8; e++) l[e] = w[1 + t - e];
#pragma unroll
          for (int e = 0; e < 8; e++) {
            sum = gl_add(gl_add(sum, sum), l[e]);
            a.t = t0 + 1 + (u32)(t - e);
            aa_emit(a, gl_mul(l[e], gl_sub(l[e], 1)));
          }
        }
        for (; t >= 0; t--) {
          u64 l = w[1 + t];
          sum = gl_add(gl_add(sum, sum), l);
          a.t = t0 + 1 + (u32)t;
          aa_emit(a, gl_mul(l, gl_sub(l, 1)));
        }
        a.t = t0;
        aa_emit(a, gl_sub(sum, w0));
        a.t = t0 + 1 + nl;
        break;
      }
      case G_ARITHMETIC: {
        const u64 c0 = cs[d.num_selectors], c1 = cs[d.num_selectors + 1];
        const u32 nops = d.gate_param[g];
        u64 nx[4];  // the next operation's wires, requested one operation ahead
#pragma unroll
        for (int e = 0; e < 4; e++) nx[e] = w[e];
        for (u32 t = 0; t < nops; t++) {
          const u64 m0 = nx[0], m1 = nx[1], ad = nx[2], o = nx[3];
          if (t + 1 < nops) {
#pragma unroll
            for (int e = 0; e < 4; e++) nx[e] = w[4 * (t + 1) + e];
          }
          u64 comp = gl_mad(gl_mul(m0, m1), c0, gl_mul(ad, c1));
          aa_emit(a, gl_sub(o, comp));
        }
        break;
      }
      case G_POSEIDON:
        poseidon_gate_eval(w, a);
        break;
      default:
        if (RECURSION)
          recursion_gate_eval(d.gate_id[g], d.gate_param[g], d.gate_param2[g], d.gate_param3[g], d.coset_aux, w,
                              cs[d.num_selectors], cs[d.num_selectors + 1], a);
        break;
    }
    aa_gate_end(a, filter);
  }
  const u64 zi = zh_inv[i & (((u64)1 << d.quotient_degree_bits) - 1)];
  for (u32 c = 0; c < nch; c++) out[(u64)c * lde + (out_bitrev ? q : i)] = gl_canon(gl_mul(acc_reduce(a.acc[c]), zi));
}

// data[c][m] *= base^m via a two-level power table (coset (i)fft shift removal).
__global__ void k_scale_by_powers(u64* __restrict__ data, u64 n, RootTab tab) {
  u64 m = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  u64* p = data + (u64)blockIdx.y * n + m;
  *p = gl_canon(gl_mul(*p, root_pow(tab, m)));
}

// ---------------------------------------------------------------------------------------------
// H10  openings: evaluate base-field coefficient columns at an extension point
// ---------------------------------------------------------------------------------------------
// pw[m] = z^m (ext), m < n.
__global__ void k_ext_powers(const u64* __restrict__ zp /* [2], device */, u64 n, u64* __restrict__ pw /*[n][2]*/) {
  u64 m = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  gl2 r = gl2_make(1, 0), b = gl2_make(__ldg(zp), __ldg(zp + 1));
  for (u64 e = m; e; e >>= 1) {
    if (e & 1) r = gl2_mul(r, b);
    b = gl2_mul(b, b);
  }
  pw[2 * m] = gl_canon(r.a);
  pw[2 * m + 1] = gl_canon(r.b);
}
// out[p] = sum_m coeffs[p][m] * pw[m]. grid (polynomials, chunks): CTA (p, c) sums coefficients
// [c * chunk, (c + 1) * chunk) of polynomial p. With one chunk the result goes straight to out; with several
// (long polynomials: one CTA per polynomial left a 2^18-coefficient opening at 1.7 waves of 8 warps per SM,
// 2.7 ms for 255 polynomials) the partial sums go to part[p][c] and k_eval_reduce adds them up.
__global__ void __launch_bounds__(256)
k_eval_at_ext(const u64* __restrict__ coeffs, u64 n, u64 chunk, const u64* __restrict__ pw, u64* __restrict__ out /*[.][2]*/,
              u64* __restrict__ part /*[.][chunks][2]*/) {
  __shared__ u64 sa[256], sb[256];
  const u64* c = coeffs + (u64)blockIdx.x * n;
  const u64 lo = (u64)blockIdx.y * chunk, hi = lo + chunk < n ? lo + chunk : n;
  Acc160 a, b;
  acc_init(a);
  acc_init(b);
  // four coefficients (and their powers) requested before the first multiply-add: with one CTA per polynomial the
  // loop was a chain of dependent-latency loads (37 us per launch at 2^14 coefficients, 5 launches per proof)
  u64 m = lo + threadIdx.x;
  const u64 st = blockDim.x;
  for (; m + 3 * st < hi; m += 4 * st) {
    u64 v[4], p0[4], p1[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      v[j] = c[m + j * st];
      p0[j] = pw[2 * (m + j * st)];
      p1[j] = pw[2 * (m + j * st) + 1];
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      acc_mac(a, v[j], p0[j]);
      acc_mac(b, v[j], p1[j]);
    }
  }
  for (; m < hi; m += st) {
    u64 v = c[m];
    acc_mac(a, v, pw[2 * m]);
    acc_mac(b, v, pw[2 * m + 1]);
  }
  sa[threadIdx.x] = acc_reduce(a);
  sb[threadIdx.x] = acc_reduce(b);
  __syncthreads();
  for (u32 off = blockDim.x / 2; off; off >>= 1) {
    if (threadIdx.x < off) {
      sa[threadIdx.x] = gl_add(sa[threadIdx.x], sa[threadIdx.x + off]);
      sb[threadIdx.x] = gl_add(sb[threadIdx.x], sb[threadIdx.x + off]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    u64* o = gridDim.y == 1 ? out + 2 * (u64)blockIdx.x : part + 2 * ((u64)blockIdx.x * gridDim.y + blockIdx.y);
    o[0] = gl_canon(sa[0]);
    o[1] = gl_canon(sb[0]);
  }
}
__global__ void k_eval_reduce(const u64* __restrict__ part, u32 npolys, u32 chunks, u64* __restrict__ out) {
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npolys) return;
  u64 a = 0, b = 0;
  for (u32 c = 0; c < chunks; c++) {
    a = gl_add(a, part[2 * ((u64)p * chunks + c)]);
    b = gl_add(b, part[2 * ((u64)p * chunks + c) + 1]);
  }
  out[2 * p] = gl_canon(a);
  out[2 * p + 1] = gl_canon(b);
}

// ---------------------------------------------------------------------------------------------
// H11  FRI batch combine: comp[m] = sum_j alpha^j f_j[m] over a list of coefficient columns
// ---------------------------------------------------------------------------------------------
struct PolyList {  // up to 4 oracles' coefficient arrays, concatenated in order
  const u64* base[4];
  u32 count[4];
  u32 noracles;
};
// apow[j] = alpha^j (ext) for j < total (device array [total][2]). comp: SoA [2][n].
__global__ void __launch_bounds__(128)
k_fri_compose(PolyList pl, u64 n, const u64* __restrict__ apow, u64* __restrict__ comp) {
  u64 m = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  Acc160 a, b;
  acc_init(a);
  acc_init(b);
  u32 j = 0;
  for (u32 o = 0; o < pl.noracles; o++) {
    // four columns requested ahead of their multiply-adds (the loop was one dependent-latency load per polynomial:
    // 115 us for the 257 polynomials of a wormhole proof)
    const u64* col = pl.base[o] + m;
    u32 p = 0;
    for (; p + 3 < pl.count[o]; p += 4, j += 4) {
      u64 v[4];
#pragma unroll
      for (int i = 0; i < 4; i++) v[i] = col[(u64)(p + i) * n];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        acc_mac(a, v[i], apow[2 * (j + i)]);
        acc_mac(b, v[i], apow[2 * (j + i) + 1]);
      }
    }
    for (; p < pl.count[o]; p++, j++) {
      u64 v = col[(u64)p * n];
      acc_mac(a, v, apow[2 * j]);
      acc_mac(b, v, apow[2 * j + 1]);
    }
  }
  comp[m] = acc_reduce(a);
  comp[n + m] = acc_reduce(b);
}

// divide_by_linear: q[m-1] = p[m] + z*q[m] (descending), q[n-1] = 0. One CTA per polynomial.
// p, q: SoA [2][n] ext coefficient vectors; they must NOT alias (the replay re-reads p while neighbours write q).
// Each thread owns a segment; a segment acts on the carry entering it from above as the affine map
// x -> L + Z*x (L = its local Horner value, Z = z^len). The carries are a SUFFIX SCAN of these maps
// under composition: log2(T) shared-memory steps instead of a T-step sequential walk (which was
// 0.24 ms of pure latency per call at n = 2^14).
__global__ void __launch_bounds__(1024)
k_divide_by_linear(const u64* __restrict__ p, u64* __restrict__ q, u64 n, const u64* __restrict__ zp /* [2], device */) {
  __shared__ u64 la[1024], lb[1024], za[1024], zb[1024];
  const gl2 z = gl2_make(__ldg(zp), __ldg(zp + 1));
  const u32 t = threadIdx.x, T = blockDim.x;
  const u64 per = (n + T - 1) / T;
  const u64 lo = (u64)t * per < n ? (u64)t * per : n, hi = lo + per < n ? lo + per : n;  // segment [lo, hi)
  gl2 acc = gl2_make(0, 0);
  for (u64 m = hi; m-- > lo;) acc = gl2_add(gl2_make(p[m], p[n + m]), gl2_mul(z, acc));
  gl2 zl = gl2_make(1, 0);
  {
    gl2 bse = z;
    for (u64 e = hi - lo; e; e >>= 1) {
      if (e & 1) zl = gl2_mul(zl, bse);
      bse = gl2_mul(bse, bse);
    }
  }
  // inclusive suffix scan: after it, (L, Z)[t] = f_t o f_{t+1} o ... o f_{T-1}; carry into t = L[t+1]
  gl2 L = acc, Z = zl;
  for (u32 d = 1; d < T; d <<= 1) {
    la[t] = L.a; lb[t] = L.b; za[t] = Z.a; zb[t] = Z.b;
    __syncthreads();
    if (t + d < T) {
      gl2 L2 = gl2_make(la[t + d], lb[t + d]), Z2 = gl2_make(za[t + d], zb[t + d]);
      L = gl2_add(L, gl2_mul(Z, L2));   // f_t..(x) = L + Z*(L2 + Z2*x)
      Z = gl2_mul(Z, Z2);
    }
    __syncthreads();
  }
  la[t] = L.a; lb[t] = L.b;
  __syncthreads();
  acc = t + 1 < T ? gl2_make(la[t + 1], lb[t + 1]) : gl2_make(0, 0);
  // replay with the true carry-in and write q[m-1]
  for (u64 m = hi; m-- > lo;) {
    acc = gl2_add(gl2_make(p[m], p[n + m]), gl2_mul(z, acc));
    if (m > 0) {
      q[m - 1] = gl_canon(acc.a);
      q[n + m - 1] = gl_canon(acc.b);
    }
  }
  if (hi == n && lo < hi) {
    q[n - 1] = 0;
    q[2 * n - 1] = 0;
  }
}

// The same division for long polynomials, spread over the machine (one CTA walked 2^18 coefficients in 1.2 ms):
// every thread of the grid owns a segment of `per` coefficients.
//   k_divlin_local : the segment's affine map x -> L + Z x (L = its Horner value, Z = z^per), to maps[t] = (L, Z)
//   k_divlin_scan  : ONE CTA turns the maps into the carries: carry[t] = value entering segment t from above
//                    (suffix composition over all T segments: each thread folds a run of them, then the
//                    1024-wide shared-memory scan above, then the runs are replayed)
//   k_divlin_apply : the segment replayed from its carry, q written
struct DivMap {
  u64 la, lb, za, zb;
};
__global__ void __launch_bounds__(256)
k_divlin_local(const u64* __restrict__ p, u64 n, u64 per, const u64* __restrict__ zp, DivMap* __restrict__ maps) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 lo = t * per;
  if (lo >= n) return;
  const u64 hi = lo + per < n ? lo + per : n;
  const gl2 z = gl2_make(__ldg(zp), __ldg(zp + 1));
  gl2 acc = gl2_make(0, 0), zl = gl2_make(1, 0);
  for (u64 m = hi; m-- > lo;) {
    acc = gl2_add(gl2_make(p[m], p[n + m]), gl2_mul(z, acc));
    zl = gl2_mul(zl, z);
  }
  DivMap o;
  o.la = acc.a; o.lb = acc.b; o.za = zl.a; o.zb = zl.b;
  maps[t] = o;
}
// maps[T] -> carry[T][2] (carry into segment t = composition of segments t+1 .. T-1 applied to 0)
__global__ void __launch_bounds__(1024)
k_divlin_scan(const DivMap* __restrict__ maps, u64 T, u64* __restrict__ carry) {
  __shared__ u64 la[1024], lb[1024], za[1024], zb[1024];
  const u32 t = threadIdx.x, NT = blockDim.x;
  const u64 run = (T + NT - 1) / NT;
  const u64 lo = (u64)t * run < T ? (u64)t * run : T, hi = lo + run < T ? lo + run : T;
  // this thread's run as one map: f_lo o f_{lo+1} o ... o f_{hi-1}
  gl2 L = gl2_make(0, 0), Z = gl2_make(1, 0);
  for (u64 i = hi; i-- > lo;) {  // prepend f_i: f_i(L + Z x) = L_i + Z_i L + Z_i Z x
    const gl2 Li = gl2_make(maps[i].la, maps[i].lb), Zi = gl2_make(maps[i].za, maps[i].zb);
    L = gl2_add(Li, gl2_mul(Zi, L));
    Z = gl2_mul(Zi, Z);
  }
  for (u32 d = 1; d < NT; d <<= 1) {  // inclusive suffix scan over the runs
    la[t] = L.a; lb[t] = L.b; za[t] = Z.a; zb[t] = Z.b;
    __syncthreads();
    if (t + d < NT) {
      gl2 L2 = gl2_make(la[t + d], lb[t + d]), Z2 = gl2_make(za[t + d], zb[t + d]);
      L = gl2_add(L, gl2_mul(Z, L2));
      Z = gl2_mul(Z, Z2);
    }
    __syncthreads();
  }
  la[t] = L.a; lb[t] = L.b;
  __syncthreads();
  gl2 acc = t + 1 < NT ? gl2_make(la[t + 1], lb[t + 1]) : gl2_make(0, 0);  // carry into the last segment of the run
  for (u64 i = hi; i-- > lo;) {
    carry[2 * i] = gl_canon(acc.a);
    carry[2 * i + 1] = gl_canon(acc.b);
    acc = gl2_add(gl2_make(maps[i].la, maps[i].lb), gl2_mul(gl2_make(maps[i].za, maps[i].zb), acc));
  }
}
__global__ void __launch_bounds__(256)
k_divlin_apply(const u64* __restrict__ p, u64* __restrict__ q, u64 n, u64 per, const u64* __restrict__ zp,
               const u64* __restrict__ carry) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 lo = t * per;
  if (lo >= n) return;
  const u64 hi = lo + per < n ? lo + per : n;
  const gl2 z = gl2_make(__ldg(zp), __ldg(zp + 1));
  gl2 acc = gl2_make(carry[2 * t], carry[2 * t + 1]);
  for (u64 m = hi; m-- > lo;) {
    acc = gl2_add(gl2_make(p[m], p[n + m]), gl2_mul(z, acc));
    if (m > 0) {
      q[m - 1] = gl_canon(acc.a);
      q[n + m - 1] = gl_canon(acc.b);
    }
  }
  if (hi == n) {
    q[n - 1] = 0;
    q[2 * n - 1] = 0;
  }
}

// final = q0 * s + q1  (ext scalar s), SoA [2][n]
__global__ void k_ext_axpy(const u64* __restrict__ q0, const u64* __restrict__ q1, const u64* __restrict__ sp /* [2], device */,
                           u64 n, u64* __restrict__ out) {
  u64 m = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const gl2 s = gl2_make(__ldg(sp), __ldg(sp + 1));
  gl2 v = gl2_add(gl2_mul(gl2_make(q0[m], q0[n + m]), s), gl2_make(q1[m], q1[n + m]));
  out[m] = gl_canon(v.a);
  out[n + m] = gl_canon(v.b);
}

// ---------------------------------------------------------------------------------------------
// H12  FRI fold and leaf packing
// ---------------------------------------------------------------------------------------------
// out[m] = sum_{j<arity} in[m*arity + j] * beta^j  (ext Horner). in: SoA [2][n_in], out: SoA [2][n_in/arity]
__global__ void k_fri_fold(const u64* __restrict__ in, u64 n_in, u32 arity_bits, const u64* __restrict__ bp /* [2], device */,
                           u64* __restrict__ out) {
  u64 n_out = n_in >> arity_bits;
  u64 m = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_out) return;
  const gl2 beta = gl2_make(__ldg(bp), __ldg(bp + 1));
  u32 arity = 1u << arity_bits;
  gl2 acc = gl2_make(0, 0);
  for (int j = (int)arity - 1; j >= 0; j--) {
    u64 idx = (m << arity_bits) + j;
    acc = gl2_add(gl2_mul(acc, beta), gl2_make(in[idx], in[n_in + idx]));
  }
  out[m] = gl_canon(acc.a);
  out[n_out + m] = gl_canon(acc.b);
}
// SoA [2][N] ext values -> AoS [N][2] (leaf = `arity` consecutive ext values = 2*arity felts)
__global__ void k_ext_interleave(const u64* __restrict__ soa, u64 N, u64* __restrict__ aos) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  ulonglong2 v = make_ulonglong2(soa[i], soa[N + i]);
  reinterpret_cast<ulonglong2*>(aos)[i] = v;
}

// ---------------------------------------------------------------------------------------------
// H13  proof of work: smallest w with leading_zeros(perm(state with w at pos)[7]) >= min_lz
// ---------------------------------------------------------------------------------------------
struct PowState {
  u64 s[12];
};
__global__ void __launch_bounds__(128)
k_pow_grind(PowState st, u32 pos, u32 min_lz, u64 start, u64 count, unsigned long long* __restrict__ best) {
  u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  u64 cand = start + t;
  if (cand >= *best) return;  // a smaller witness was already found
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = st.s[i];
#pragma unroll
  for (int i = 0; i < 12; i++)
    if (i == (int)pos) s[i] = cand;
  poseidon_permute(s);
  u64 resp = gl_canon(s[7]);
  u32 lz = resp ? (u32)__clzll((long long)resp) : 64;
  if (lz >= min_lz) atomicMin(best, (unsigned long long)cand);
}

// The same search driven from the device-resident transcript, without a host round trip per window: the
// sponge state and the input position come from `T`, the grid strides over the candidates and a thread
// stops at the first candidate of its own that is not below the best witness found so far. Every candidate
// below the final witness has then been tried by the thread that owns it, so the result is still the
// SMALLEST valid witness. T->pow_witness must be ~0 on entry (k_transcript_init).
__global__ void __launch_bounds__(128)
k_pow_grind_dev(TranscriptDev* __restrict__ T, u32 min_lz) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  u64 cand = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u32 pos = T->in_len;
  u64 s0[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s0[i] = T->state[i];
  volatile unsigned long long* best = reinterpret_cast<volatile unsigned long long*>(&T->pow_witness);
  while (cand < *best) {
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = i == (int)pos ? cand : s0[i];
    poseidon_permute(s);
    const u64 resp = gl_canon(s[7]);
    const u32 lz = resp ? (u32)__clzll((long long)resp) : 64;
    if (lz >= min_lz) atomicMin(reinterpret_cast<unsigned long long*>(&T->pow_witness), (unsigned long long)cand);
    cand += stride;
  }
}

// ---------------------------------------------------------------------------------------------
// H14  batched openings: for each query q: salted row (column-major source) + Merkle path
// ---------------------------------------------------------------------------------------------
// grid = nq blocks. out[q] = [width felts][L*4 felts]
// A multi-GPU shard serves the leaves [leaf0, leaf1) it holds and writes ZEROS for the others: the ranks'
// outputs are then summed.
__global__ void k_gather_openings(const u64* __restrict__ lde, u64 row_stride, u64 col_stride, u32 width,
                                  const u64* __restrict__ levels, u32 log_n, u32 cap_height,
                                  const u64* __restrict__ leaf_idx, u32 shift_bits, u64 leaf0, u64 leaf1,
                                  u64* __restrict__ out) {
  const u32 q = blockIdx.x;
  const u64 leaf = leaf_idx[q] >> shift_bits;
  const u32 L = log_n - cap_height;
  u64* o = out + (u64)q * (width + 4 * L);
  QPZK_CHECK(leaf < ((u64)1 << log_n) && cap_height <= log_n);
  if (leaf < leaf0 || leaf >= leaf1) {
    for (u32 e = threadIdx.x; e < width + 4 * L; e += blockDim.x) o[e] = 0;
    return;
  }
  for (u32 c = threadIdx.x; c < width; c += blockDim.x) o[c] = lde[leaf * row_stride + (u64)c * col_stride];
  const u64 twoN = (u64)2 << log_n;
  for (u32 e = threadIdx.x; e < 4 * L; e += blockDim.x) {
    u32 l = e >> 2;
    u64 off = twoN - (twoN >> l);
    o[width + e] = levels[(off + ((leaf >> l) ^ 1)) * 4 + (e & 3)];
  }
}

}  // namespace qpzk
