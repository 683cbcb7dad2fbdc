// Goldilocks NTT kernels for sm_100a: batched IFFT (values -> coefficients) and coset low-degree
// extension, shared-memory staged, at most two passes over HBM.
//
// Replaces `PolynomialValues::ifft`, `PolynomialCoeffs::lde` + `coset_fft` and the
// `transpose` + `reverse_index_bits_in_place` that follow them inside
// `PolynomialBatch::from_values / from_coeffs` (qp-plonky2 1.1.1 fri/oracle.rs and
// qp-plonky2-field fft.rs, un-vendored; reached from
// /root/reference/wormhole/prover/src/lib.rs:233-237 and
// /root/reference/wormhole/circuit/src/circuit.rs:98-108).
//
// Design (not plonky2's):
//  * The rate-2^r LDE of a degree-<n column on g*<w_N> is computed as 2^r independent n-point
//    coset NTTs (shift g*w_N^t): no zero padding, no wasted butterflies. Natural LDE index
//    i = t + 2^r*m maps to bit-reversed leaf index rev_r(t)*n + rev_k(m), so coset t owns a
//    contiguous block of n leaves and a DIF transform (natural in, bit-reversed out) lands every
//    value in its final place. The reference's transpose and bit-reversal passes do not exist here.
//  * n = n1*n2. Pass A: for 16 adjacent columns j2 of the [n1][n2] view, an n1-point DIF over j1 in
//    shared memory, then the inter-pass twiddle w^(j2*k1). Pass B: n2-point DIF along contiguous
//    rows. Every global access is a full 128-byte line.
//  * Twiddles come from two-level tables (w^e = lo[e & m] * hi[e >> lk]) that stay L1/L2 resident.
#pragma once
#include <cooperative_groups.h>

#include "gl.cuh"

namespace qpzk {
namespace cg = cooperative_groups;

struct RootTab {
  const u64* lo;  // [2^lk]      root^e
  const u64* hi;  // [2^(k-lk)]  root^(e << lk)
  int k, lk;
};
GL_DEV u64 root_pow(const RootTab& t, u64 e) {
  e &= ((u64)1 << t.k) - 1;
  u64 l = t.lo[e & (((u64)1 << t.lk) - 1)];
  u64 h = t.hi[e >> t.lk];
  return gl_mul(l, h);
}

GL_DEV u32 brev(u32 x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0; }

// tables: lo[e] = root^e, hi[e] = root^(e << lk)
__global__ void k_build_root_tab(u64* lo, u64* hi, u64 root, int k, int lk) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ((u64)1 << lk)) lo[i] = gl_canon(gl_pow(root, i));
  if (i < ((u64)1 << (k - lk))) hi[i] = gl_canon(gl_pow(root, i << lk));
}
// pm[t][m] = (shift * w_N^t)^m for t < 2^r, m < n   (coset pre-multipliers; shift = g for commits,
// g^(arity^i) for the i-th FRI round)
__global__ void k_build_coset_pm(u64* pm, u64 shift0, u64 wN, int k, int r) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  u64 n = (u64)1 << k;
  if (i >= (n << r)) return;
  u64 t = i >> k, m = i & (n - 1);
  u64 shift = gl_mul(shift0, gl_pow(wN, t));
  pm[i] = gl_canon(gl_pow(shift, m));
}

// Twiddle matrix M[k1][j] = root^(j*k1), k1 < 2^a, j < 2^(k-a), stored at (k1 << (k-a)) + j: the factors
// between the first 2^a-point stage of a 2^k-point DIF and the rest (pass A's inter-pass twiddle with
// a = log n1; the first radix-16 stage of a long single-CTA transform with a = 4). One coalesced load per
// element instead of two table lookups and a multiplication.
__global__ void k_build_tw_matrix(u64* out, RootTab tab, int k, int a) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ((u64)1 << k)) return;
  u64 k1 = i >> (k - a), j = i & (((u64)1 << (k - a)) - 1);
  out[i] = gl_canon(root_pow(tab, j * k1));
}

// ---- register-resident radix-16 butterflies -------------------------------------------------------
// plonky2's power-of-two roots of unity are powers of TWO in Goldilocks up to order 64:
// POWER_OF_TWO_GENERATOR^(2^26) = 8, so w_16 = 2^12, w_8 = 2^24, w_4 = 2^48 (2^96 = -1, 2^192 = 1).
// A 16-point DFT therefore needs no multiplier at all: its 17 non-trivial twiddles are shifts
// followed by one fold, and a 2^lg-point transform is ceil(lg/4) such stages with ONE general
// multiplication per element between stages (instead of one per butterfly per layer).

// x * 2^s (mod p), 0 <= s < 96; s is a compile-time constant after unrolling.
GL_DEV u64 gl_mul_pow2(u64 x, int s) {
  typedef unsigned __int128 u128;
  if (s == 0) return x;
  if (s < 64) {
    u128 v = (u128)x << s;
    u64 hi = (u64)(v >> 64);
    return gl_fold((u64)v, (u32)hi, hi >> 32);
  }
  u128 w = (u128)x << (s - 64);  // x * 2^s = w * 2^64 = 2^64 * w0 + 2^96 * (w >> 32)
  return gl_fold(0, (u32)(u64)w, (u64)(w >> 32));
}

// DIF butterfly with twiddle 2^e, 0 <= e < 192 (2^(96+t) = -2^t: swap the operands of the subtraction)
GL_DEV void bfly_pow2(u64& a, u64& b, int e) {
  u64 s = gl_add(a, b);
  u64 d = e >= 96 ? gl_sub(b, a) : gl_sub(a, b);
  a = s;
  b = gl_mul_pow2(d, e >= 96 ? e - 96 : e);
}

// In-register 2^LOGR-point DIF DFT (LOGR <= 4); x[p] ends up holding X[rev_LOGR(p)].
template <int LOGR, bool INV>
GL_DEV void dft_regs(u64 (&x)[1 << LOGR]) {
#pragma unroll
  for (int l = 0; l < LOGR; l++) {
    const int half = (1 << LOGR) >> (l + 1);
    const int unit = 192 / (2 * half);  // w_{2*half} = 2^unit
#pragma unroll
    for (int blk = 0; blk < (1 << l); blk++) {
#pragma unroll
      for (int j = 0; j < half; j++) {
        int e = (INV ? 192 - unit * j : unit * j) % 192;
        bfly_pow2(x[blk * 2 * half + j], x[blk * 2 * half + j + half], e);
      }
    }
  }
}

// Shared-memory tile layouts. `cnt` transforms of length len = 2^lg:
//  INTERLEAVED : element e of transform c at e*cnt + c        (pass A: 16 adjacent columns)
//  otherwise   : at c*pitch + e + (e >> 4), pitch = len + len/16 + 1. The 1-in-16 padding keeps every
//                radix-16 stage (stride q = 1, 16, 256 ...) and the transposed read of the inverse
//                pass free of bank conflicts.
GL_HD u32 tile_pitch(u32 len) { return len + (len >> 4) + 1; }
template <bool INTERLEAVED>
GL_DEV u32 tile_pos(u32 c, u32 e, u32 cnt_log, u32 pitch) {
  return INTERLEAVED ? (e << cnt_log) + c : c * pitch + e + (e >> 4);
}

// Tile accessors for the stages. A stage reads its 2^LOGR inputs through LD and writes its outputs
// through ST; inside a pass both are the shared-memory tile, but the FIRST stage of a kernel reads
// straight from global memory (all of a thread's loads are issued before any arithmetic, and one
// shared-memory round trip disappears) and the LAST stage may write straight to global memory.
// (checked build: an access must fall inside the dynamic shared memory the kernel was launched with)
GL_DEV u32 dyn_smem_words() {
  u32 bytes;
  asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(bytes));
  return bytes >> 3;
}
template <bool INTERLEAVED>
struct TileLd {
  const u64* sm; u32 cnt_log, pitch;
  GL_DEV u64 operator()(u32 c, u32 e) const {
    QPZK_CHECK(tile_pos<INTERLEAVED>(c, e, cnt_log, pitch) < dyn_smem_words());
    return sm[tile_pos<INTERLEAVED>(c, e, cnt_log, pitch)];
  }
};
template <bool INTERLEAVED>
struct TileSt {
  u64* sm; u32 cnt_log, pitch;
  GL_DEV void operator()(u32 c, u32 e, u64 v) const {
    QPZK_CHECK(tile_pos<INTERLEAVED>(c, e, cnt_log, pitch) < dyn_smem_words());
    sm[tile_pos<INTERLEAVED>(c, e, cnt_log, pitch)] = v;
  }
};
struct NoPre {
  GL_DEV void operator()() const {}
};
// Inter-stage twiddle w_m^(e_lo * k1) of a stage over blocks of length m inside a 2^lg-point transform:
//  TwTable  : from a table tw[e] = w_L^e (shared memory), index (e_lo * k1) << sh, sh = log(L / m)
//  TwMatrix : from a global matrix t[(k1 << q_log) + e_lo] (k_build_tw_matrix), consecutive threads read
//             consecutive words
struct TwTable {
  const u64* tw; int sh;
  GL_DEV u64 operator()(u32 e_lo, u32 k1) const {
    QPZK_CHECK(((e_lo * k1) << sh) < dyn_smem_words());
    return tw[(e_lo * k1) << sh];
  }
};
struct TwMatrix {
  const u64* t; u32 q_log;
  GL_DEV u64 operator()(u32 e_lo, u32 k1) const { return t[((u64)k1 << q_log) + e_lo]; }
};

// One stage: 2^LOGR-point DFTs over the elements base + i*q of every block of length m = 2^m_log,
// then the inter-stage twiddle w_m^(e_lo * k1) = tw(e_lo, k1); output k1 goes to sub-block rev(k1) (DIF
// order). `pre` runs after the first item's loads have been issued and before any shared-memory twiddle
// is read (a first stage uses it to fill the twiddle table and __syncthreads()).
template <int LOGR, bool INTERLEAVED, bool INV, class TW, class LD, class ST, class PRE>
GL_DEV void dif_stage(TW tw, int lg, int m_log, u32 cnt_log, LD ld, ST st, PRE pre) {
  constexpr int R = 1 << LOGR;
  const u32 q_log = m_log - LOGR, q = 1u << q_log;
  const u32 per_c_log = lg - LOGR;
  const u32 total = 1u << (per_c_log + cnt_log);
  u32 id = threadIdx.x;
  u32 c = 0, e_lo = 0, base = 0;
  u64 x[R];
  auto fetch = [&]() {
    u32 rest;
    if (INTERLEAVED) {
      c = id & ((1u << cnt_log) - 1);
      rest = id >> cnt_log;
    } else {
      rest = id & ((1u << per_c_log) - 1);
      c = id >> per_c_log;
    }
    e_lo = rest & (q - 1);
    base = ((rest >> q_log) << m_log) + e_lo;
#pragma unroll
    for (int i = 0; i < R; i++) x[i] = ld(c, base + ((u32)i << q_log));
  };
  bool has = id < total;
  if (has) fetch();
  pre();
  while (has) {
    dft_regs<LOGR, INV>(x);
#pragma unroll
    for (int p = 0; p < R; p++) {
      const u32 k1 = __brev((u32)p) >> (32 - LOGR);
      u64 v = x[p];
      if (p != 0 && q > 1) v = gl_mul(v, tw(e_lo, k1));
      st(c, base + ((u32)p << q_log), v);
    }
    id += blockDim.x;
    has = id < total;
    if (has) fetch();
  }
  __syncthreads();
}

// Full tile DIF of `cnt` transforms of length 2^lg: first stage reads through `ld0` (after which
// `pre` runs once), stages in between use the shared-memory tile, the last stage writes through `stN`.
// Afterwards position p of every transform holds X[rev_lg(p)].
// Twiddles: tw1 == nullptr: tw[e] = w_len^e, e < len, in shared memory serves every stage.
//           tw1 != nullptr (lg > 4): the first radix-16 stage reads the global matrix tw1[16][len/16] and
//           tw[e] = w_(len/16)^e, e < len/16, serves the rest - a 2^14-point tile then needs 8 KB of
//           twiddles next to its 136 KB of data instead of another 128 KB.
template <bool INTERLEAVED, bool INV, class LD0, class STN, class PRE>
GL_DEV void tile_dif(u64* sm, const u64* tw, const u64* tw1, int lg, u32 cnt_log, u32 pitch, LD0 ld0, STN stN,
                     PRE pre) {
  TileLd<INTERLEAVED> tl{sm, cnt_log, pitch};
  TileSt<INTERLEAVED> ts{sm, cnt_log, pitch};
  if (lg == 0) {  // nothing to transform: copy through
    pre();
    for (u32 c = threadIdx.x; c < (1u << cnt_log); c += blockDim.x) stN(c, 0, ld0(c, 0));
    __syncthreads();
    return;
  }
  const int sub = tw1 ? 4 : 0;
#define QPZK_STAGE(LOGR, LD, ST, PRE_) \
  dif_stage<LOGR, INTERLEAVED, INV>(TwTable{tw, lg - m_log - sub}, lg, m_log, cnt_log, LD, ST, PRE_)
  int m_log = lg;
  if (lg <= 4) {  // single stage: global in, `stN` out
    if (lg == 4) QPZK_STAGE(4, ld0, stN, pre);
    if (lg == 3) QPZK_STAGE(3, ld0, stN, pre);
    if (lg == 2) QPZK_STAGE(2, ld0, stN, pre);
    if (lg == 1) QPZK_STAGE(1, ld0, stN, pre);
    return;
  }
  if (tw1)
    dif_stage<4, INTERLEAVED, INV>(TwMatrix{tw1, (u32)(lg - 4)}, lg, m_log, cnt_log, ld0, ts, pre);
  else
    QPZK_STAGE(4, ld0, ts, pre);
  m_log -= 4;
  while (m_log > 4) {
    QPZK_STAGE(4, tl, ts, NoPre());
    m_log -= 4;
  }
  if (m_log == 4) QPZK_STAGE(4, tl, stN, NoPre());
  if (m_log == 3) QPZK_STAGE(3, tl, stN, NoPre());
  if (m_log == 2) QPZK_STAGE(2, tl, stN, NoPre());
  if (m_log == 1) QPZK_STAGE(1, tl, stN, NoPre());
#undef QPZK_STAGE
}

// Twiddle-table fill used as the `pre` step of a kernel's first stage.
struct TwFill {
  u64* tw; RootTab tab; u32 len; int shift;
  GL_DEV void operator()() const {
    for (u32 e = threadIdx.x; e < len; e += blockDim.x) tw[e] = root_pow(tab, (u64)e << shift);
    __syncthreads();
  }
};

// ---- single-CTA transform for n <= 2^14 ----
// grid (ncols, ncosets). src column c at src + c*src_stride (natural order).
// Output column c, coset t at dst + c*dst_stride + brev(t, r)*n:
//   NATURAL_OUT = false : DIF order (position = bit-reversed index)  [LDE flavour]
//   NATURAL_OUT = true  : natural order                              [IFFT flavour]
// tw1 == nullptr: n <= 2^12, the whole twiddle table sits in shared memory. Otherwise (n = 2^13, 2^14; the
// wormhole proof sizes) the first stage takes its twiddles from the global matrix tw1[16][n/16] and the
// tile (68 / 136 KB) plus n/16 twiddles fill the SM's shared memory: one read of the coefficients, one
// write of the evaluations, nothing in between touches HBM.
struct SmallLd {  // natural-order input column, optional coset pre-multiplier
  const u64* s; const u64* pmt;
  GL_DEV u64 operator()(u32, u32 j) const {
    u64 v = s[j];
    if (pmt) v = gl_mul(v, pmt[j]);
    return v;
  }
};
template <bool NATURAL_OUT, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_ntt_small(const u64* __restrict__ src, u64 src_stride, u64* __restrict__ dst, u64 dst_stride,
            const u64* __restrict__ pm, RootTab tab, const u64* __restrict__ tw1, int k, int r, u64 scale,
            u32 blk0) {
  extern __shared__ u64 smem[];
  const u32 n = 1u << k;
  const u32 pitch = tile_pitch(n);
  u64* x = smem;
  u64* tw = smem + pitch;
  // leaf block blk (n consecutive bit-reversed leaves) holds coset t = rev_r(blk)
  const u32 col = blockIdx.x, blk = blk0 + blockIdx.y, t = brev(blk, r);
  SmallLd ld{src + (u64)col * src_stride, pm ? pm + ((u64)t << k) : nullptr};
  // the natural-order flavour is the inverse transform
  tile_dif<false, NATURAL_OUT>(x, tw, tw1, k, 0, pitch, ld, TileSt<false>{x, 0, pitch},
                               TwFill{tw, tab, tw1 ? n >> 4 : n, tw1 ? 4 : 0});
  u64* d = dst + (u64)col * dst_stride + ((u64)blk << k);
  for (u32 q = threadIdx.x; q < n; q += blockDim.x) {
    u64 v = x[tile_pos<false>(0, NATURAL_OUT ? brev(q, k) : q, 0, pitch)];
    if (scale != 1) v = gl_mul(v, scale);
    d[q] = gl_canon(v);
  }
}

// ---- one pass over HBM for 2^15 and 2^16 points: a thread-block CLUSTER holds the (column, coset) tile ----
// Above 2^14 points a tile no longer fits one SM's shared memory and the two-pass kernels below write the
// half-transformed data to HBM and read it back (DRAM traffic 2.9x the algorithmic bytes at 2^16 x 135,
// profiles/r1_ntt_v2_ncu_full.txt). Eight CTAs of a cluster have 8 x 227 KB between them: with n = R * B,
// R = 2^A1 (8 or 16),
//   phase 1  CTA c takes the columns j2 in [c*B/8, (c+1)*B/8) of the [R][B] view: R coalesced loads per thread
//            (coset pre-multiplier applied), an R-point DFT in registers (w_16 = 2^12: shifts only), the
//            twiddle w_n^(j2*k1), and each result goes straight into the shared memory of the CTA that owns its
//            row - the row at DIF position p = rev(k1) lives in CTA p / (R/8) - as a distributed-shared-memory
//            store;
//   phase 2  after one cluster barrier every CTA holds R/8 complete rows: B-point DIFs in its own shared memory
//            (three radix-16 stages at B = 2^12), written to the contiguous block [p*B, (p+1)*B) of the
//            bit-reversed output.
// The coefficients are read once and the evaluations written once. Three CTAs per SM (launch bounds) so that
// the CTAs of other clusters run while one waits at its barrier.
#define QPZK_NTT_CLUSTER 8
template <int A1, int THREADS>
__global__ void __cluster_dims__(QPZK_NTT_CLUSTER, 1, 1) __launch_bounds__(THREADS, 3)
k_ntt_cluster(const u64* __restrict__ src, u64 src_stride, u64* __restrict__ dst, u64 dst_stride,
              const u64* __restrict__ pm, RootTab tabB, const u64* __restrict__ tw1, const u64* __restrict__ twc,
              int k, int r, u32 blk0) {
  constexpr int R = 1 << A1;
  constexpr int RPC = R / QPZK_NTT_CLUSTER;           // rows per CTA
  constexpr u32 RPC_LOG = A1 - 3;
  extern __shared__ u64 smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int lb = k - A1;                               // log2 B
  const u32 B = 1u << lb;
  const u32 pitch = tile_pitch(B);
  u64* x = smem;                                       // [RPC][pitch]
  u64* tw = smem + RPC * pitch;
  const u32 rank = cluster.block_rank();
  const u32 col = blockIdx.z, blk = blk0 + blockIdx.y, t = brev(blk, r);
  // twiddles of the local transform, then make sure every CTA of the cluster is running before its shared
  // memory is written from outside
  TwFill{tw, tabB, tw1 ? B >> 4 : B, tw1 ? 4 : 0}();
  cluster.sync();
  {
    const u64* s = src + (u64)col * src_stride;
    const u64* pmt = pm ? pm + ((u64)t << k) : nullptr;
    u64* remote[QPZK_NTT_CLUSTER];
#pragma unroll
    for (int c = 0; c < QPZK_NTT_CLUSTER; c++) remote[c] = cluster.map_shared_rank(x, c);
    const u32 per = B >> 3;  // columns j2 of this CTA
    for (u32 i = threadIdx.x; i < per; i += THREADS) {
      const u32 j2 = rank * per + i;
      u64 v[R];
#pragma unroll
      for (int j1 = 0; j1 < R; j1++) v[j1] = s[((u64)j1 << lb) + j2];
      if (pmt) {
#pragma unroll
        for (int j1 = 0; j1 < R; j1++) v[j1] = gl_mul(v[j1], pmt[((u64)j1 << lb) + j2]);
      }
      dft_regs<A1, false>(v);
      const u32 pos = j2 + (j2 >> 4);
#pragma unroll
      for (int p = 0; p < R; p++) {
        const u32 k1 = __brev((u32)p) >> (32 - A1);
        u64 e = v[p];
        if (p != 0) e = gl_mul(e, twc[((u64)k1 << lb) + j2]);
        QPZK_CHECK((p & (RPC - 1)) * pitch + pos < RPC * pitch);
        remote[p >> RPC_LOG][(p & (RPC - 1)) * pitch + pos] = e;
      }
    }
  }
  cluster.sync();
  tile_dif<false, false>(x, tw, tw1, lb, RPC_LOG, pitch, TileLd<false>{x, RPC_LOG, pitch}, TileSt<false>{x, RPC_LOG, pitch},
                         NoPre());
  u64* d = dst + (u64)col * dst_stride + ((u64)blk << k) + ((u64)(rank * RPC) << lb);
  for (u32 q = threadIdx.x; q < RPC * B; q += THREADS)
    d[q] = gl_canon(x[tile_pos<false>(q >> lb, q & (B - 1), 0, pitch)]);
}

// ---- pass A: strided n1-point DIF + inter-pass twiddle ----
// View column as [n1 = 2^a][n2 = 2^b]. grid (n2/COLS, ncosets, ncols); tile [n1][COLS]. The coset index runs
// faster than the column index so that the 2^r CTAs reading the same coefficient tile are scheduled
// together and all but the first find it in L2 (with columns running faster the 70.8 MB of coefficients of
// the 2^16 x 135 commit were streamed from HBM once per coset: 560 MB of DRAM reads, ncu).
//   ROW_BITREV = true  : result for frequency k1 is stored at row rev_a(k1)   [LDE flavour]
//   ROW_BITREV = false : stored at row k1                                     [IFFT flavour]
struct PassALd {  // element j1 of column j2_base + c of the [n1][n2] view, optional coset pre-multiplier
  const u64* s; const u64* pmt; int b; u32 j2_base;
  GL_DEV u64 operator()(u32 c, u32 j1) const {
    u64 j = ((u64)j1 << b) + j2_base + c;
    u64 v = s[j];
    if (pmt) v = gl_mul(v, pmt[j]);
    return v;
  }
};
template <bool ROW_BITREV>
struct PassASt {  // DIF position p = frequency k1 = rev_a(p): inter-pass twiddle w_n^(j2*k1), row p or k1
  u64* d; const u64* twm; int a, b; u32 j2_base;  // twm[k1][j2] = w_n^(j2*k1) (k_build_tw_matrix)
  GL_DEV void operator()(u32 c, u32 p, u64 v) const {
    u32 k1 = brev(p, a), j2 = j2_base + c;
    v = gl_mul(v, twm[((u64)k1 << b) + j2]);
    d[((u64)(ROW_BITREV ? p : k1) << b) + j2] = v;  // non-canonical is fine: pass B canonicalises
  }
};
template <bool ROW_BITREV>
__global__ void __launch_bounds__(256)
k_ntt_pass_a(const u64* __restrict__ src, u64 src_stride, u64* __restrict__ dst, u64 dst_stride,
             const u64* __restrict__ pm, RootTab tab, const u64* __restrict__ twm, int k, int a, int r,
             u32 cols_log, u32 blk0) {
  const u32 cols = 1u << cols_log;
  extern __shared__ u64 smem[];
  const int b = k - a;
  const u32 n1 = 1u << a;
  u64* x = smem;               // [n1][cols]
  u64* tw = smem + n1 * cols;  // [n1]
  const u32 col = blockIdx.z, blk = blk0 + blockIdx.y, t = brev(blk, r);
  const u32 j2_base = blockIdx.x * cols;
  PassALd ld{src + (u64)col * src_stride, pm ? pm + ((u64)t << k) : nullptr, b, j2_base};
  PassASt<ROW_BITREV> st{dst + (u64)col * dst_stride + ((u64)blk << k), twm, a, b, j2_base};
  // ROW_BITREV = forward (LDE), otherwise inverse
  tile_dif<true, !ROW_BITREV>(x, tw, nullptr, a, cols_log, 0, ld, st, TwFill{tw, tab, n1, b});
}

// ---- pass B (LDE flavour): n2-point DIF along contiguous rows, in place, DIF output order ----
// grid (n1/rows_per_cta, ncols, ncosets)
struct RowsLd {
  const u64* d; int b;
  GL_DEV u64 operator()(u32 row, u32 e) const { return d[((u64)row << b) + e]; }
};
__global__ void __launch_bounds__(256)
k_ntt_pass_b_rows(u64* __restrict__ data, u64 stride, RootTab tab, int k, int a, int r,
                  u32 rows_log, u32 blk0) {
  extern __shared__ u64 smem[];
  const u32 rows_per_cta = 1u << rows_log;
  const int b = k - a;
  const u32 n2 = 1u << b, pitch = tile_pitch(n2);
  u64* x = smem;                         // [rows][pitch]
  u64* tw = smem + rows_per_cta * pitch; // [n2]
  const u32 col = blockIdx.y, blk = blk0 + blockIdx.z;
  u64* d = data + (u64)col * stride + ((u64)blk << k) + (u64)blockIdx.x * rows_per_cta * n2;
  const u32 total = rows_per_cta * n2;
  tile_dif<false, false>(x, tw, nullptr, b, rows_log, pitch, RowsLd{d, b}, TileSt<false>{x, rows_log, pitch},
                         TwFill{tw, tab, n2, a});
  for (u32 idx = threadIdx.x; idx < total; idx += blockDim.x)
    d[idx] = gl_canon(x[tile_pos<false>(idx >> b, idx & (n2 - 1), 0, pitch)]);
}

// ---- pass B (IFFT flavour): n2-point DIF along rows k1, natural-order output X[k1 + n1*k2] ----
// grid (n1/rc, ncols). Tile = rc consecutive rows, padded by one element per row.
__global__ void __launch_bounds__(256)
k_ntt_pass_b_transpose(const u64* __restrict__ tmp, u64 tmp_stride, u64* __restrict__ dst,
                       u64 dst_stride, RootTab tab, int k, int a, u32 rc_log, u64 scale) {
  extern __shared__ u64 smem[];
  const u32 rc = 1u << rc_log;
  const int b = k - a;
  const u32 n2 = 1u << b, pitch = tile_pitch(n2);
  u64* x = smem;               // [rc][pitch]
  u64* tw = smem + rc * pitch; // [n2]
  const u32 col = blockIdx.y;
  const u32 k1_base = blockIdx.x * rc;
  const u64* s = tmp + (u64)col * tmp_stride + ((u64)k1_base << b);
  tile_dif<false, true>(x, tw, nullptr, b, rc_log, pitch, RowsLd{s, b}, TileSt<false>{x, rc_log, pitch},
                        TwFill{tw, tab, n2, a});
  u64* d = dst + (u64)col * dst_stride;
  for (u32 idx = threadIdx.x; idx < rc * n2; idx += blockDim.x) {
    u32 k2 = idx >> rc_log, rr = idx & (rc - 1);
    u64 v = gl_mul(x[tile_pos<false>(rr, brev(k2, b), 0, pitch)], scale);
    d[((u64)k2 << a) + k1_base + rr] = gl_canon(v);
  }
}

// Salt columns arrive in natural LDE order ([ncols][N]); leaves are stored bit-reversed. Leaves
// [leaf0, leaf0 + count) are written, column c at dst + c * dst_stride + leaf (a multi-GPU shard passes its
// own range and a destination shifted back by leaf0, see qpzk_batch::lde).
__global__ void k_bitrev_rows(const u64* __restrict__ src, u64* __restrict__ dst, u64 dst_stride, u64 leaf0, u64 count,
                              int log_n, u32 ncols) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  u64 N = (u64)1 << log_n;
  if (i >= count) return;
  i += leaf0;
  u64 j = log_n ? __brevll(i) >> (64 - log_n) : 0;
  for (u32 c = blockIdx.y; c < ncols; c += gridDim.y) dst[(u64)c * dst_stride + i] = gl_canon(src[(u64)c * N + j]);
}

// Column-major [width][N] -> row-major [N][width] (export of `merkle_tree.leaves`).
__global__ void k_transpose_to_rows(const u64* __restrict__ src, u64* __restrict__ dst, u64 N,
                                    u32 width) {
  __shared__ u64 tile[32][33];
  u64 r0 = (u64)blockIdx.x * 32;
  u32 c0 = blockIdx.y * 32;
  for (u32 cc = threadIdx.y; cc < 32; cc += blockDim.y) {
    u64 r = r0 + threadIdx.x;
    if (c0 + cc < width && r < N) tile[cc][threadIdx.x] = src[(u64)(c0 + cc) * N + r];
  }
  __syncthreads();
  for (u32 rr = threadIdx.y; rr < 32; rr += blockDim.y) {
    u32 c = c0 + threadIdx.x;
    if (c < width && r0 + rr < N) dst[(r0 + rr) * width + c] = tile[threadIdx.x][rr];
  }
}

// get_lde_values for a list of indices: out[i][c] = lde[c][rev(idx[i]*step)]
__global__ void k_gather_rows(const u64* __restrict__ lde, u64 stride, int log_n, const u32* __restrict__ idx,
                              u32 nidx, u32 step, u32 ncols, u64* __restrict__ out) {
  u32 i = blockIdx.x;
  if (i >= nidx) return;
  u64 nat = (u64)idx[i] * step;
  u64 leaf = log_n ? __brevll(nat) >> (64 - log_n) : 0;
  for (u32 c = threadIdx.x; c < ncols; c += blockDim.x) out[(u64)i * ncols + c] = lde[(u64)c * stride + leaf];
}

// One salted leaf row by leaf index (already bit-reversed position).
__global__ void k_gather_leaf(const u64* __restrict__ lde, u64 stride, u64 leaf, u32 width,
                              u64* __restrict__ out) {
  for (u32 c = threadIdx.x; c < width; c += blockDim.x) out[c] = lde[(u64)c * stride + leaf];
}

}  // namespace qpzk
