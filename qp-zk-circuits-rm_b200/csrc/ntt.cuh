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
#include "gl.cuh"

namespace qpzk {

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

// In-shared-memory radix-2 DIF over `cnt` transforms of length 2^lg held in one tile.
// Element e of transform c lives at sm[e*estride + c*cstride]. tw[e] = w_len^e, e < len/2.
// C_FASTEST picks the thread->butterfly map so that a warp walks the unit-stride dimension:
//   true  : transforms are interleaved (cstride == 1), adjacent threads take adjacent transforms
//   false : each transform is contiguous (estride == 1), adjacent threads take adjacent elements
// After the call position p holds X[rev_lg(p)].
template <bool C_FASTEST>
GL_DEV void smem_dif(u64* sm, const u64* tw, int lg, u32 cnt_log, u32 estride, u32 cstride) {
  const u32 cnt = 1u << cnt_log;
  if (lg == 0) return;
  const u32 hcount = 1u << (lg - 1);
  const u32 half_total = hcount * cnt;
  for (int s = lg - 1; s >= 0; s--) {
    const u32 half = 1u << s;
    for (u32 b = threadIdx.x; b < half_total; b += blockDim.x) {
      u32 c, bb;
      if (C_FASTEST) {
        c = b & (cnt - 1);
        bb = b >> cnt_log;
      } else {
        c = b >> (lg - 1);
        bb = b & (hcount - 1);
      }
      u32 lowbits = bb & (half - 1);
      u32 i = ((bb >> s) << (s + 1)) | lowbits;
      u64* p0 = sm + (u64)i * estride + (u64)c * cstride;
      u64* p1 = p0 + (u64)half * estride;
      u64 u = *p0, v = *p1;
      *p0 = gl_add(u, v);
      u64 d = gl_sub(u, v);
      *p1 = s == 0 ? d : gl_mul(d, tw[lowbits << (lg - 1 - s)]);
    }
    __syncthreads();
  }
}

// ---- single-CTA transform for n <= 2^12 ----
// grid (ncols, ncosets). src column c at src + c*src_stride (natural order).
// Output column c, coset t at dst + c*dst_stride + brev(t, r)*n:
//   NATURAL_OUT = false : DIF order (position = bit-reversed index)  [LDE flavour]
//   NATURAL_OUT = true  : natural order                              [IFFT flavour]
template <bool NATURAL_OUT>
__global__ void __launch_bounds__(256)
k_ntt_small(const u64* __restrict__ src, u64 src_stride, u64* __restrict__ dst, u64 dst_stride,
            const u64* __restrict__ pm, RootTab tab, int k, int r, u64 scale, u32 blk0) {
  extern __shared__ u64 smem[];
  const u32 n = 1u << k;
  u64* x = smem;
  u64* tw = smem + n;
  // leaf block blk (n consecutive bit-reversed leaves) holds coset t = rev_r(blk)
  const u32 col = blockIdx.x, blk = blk0 + blockIdx.y, t = brev(blk, r);
  const u64* s = src + (u64)col * src_stride;
  const u64* pmt = pm ? pm + ((u64)t << k) : nullptr;
  for (u32 j = threadIdx.x; j < n; j += blockDim.x) {
    u64 v = s[j];
    if (pmt) v = gl_mul(v, pmt[j]);
    x[j] = v;
  }
  for (u32 e = threadIdx.x; e < n / 2; e += blockDim.x) tw[e] = root_pow(tab, e);
  __syncthreads();
  smem_dif<false>(x, tw, k, 0, 1, 0);
  u64* d = dst + (u64)col * dst_stride + ((u64)blk << k);
  for (u32 q = threadIdx.x; q < n; q += blockDim.x) {
    u64 v = NATURAL_OUT ? x[brev(q, k)] : x[q];
    if (scale != 1) v = gl_mul(v, scale);
    d[q] = gl_canon(v);
  }
}

// ---- pass A: strided n1-point DIF + inter-pass twiddle ----
// View column as [n1 = 2^a][n2 = 2^b]. grid (n2/COLS, ncols, ncosets); tile [n1][COLS].
//   ROW_BITREV = true  : result for frequency k1 is stored at row rev_a(k1)   [LDE flavour]
//   ROW_BITREV = false : stored at row k1                                     [IFFT flavour]
template <bool ROW_BITREV>
__global__ void __launch_bounds__(256)
k_ntt_pass_a(const u64* __restrict__ src, u64 src_stride, u64* __restrict__ dst, u64 dst_stride,
             const u64* __restrict__ pm, RootTab tab, int k, int a, int r, u32 cols_log, u32 blk0) {
  const u32 cols = 1u << cols_log;
  extern __shared__ u64 smem[];
  const int b = k - a;
  const u32 n1 = 1u << a;
  u64* x = smem;               // [n1][cols]
  u64* tw = smem + n1 * cols;  // [n1/2]
  const u32 col = blockIdx.y, blk = blk0 + blockIdx.z, t = brev(blk, r);
  const u32 j2_base = blockIdx.x * cols;
  const u64* s = src + (u64)col * src_stride;
  const u64* pmt = pm ? pm + ((u64)t << k) : nullptr;
  for (u32 idx = threadIdx.x; idx < n1 * cols; idx += blockDim.x) {
    u32 j1 = idx >> cols_log, c = idx & (cols - 1);
    u64 j = ((u64)j1 << b) + j2_base + c;
    u64 v = s[j];
    if (pmt) v = gl_mul(v, pmt[j]);
    x[idx] = v;
  }
  for (u32 e = threadIdx.x; e < n1 / 2; e += blockDim.x) tw[e] = root_pow(tab, (u64)e << b);
  __syncthreads();
  smem_dif<true>(x, tw, a, cols_log, cols, 1);
  u64* d = dst + (u64)col * dst_stride + ((u64)blk << k);
  for (u32 idx = threadIdx.x; idx < n1 * cols; idx += blockDim.x) {
    u32 p = idx >> cols_log, c = idx & (cols - 1);
    u32 k1 = brev(p, a);
    u32 j2 = j2_base + c;
    u64 v = gl_mul(x[idx], root_pow(tab, (u64)j2 * k1));
    u32 row = ROW_BITREV ? p : k1;
    d[((u64)row << b) + j2] = v;  // non-canonical is fine: pass B canonicalises
  }
}

// ---- pass B (LDE flavour): n2-point DIF along contiguous rows, in place, DIF output order ----
// grid (n1/rows_per_cta, ncols, ncosets)
__global__ void __launch_bounds__(256)
k_ntt_pass_b_rows(u64* __restrict__ data, u64 stride, RootTab tab, int k, int a, int r,
                  u32 rows_log, u32 blk0) {
  extern __shared__ u64 smem[];
  const u32 rows_per_cta = 1u << rows_log;
  const int b = k - a;
  const u32 n2 = 1u << b;
  u64* x = smem;                      // [rows][n2]
  u64* tw = smem + rows_per_cta * n2; // [n2/2]
  const u32 col = blockIdx.y, blk = blk0 + blockIdx.z;
  u64* d = data + (u64)col * stride + ((u64)blk << k) + (u64)blockIdx.x * rows_per_cta * n2;
  const u32 total = rows_per_cta * n2;
  for (u32 idx = threadIdx.x; idx < total; idx += blockDim.x) x[idx] = d[idx];
  for (u32 e = threadIdx.x; e < n2 / 2; e += blockDim.x) tw[e] = root_pow(tab, (u64)e << a);
  __syncthreads();
  smem_dif<false>(x, tw, b, rows_log, 1, n2);
  for (u32 idx = threadIdx.x; idx < total; idx += blockDim.x) d[idx] = gl_canon(x[idx]);
}

// ---- pass B (IFFT flavour): n2-point DIF along rows k1, natural-order output X[k1 + n1*k2] ----
// grid (n1/rc, ncols). Tile = rc consecutive rows, padded by one element per row.
__global__ void __launch_bounds__(256)
k_ntt_pass_b_transpose(const u64* __restrict__ tmp, u64 tmp_stride, u64* __restrict__ dst,
                       u64 dst_stride, RootTab tab, int k, int a, u32 rc_log, u64 scale) {
  extern __shared__ u64 smem[];
  const u32 rc = 1u << rc_log;
  const int b = k - a;
  const u32 n2 = 1u << b, pitch = n2 + 1;
  u64* x = smem;               // [rc][pitch]
  u64* tw = smem + rc * pitch; // [n2/2]
  const u32 col = blockIdx.y;
  const u32 k1_base = blockIdx.x * rc;
  const u64* s = tmp + (u64)col * tmp_stride + ((u64)k1_base << b);
  for (u32 idx = threadIdx.x; idx < rc * n2; idx += blockDim.x) {
    u32 rr = idx >> b, j2 = idx & (n2 - 1);
    x[rr * pitch + j2] = s[idx];
  }
  for (u32 e = threadIdx.x; e < n2 / 2; e += blockDim.x) tw[e] = root_pow(tab, (u64)e << a);
  __syncthreads();
  smem_dif<false>(x, tw, b, rc_log, 1, pitch);
  u64* d = dst + (u64)col * dst_stride;
  for (u32 idx = threadIdx.x; idx < rc * n2; idx += blockDim.x) {
    u32 k2 = idx >> rc_log, rr = idx & (rc - 1);
    u64 v = gl_mul(x[rr * pitch + brev(k2, b)], scale);
    d[((u64)k2 << a) + k1_base + rr] = gl_canon(v);
  }
}

// Salt columns arrive in natural LDE order; leaves are stored bit-reversed.
__global__ void k_bitrev_rows(const u64* __restrict__ src, u64* __restrict__ dst, int log_n,
                              u32 ncols) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  u64 N = (u64)1 << log_n;
  if (i >= N) return;
  u64 j = __brevll(i) >> (64 - log_n);
  for (u32 c = blockIdx.y; c < ncols; c += gridDim.y) dst[(u64)c * N + i] = gl_canon(src[(u64)c * N + j]);
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
__global__ void k_gather_rows(const u64* __restrict__ lde, int log_n, const u32* __restrict__ idx,
                              u32 nidx, u32 step, u32 ncols, u64* __restrict__ out) {
  u32 i = blockIdx.x;
  if (i >= nidx) return;
  u64 N = (u64)1 << log_n;
  u64 nat = (u64)idx[i] * step;
  u64 leaf = __brevll(nat) >> (64 - log_n);
  for (u32 c = threadIdx.x; c < ncols; c += blockDim.x) out[(u64)i * ncols + c] = lde[(u64)c * N + leaf];
}

// One salted leaf row by leaf index (already bit-reversed position).
__global__ void k_gather_leaf(const u64* __restrict__ lde, u64 N, u64 leaf, u32 width,
                              u64* __restrict__ out) {
  for (u32 c = threadIdx.x; c < width; c += blockDim.x) out[c] = lde[(u64)c * N + leaf];
}

}  // namespace qpzk
