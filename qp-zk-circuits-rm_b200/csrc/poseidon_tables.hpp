// Host-side generation of the Poseidon-Goldilocks constant tables uploaded to __constant__ memory.
//
// Replaces the constant arrays of `PoseidonGoldilocksConfig` (qp-plonky2 1.1.1, un-vendored;
// used through `C` at /root/reference/common/src/circuit.rs:10 and directly at
// /root/reference/wormhole/circuit/src/nullifier.rs:64-65). Nothing is copied: the 360 round
// constants are re-drawn from ChaCha8 keyed by seed 0 and the sparse ("fast") partial-round
// factorisation is derived from the circulant MDS matrix here, in the transposed (row-vector)
// formulation. Runs once per context; never on the data path.
#pragma once
#include <cstring>
#include <vector>

#include "gl.cuh"

namespace qpzk {

struct PoseidonTablesHost {
  u64 rc[360];
  u64 fast_first[12];
  u64 fast_rc[22];
  u64 fast_init[121];  // [r-1][c-1], out[c] += in[r] * init
  u64 fast_w_hat[242]; // [round][i-1]
  u64 fast_v[242];     // [round][i-1]
  // The same sparse tables for a permutation whose first `dense` partial rounds run in the textbook form
  // (constants, s-box on lane 0, dense MDS) and only the remaining 22 - dense in the sparse form.
  int dense;
  u64 h_first[12], h_rc[22], h_init[121], h_w_hat[242], h_v[242];
  // Linearised partial rounds for the 16-lane permutation (build_linear_tables): 32 accumulators, two per lane.
  u64 lin_p[2][11][16];    // [slot][i - 1][lane]: coefficient of t0[i], i = 1..11
  u64 lin_c[2][16];        // [slot][lane]: the constant term
  u64 lin_coef[22][2][16]; // [k][slot][lane]: coefficient of y_k
};

static const u64 kMdsCirc[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static const u64 kMdsDiag0 = 8;

namespace detail {

// Keystream of ChaCha with 8 rounds, 64-bit block counter, zero nonce.
class ChaCha8Stream {
 public:
  explicit ChaCha8Stream(u64 seed) {
    // rand_core::SeedableRng::seed_from_u64: PCG32 output function over an LCG
    u64 s = seed;
    for (int w = 0; w < 8; w++) {
      s = s * 6364136223846793005ULL + 11634580027462260723ULL;
      u32 x = (u32)(((s >> 18) ^ s) >> 27);
      u32 r = (u32)(s >> 59);
      key_[w] = (x >> r) | (x << ((32u - r) & 31u));
    }
  }
  u64 next64() {
    u64 lo = next32();
    u64 hi = next32();
    return lo | (hi << 32);
  }

 private:
  u32 key_[8];
  u64 block_ = 0;
  u32 out_[16];
  int pos_ = 16;
  static u32 rol(u32 v, int n) { return (v << n) | (v >> (32 - n)); }
  static void quarter(u32* x, int a, int b, int c, int d) {
    x[a] += x[b]; x[d] = rol(x[d] ^ x[a], 16);
    x[c] += x[d]; x[b] = rol(x[b] ^ x[c], 12);
    x[a] += x[b]; x[d] = rol(x[d] ^ x[a], 8);
    x[c] += x[d]; x[b] = rol(x[b] ^ x[c], 7);
  }
  u32 next32() {
    if (pos_ == 16) {
      u32 in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
      memcpy(in + 4, key_, sizeof key_);
      in[12] = (u32)block_;
      in[13] = (u32)(block_ >> 32);
      in[14] = in[15] = 0;
      u32 x[16];
      memcpy(x, in, sizeof x);
      for (int dr = 0; dr < 4; dr++) {
        quarter(x, 0, 4, 8, 12); quarter(x, 1, 5, 9, 13); quarter(x, 2, 6, 10, 14); quarter(x, 3, 7, 11, 15);
        quarter(x, 0, 5, 10, 15); quarter(x, 1, 6, 11, 12); quarter(x, 2, 7, 8, 13); quarter(x, 3, 4, 9, 14);
      }
      for (int i = 0; i < 16; i++) out_[i] = x[i] + in[i];
      block_++;
      pos_ = 0;
    }
    return out_[pos_++];
  }
};

// Flat row-major square matrices mod p.
typedef std::vector<u64> Mx;
static inline Mx mx_mul(const Mx& A, const Mx& B, int n) {
  Mx C((size_t)n * n, 0);
  for (int i = 0; i < n; i++)
    for (int k = 0; k < n; k++) {
      u64 a = A[i * n + k];
      if (!a) continue;
      for (int j = 0; j < n; j++) C[i * n + j] = glh::add(C[i * n + j], glh::mul(a, B[k * n + j]));
    }
  return C;
}
static inline Mx mx_inv(Mx A, int n) {  // Gauss-Jordan
  Mx R((size_t)n * n, 0);
  for (int i = 0; i < n; i++) R[i * n + i] = 1;
  for (int col = 0; col < n; col++) {
    int p = col;
    while (A[p * n + col] == 0) p++;
    for (int j = 0; j < n; j++) {
      std::swap(A[p * n + j], A[col * n + j]);
      std::swap(R[p * n + j], R[col * n + j]);
    }
    u64 s = glh::inv(A[col * n + col]);
    for (int j = 0; j < n; j++) {
      A[col * n + j] = glh::mul(A[col * n + j], s);
      R[col * n + j] = glh::mul(R[col * n + j], s);
    }
    for (int r = 0; r < n; r++) {
      u64 f = A[r * n + col];
      if (r == col || !f) continue;
      for (int j = 0; j < n; j++) {
        A[r * n + j] = glh::sub(A[r * n + j], glh::mul(f, A[col * n + j]));
        R[r * n + j] = glh::sub(R[r * n + j], glh::mul(f, R[col * n + j]));
      }
    }
  }
  return R;
}

}  // namespace detail

// Sparse form of the partial rounds 4 + D .. 25 (D = number of leading partial rounds kept dense):
// first[12] replaces the constants of round 4 + D, frc[r] is added to lane 0 after the s-box of sparse round r,
// init is the 11x11 matrix applied before the first sparse round, w_hat / v the per-round sparse factors.
static inline void build_sparse_tables(const u64* rc, const detail::Mx& Mt, const detail::Mx& Minv, int D, u64* first,
                                       u64* frc, u64* init, u64* w_hat, u64* vv) {
  using namespace detail;
  const int W = 12, R = 22 - D, p0 = 4 + D;
  // Equivalent round constants for the sparse rounds (rounds p0..25).
  u64 c[30][12];
  memcpy(c, rc, sizeof c);
  for (int i = 24; i >= p0; i--) {
    u64 t[12];
    for (int r = 0; r < W; r++) {
      u64 s = 0;
      for (int k = 0; k < W; k++) s = glh::add(s, glh::mul(Minv[r * W + k], c[i + 1][k]));
      t[r] = s;
    }
    for (int k = 1; k < W; k++) c[i][k] = glh::add(c[i][k], t[k]);
    memset(c[i + 1], 0, sizeof c[i + 1]);
    c[i + 1][0] = t[0];
  }
  memcpy(first, c[p0], 12 * sizeof(u64));
  for (int r = 0; r < 22; r++) frc[r] = r < R - 1 ? c[p0 + 1 + r][0] : 0;

  // Sparse factorisation, walking the rounds backwards.
  Mx Mmul = Mt;
  for (int i = R - 1; i >= 0; i--) {
    Mx Mhat(11 * 11), w(11), v(11);
    for (int a = 0; a < 11; a++) {
      for (int b = 0; b < 11; b++) Mhat[a * 11 + b] = Mmul[(a + 1) * W + (b + 1)];
      w[a] = Mmul[(a + 1) * W + 0];
      v[a] = Mmul[0 * W + (a + 1)];
    }
    Mx MhatInv = mx_inv(Mhat, 11);
    for (int a = 0; a < 11; a++) {
      u64 s = 0;
      for (int b = 0; b < 11; b++) s = glh::add(s, glh::mul(MhatInv[a * 11 + b], w[b]));
      w_hat[i * 11 + a] = s;
      vv[i * 11 + a] = v[a];
    }
    Mx Mi(W * W, 0);
    Mi[0] = 1;
    for (int a = 0; a < 11; a++)
      for (int b = 0; b < 11; b++) Mi[(a + 1) * W + (b + 1)] = Mhat[a * 11 + b];
    if (i > 0) {
      Mmul = mx_mul(Mt, Mi, W);
    } else {
      // state_row * Mi is applied before the first sparse round
      for (int a = 0; a < 11; a++)
        for (int b = 0; b < 11; b++) init[a * 11 + b] = Mhat[a * 11 + b];
    }
  }
}

// The 22 partial rounds as ONE linear recurrence driven by the 22 s-box outputs (poseidon_permute_coop).
// With t_k the state of partial round k after its constants (t_0 = the state entering round 4 plus rc[4]),
// x_k = t_k[0], y_k = x_k^7, A = the MDS matrix with its first column zeroed and m0 that first column,
//     t_(k+1) = A t_k + m0 y_k + rc[5 + k]             (k = 0..21; t_22 enters the s-boxes of round 26)
// so every x_m and every word of t_22 is an affine function of t_0[1..11] and of y_0..y_(m-1):
//     x_m  = row_0(A^m) t_0 + sum_(j<m) (row_0(A^(m-1-j)) m0) y_j + sum_(j<m) row_0(A^(m-1-j)) rc[5 + j]
//     t_22 = A^22 t_0       + sum_(j<22) (A^(21-j) m0) y_j      + sum_(j<22) A^(21-j) rc[5 + j]
// Column 0 of A^m is zero, so t_0[0] enters only through y_0. x_1 has small coefficients (row 0 of the MDS
// matrix) and is computed by every lane; the other 20 + 12 accumulators sit two per lane:
//   slot 0, lane L       : x_(L+2)            (x_2 .. x_17)
//   slot 1, lane L < 12  : t_22[L]
//   slot 1, lane L >= 12 : x_(L+6)            (x_18 .. x_21)
// A coefficient of y_k in an x_m with m <= k (already consumed) is zero.
static inline void build_linear_tables(PoseidonTablesHost* T, const detail::Mx& M) {
  using namespace detail;
  const int W = 12;
  Mx A = M;
  u64 m0[12];
  for (int r = 0; r < W; r++) {
    m0[r] = M[r * W];
    A[r * W] = 0;
  }
  std::vector<Mx> P(23);
  P[0] = Mx(W * W, 0);
  for (int i = 0; i < W; i++) P[0][i * W + i] = 1;
  for (int d = 1; d <= 22; d++) P[d] = mx_mul(A, P[d - 1], W);
  auto row_dot = [&](const Mx& X, int row, const u64* v) {
    u64 s = 0;
    for (int c = 0; c < W; c++) s = glh::add(s, glh::mul(X[row * W + c], v[c]));
    return s;
  };
  // accumulator `slot` of `lane` is row `row` of the state m rounds in (m = 22: the output)
  for (int slot = 0; slot < 2; slot++)
    for (int lane = 0; lane < 16; lane++) {
      int m, row;
      if (slot == 0) m = lane + 2, row = 0;
      else if (lane < 12) m = 22, row = lane;
      else m = lane + 6, row = 0;
      for (int i = 1; i < W; i++) T->lin_p[slot][i - 1][lane] = P[m][row * W + i];
      u64 c = 0;
      for (int j = 0; j < m; j++) c = glh::add(c, row_dot(P[m - 1 - j], row, T->rc + 12 * (5 + j)));
      T->lin_c[slot][lane] = c;
      for (int k = 0; k < 22; k++) T->lin_coef[k][slot][lane] = k < m ? row_dot(P[m - 1 - k], row, m0) : 0;
    }
}

static inline void build_poseidon_tables(PoseidonTablesHost* T, int dense = 0) {
  using namespace detail;
  typedef unsigned __int128 u128;
  // 360 x gen_range(0..p) (rand 0.8 widening-multiply rejection sampling; zone = p - 1)
  ChaCha8Stream rng(0);
  for (int i = 0; i < 360; i++) {
    for (;;) {
      u128 m = (u128)rng.next64() * GL_P;
      if ((u64)m <= GL_P - 1) {
        T->rc[i] = (u64)(m >> 64);
        break;
      }
    }
  }
  // Transposed MDS: row-vector convention, new_row = state_row * Mt, Mt[c][r] = circ[(c - r) mod 12].
  const int W = 12;
  Mx Mt(W * W);
  for (int r = 0; r < W; r++)
    for (int c = 0; c < W; c++)
      Mt[c * W + r] = kMdsCirc[((c - r) % W + W) % W] + ((r == c && r == 0) ? kMdsDiag0 : 0);
  Mx M(W * W);
  for (int r = 0; r < W; r++)
    for (int c = 0; c < W; c++) M[r * W + c] = Mt[c * W + r];
  Mx Minv = mx_inv(M, W);

  build_sparse_tables(T->rc, Mt, Minv, 0, T->fast_first, T->fast_rc, T->fast_init, T->fast_w_hat, T->fast_v);
  T->dense = dense;
  build_sparse_tables(T->rc, Mt, Minv, dense, T->h_first, T->h_rc, T->h_init, T->h_w_hat, T->h_v);
  build_linear_tables(T, M);
}

// Constants added by FP64 MDS layer L on behalf of the round that follows it, as the 32-bit halves of each word
// in doubles: [L][half][lane]. Layers 0..3 are the first four full rounds, 4..3+D the dense partial rounds,
// 4+D..7+D the last four full rounds. Every layer carries the next round's constants; the last layer before the
// sparse rounds carries their `first` constants; layer 3+D... the very last layer carries zeros. (The first
// round of each block of full rounds that follows sparse rounds adds its own constants.)
#define QPZK_MDS_LAYERS_MAX 30
static inline void poseidon_next_rc_f64(const PoseidonTablesHost& T, double (*out)[2][12], bool split) {
  const int D = T.dense;
  for (int L = 0; L < QPZK_MDS_LAYERS_MAX; L++)
    for (int i = 0; i < 12; i++) {
      u64 c = 0;
      if (L < 3 + D) c = T.rc[12 * (L + 1) + i];                  // next full / dense partial round
      else if (L == 3 + D) c = D == 22 ? T.rc[12 * 26 + i] : T.h_first[i];  // first sparse round (or round 26)
      else if (L < 7 + D) c = T.rc[12 * (26 + (L - 4 - D) + 1) + i];
      out[L][0][i] = (double)(uint32_t)c;
      out[L][1][i] = (double)(uint32_t)(c >> 32);
    }
  // the split MDS layer (poseidon.cuh) wants (rc[r] + rc[r+6]) / 2 in slot r and (rc[r] - rc[r+6]) / 2 in slot r+6
  if (split)
    for (int L = 0; L < QPZK_MDS_LAYERS_MAX; L++)
      for (int k = 0; k < 2; k++)
        for (int r = 0; r < 6; r++) {
          double a = out[L][k][r], b = out[L][k][r + 6];
          out[L][k][r] = 0.5 * (a + b);
          out[L][k][r + 6] = 0.5 * (a - b);
        }
}
// (c[i] + c[i+6]) / 2, i < 6, then (c[i] - c[i+6]) / 2: the constants of the split circulant product
static inline void poseidon_mds_half_f64(double* out) {
  for (int i = 0; i < 6; i++) {
    out[i] = 0.5 * ((double)kMdsCirc[i] + (double)kMdsCirc[i + 6]);
    out[6 + i] = 0.5 * ((double)kMdsCirc[i] - (double)kMdsCirc[i + 6]);
  }
}

}  // namespace qpzk
