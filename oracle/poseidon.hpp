// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// Poseidon over Goldilocks: width 12, rate 8, x^7, 4 + 22 + 4 rounds, as used by
// `PoseidonGoldilocksConfig` (aliased `C` at /root/reference/common/src/circuit.rs:10) and
// called directly at /root/reference/wormhole/circuit/src/nullifier.rs:64-65 and
// /root/reference/wormhole/circuit/src/unspendable_account.rs:54-56.
// The algorithm lives in the un-vendored crate qp-plonky2 1.1.1
// (/root/reference/Cargo.lock:489-490); restated from SURVEY.md App. A.2-A.4. Round constants are
// REGENERATED (ChaCha8 seeded with 0), not copied, and the fast partial-round tables are DERIVED
// from the MDS matrix; `poseidon_selfcheck()` pins both against published known answers.
#pragma once
#include <array>
#include <cstring>
#include <vector>

#include "gl.hpp"

namespace orc {

static const int SPONGE_WIDTH = 12, SPONGE_RATE = 8, HALF_N_FULL = 4, N_PARTIAL = 22;
static const int N_ROUNDS = 2 * HALF_N_FULL + N_PARTIAL;
static const u64 MDS_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static const u64 MDS_DIAG[12] = {8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

typedef std::array<u64, 12> State;

struct PoseidonTables {
  u64 rc[N_ROUNDS * 12];             // ALL_ROUND_CONSTANTS
  u64 fast_first[12];                // FAST_PARTIAL_FIRST_ROUND_CONSTANT
  u64 fast_rc[N_PARTIAL];            // FAST_PARTIAL_ROUND_CONSTANTS (last entry unused = 0)
  u64 fast_init[11][11];             // FAST_PARTIAL_ROUND_INITIAL_MATRIX: out[c] += in[r]*init[r-1][c-1]
  u64 fast_init_t[11][11];           // transpose of fast_init (one contiguous dot product per lane)
  u64 fast_w_hat[N_PARTIAL][11];     // d = M00*s0 + sum w_hat[r][i-1]*s_i
  u64 fast_v[N_PARTIAL][11];         // s_i += s0 * v[r][i-1]
};

// ---- ChaCha8Rng::seed_from_u64(0) + rand 0.8 gen_range(0..p)   ----
struct ChaCha8 {
  uint32_t key[8];
  uint64_t counter = 0;
  uint32_t buf[16];
  int idx = 16;
  explicit ChaCha8(uint64_t seed) {
    uint64_t st = seed;
    for (int i = 0; i < 8; i++) {  // rand_core seed_from_u64: PCG32 expansion
      st = st * 6364136223846793005ULL + 11634580027462260723ULL;
      uint32_t xs = (uint32_t)(((st >> 18) ^ st) >> 27);
      uint32_t rot = (uint32_t)(st >> 59);
      key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
    }
  }
  static inline uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
  void block() {
    uint32_t s[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
    for (int i = 0; i < 8; i++) s[4 + i] = key[i];
    s[12] = (uint32_t)counter;
    s[13] = (uint32_t)(counter >> 32);
    s[14] = s[15] = 0;
    uint32_t x[16];
    memcpy(x, s, sizeof x);
#define QR(a, b, c, d)                                   \
  x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);            \
  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);            \
  x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);             \
  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    for (int r = 0; r < 4; r++) {  // 8 rounds = 4 double rounds
      QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
      QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
    }
#undef QR
    for (int i = 0; i < 16; i++) buf[i] = x[i] + s[i];
    counter++;
    idx = 0;
  }
  uint32_t next_u32() {
    if (idx == 16) block();
    return buf[idx++];
  }
  uint64_t next_u64() {
    uint64_t lo = next_u32();
    uint64_t hi = next_u32();
    return lo | (hi << 32);
  }
  uint64_t gen_range_p() {  // UniformInt::sample_single(0, p)
    const u64 zone = P - 1;  // (p << lz(p)) - 1
    for (;;) {
      u128 m = (u128)next_u64() * P;
      if ((u64)m <= zone) return (u64)(m >> 64);
    }
  }
};

// ---- small dense linear algebra mod p (for the fast partial-round derivation) ----
typedef std::vector<std::vector<u64>> Mat;
static inline Mat mat_mul(const Mat& A, const Mat& B) {
  size_t n = A.size(), m = B[0].size(), k = B.size();
  Mat C(n, std::vector<u64>(m, 0));
  for (size_t i = 0; i < n; i++)
    for (size_t j = 0; j < m; j++) {
      u64 s = 0;
      for (size_t t = 0; t < k; t++) s = add(s, mul(A[i][t], B[t][j]));
      C[i][j] = s;
    }
  return C;
}
static inline Mat mat_inv(Mat A) {
  size_t n = A.size();
  Mat I(n, std::vector<u64>(n, 0));
  for (size_t i = 0; i < n; i++) I[i][i] = 1;
  for (size_t c = 0; c < n; c++) {
    size_t piv = c;
    while (A[piv][c] == 0) piv++;
    std::swap(A[piv], A[c]);
    std::swap(I[piv], I[c]);
    u64 iv = inv(A[c][c]);
    for (size_t j = 0; j < n; j++) {
      A[c][j] = mul(A[c][j], iv);
      I[c][j] = mul(I[c][j], iv);
    }
    for (size_t r = 0; r < n; r++)
      if (r != c && A[r][c]) {
        u64 f = A[r][c];
        for (size_t j = 0; j < n; j++) {
          A[r][j] = sub(A[r][j], mul(f, A[c][j]));
          I[r][j] = sub(I[r][j], mul(f, I[c][j]));
        }
      }
  }
  return I;
}

static inline Mat mds_matrix() {  // new = M * state (column vector), M[r][c] = CIRC[(c-r)%12] + diag
  Mat M(12, std::vector<u64>(12));
  for (int r = 0; r < 12; r++)
    for (int c = 0; c < 12; c++) M[r][c] = MDS_CIRC[(c - r + 12) % 12] + (r == c ? MDS_DIAG[r] : 0);
  return M;
}

static inline PoseidonTables* build_tables() {
  PoseidonTables* t = new PoseidonTables();
  ChaCha8 rng(0);
  for (int i = 0; i < N_ROUNDS * 12; i++) t->rc[i] = rng.gen_range_p();

  Mat M = mds_matrix();
  Mat Minv = mat_inv(M);
  // (i) equivalent constants: push lanes 1..11 of each partial-round constant one round earlier.
  std::vector<std::array<u64, 12>> c(N_ROUNDS);
  for (int r = 0; r < N_ROUNDS; r++)
    for (int i = 0; i < 12; i++) c[r][i] = t->rc[12 * r + i];
  const int first = HALF_N_FULL, last = HALF_N_FULL + N_PARTIAL - 1;  // 4 .. 25
  for (int i = last - 1; i >= first; i--) {
    std::array<u64, 12> cp{};
    for (int r = 0; r < 12; r++) {
      u64 s = 0;
      for (int k = 0; k < 12; k++) s = add(s, mul(Minv[r][k], c[i + 1][k]));
      cp[r] = s;
    }
    for (int k = 1; k < 12; k++) c[i][k] = add(c[i][k], cp[k]);
    c[i + 1].fill(0);
    c[i + 1][0] = cp[0];
  }
  for (int i = 0; i < 12; i++) t->fast_first[i] = c[first][i];
  for (int r = 0; r < N_PARTIAL; r++) t->fast_rc[r] = (r < N_PARTIAL - 1) ? c[first + 1 + r][0] : 0;

  // (ii) sparse factorisation, column-vector form. Round r applies Meff_r = D_{r+1} * M.
  // Meff = [[m00, row],[col, Mh]] = [[m00, row*Mh^-1],[col, I]] * diag(1, Mh): the right factor
  // D_r commutes with the s-box and is pushed into the previous round; the left factor is the
  // sparse matrix with first row (m00, w_hat) and first column (m00; v).
  Mat Meff = M;
  for (int r = N_PARTIAL - 1; r >= 0; r--) {
    Mat Mh(11, std::vector<u64>(11));
    for (int i = 0; i < 11; i++)
      for (int j = 0; j < 11; j++) Mh[i][j] = Meff[i + 1][j + 1];
    Mat MhInv = mat_inv(Mh);
    for (int j = 0; j < 11; j++) {
      u64 s = 0;
      for (int k = 0; k < 11; k++) s = add(s, mul(Meff[0][k + 1], MhInv[k][j]));
      t->fast_w_hat[r][j] = s;
      t->fast_v[r][j] = Meff[j + 1][0];
    }
    Mat D(12, std::vector<u64>(12, 0));
    D[0][0] = 1;
    for (int i = 0; i < 11; i++)
      for (int j = 0; j < 11; j++) D[i + 1][j + 1] = Mh[i][j];
    if (r > 0) {
      Meff = mat_mul(D, M);
    } else {
      // remaining diag(1, Mh) is applied before the first partial round:
      // out[c] = sum_r Mh[c-1][r-1] * in[r]  ==> init[r-1][c-1] = Mh[c-1][r-1]
      for (int i = 0; i < 11; i++)
        for (int j = 0; j < 11; j++) t->fast_init[i][j] = Mh[j][i];
    }
  }
  for (int r = 0; r < 11; r++)
    for (int c = 0; c < 11; c++) t->fast_init_t[c][r] = t->fast_init[r][c];
  return t;
}
// Built once at load time (before any thread can call in), so the hot path pays no init guard.
static const PoseidonTables* const G_TABLES = build_tables();
static inline const PoseidonTables& tables() { return *G_TABLES; }

static inline u64 sbox(u64 x) {
  u64 x2 = sqr(x), x4 = sqr(x2), x3 = mul(x, x2);
  return mul(x3, x4);
}
// MDS entries are < 2^6, so the 32-bit halves of the state can be accumulated separately in
// u64 without overflow and recombined once per output lane.
static inline void mds_layer(State& s) {
  u64 lo[24], hi[24];
  for (int i = 0; i < 12; i++) {
    lo[i] = lo[i + 12] = s[i] & EPS;
    hi[i] = hi[i + 12] = s[i] >> 32;
  }
  State o;
  for (int r = 0; r < 12; r++) {
    u64 al = 0, ah = 0;
    for (int i = 0; i < 12; i++) {
      al += lo[i + r] * MDS_CIRC[i];
      ah += hi[i + r] * MDS_CIRC[i];
    }
    al += lo[r] * MDS_DIAG[r];
    ah += hi[r] * MDS_DIAG[r];
    o[r] = reduce128((u128)al + ((u128)ah << 32));
  }
  s = o;
}
// sum_i a[i]*b[i] for n <= 16 terms with two reductions instead of n.
static inline u64 dot(const u64* a, const u64* b, int n) {
  u128 lo = 0, hi = 0;
  for (int i = 0; i < n; i++) {
    u128 p = (u128)a[i] * b[i];
    lo += (u64)p;
    hi += (u64)(p >> 64);
  }
  // lo + 2^64*hi, with 2^64 == EPS
  return add(reduce128(lo), reduce128((u128)reduce128(hi) * EPS));
}

// Textbook form: 30 rounds of (add constants, s-box, MDS).
static inline void poseidon_naive(State& s) {
  const PoseidonTables& T = tables();
  for (int r = 0; r < N_ROUNDS; r++) {
    for (int i = 0; i < 12; i++) s[i] = add(s[i], T.rc[12 * r + i]);
    if (r < HALF_N_FULL || r >= HALF_N_FULL + N_PARTIAL) {
      for (int i = 0; i < 12; i++) s[i] = sbox(s[i]);
    } else {
      s[0] = sbox(s[0]);
    }
    mds_layer(s);
  }
}

// Fast form (sparse partial rounds) - same function; this is the form the PoseidonGate's
// constraints walk through.
static inline void partial_first_constant_layer(State& s) {
  const PoseidonTables& T = tables();
  for (int i = 0; i < 12; i++) s[i] = add(s[i], T.fast_first[i]);
}
static inline void mds_partial_layer_init(State& s) {
  const PoseidonTables& T = tables();
  State o{};
  o[0] = s[0];
  for (int c = 1; c < 12; c++) o[c] = dot(&s[1], T.fast_init_t[c - 1], 11);
  s = o;
}
static inline void mds_partial_layer_fast(State& s, int r) {
  const PoseidonTables& T = tables();
  u64 d = add(mul(s[0], MDS_CIRC[0] + MDS_DIAG[0]), dot(&s[1], T.fast_w_hat[r], 11));
  u64 s0 = s[0];
  for (int i = 1; i < 12; i++) s[i] = add(s[i], mul(s0, T.fast_v[r][i - 1]));
  s[0] = d;
}
static inline void poseidon(State& s) {
  const PoseidonTables& T = tables();
  int round = 0;
  for (int r = 0; r < HALF_N_FULL; r++, round++) {
    for (int i = 0; i < 12; i++) s[i] = sbox(add(s[i], T.rc[12 * round + i]));
    mds_layer(s);
  }
  partial_first_constant_layer(s);
  mds_partial_layer_init(s);
  for (int r = 0; r < N_PARTIAL; r++) {
    s[0] = sbox(s[0]);
    if (r < N_PARTIAL - 1) s[0] = add(s[0], T.fast_rc[r]);
    mds_partial_layer_fast(s, r);
  }
  round += N_PARTIAL;
  for (int r = 0; r < HALF_N_FULL; r++, round++) {
    for (int i = 0; i < 12; i++) s[i] = sbox(add(s[i], T.rc[12 * round + i]));
    mds_layer(s);
  }
}

// ---- hashing ----
struct Hash {
  u64 e[4];
};
static inline bool operator==(const Hash& a, const Hash& b) {
  return a.e[0] == b.e[0] && a.e[1] == b.e[1] && a.e[2] == b.e[2] && a.e[3] == b.e[3];
}
static inline Hash hash_no_pad(const u64* x, size_t n) {
  State s{};
  for (size_t off = 0; off < n; off += SPONGE_RATE) {
    size_t len = n - off < (size_t)SPONGE_RATE ? n - off : SPONGE_RATE;
    for (size_t i = 0; i < len; i++) s[i] = x[off + i];  // overwrite mode
    poseidon(s);
  }
  return Hash{{s[0], s[1], s[2], s[3]}};
}
static inline Hash hash_or_noop(const u64* x, size_t n) {
  if (n <= 4) {
    Hash h{{0, 0, 0, 0}};
    for (size_t i = 0; i < n; i++) h.e[i] = x[i];
    return h;
  }
  return hash_no_pad(x, n);
}
static inline Hash two_to_one(const Hash& l, const Hash& r) {
  State s{};
  for (int i = 0; i < 4; i++) {
    s[i] = l.e[i];
    s[4 + i] = r.e[i];
  }
  poseidon(s);
  return Hash{{s[0], s[1], s[2], s[3]}};
}

// Known-answer self check. Returns 0 when everything matches.
static inline int poseidon_selfcheck() {
  const PoseidonTables& T = tables();
  const u64 rc0[4] = {0xb585f766f2144405ULL, 0x7746a55f43921ad7ULL, 0xb2fb0d31cee799b4ULL,
                      0x0f6760a4803427d7ULL};
  for (int i = 0; i < 4; i++)
    if (T.rc[i] != rc0[i]) return 1;
  if (T.rc[359] != 0xbc8dfb627fe558fcULL) return 2;
  const u64 kat0[12] = {0x3c18a9786cb0b359ULL, 0xc4055e3364a246c3ULL, 0x7953db0ab48808f4ULL,
                        0xc71603f33a1144caULL, 0xd7709673896996dcULL, 0x46a84e87642f44edULL,
                        0xd032648251ee0b3cULL, 0x1c687363b207df62ULL, 0xdf8565563e8045feULL,
                        0x40f5b37ff4254daeULL, 0xd070f637b431067cULL, 0x1792b1c4342109d7ULL};
  State s{};
  poseidon_naive(s);
  for (int i = 0; i < 12; i++)
    if (s[i] != kat0[i]) return 3;
  State f{};
  poseidon(f);
  if (f != s) return 4;
  State a, b;
  for (int i = 0; i < 12; i++) a[i] = b[i] = i;
  poseidon_naive(a);
  poseidon(b);
  if (a != b || a[0] != 0xd64e1e3efc5b8e9eULL || a[11] != 0x5c0a27fcb0e1459bULL) return 5;
  for (int i = 0; i < 12; i++) a[i] = b[i] = P - 1;
  poseidon_naive(a);
  poseidon(b);
  if (a != b || a[0] != 0xbe0085cfc57a8357ULL) return 6;
  if (T.fast_first[0] != 0x3cc3f892184df408ULL || T.fast_first[1] != 0xe993fd841e7e97f1ULL ||
      T.fast_first[2] != 0xf2831d3575f0f3afULL)
    return 7;
  if (T.fast_rc[0] != 0x74cb2e819ae421abULL || T.fast_rc[1] != 0xd2559d2370e7f663ULL ||
      T.fast_rc[2] != 0x62bf78acf843d17cULL)
    return 8;
  return 0;
}

}  // namespace orc
