// Host-only check (compiled with nvcc, runs without a GPU) of the table derivations in
// qp-zk-circuits-rm_b200/csrc/poseidon_tables.hpp: the sparse partial-round form for every split
// "D dense partial rounds + 22 - D sparse ones" must compute the textbook Poseidon permutation, and the
// constants the FP64 MDS layers fold in must be the next round's constants. Not product code.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../qp-zk-circuits-rm_b200/csrc/gl.cuh"
#include "../../qp-zk-circuits-rm_b200/csrc/poseidon_tables.hpp"

using namespace qpzk;

static u64 sbox(u64 x) {
  u64 x2 = glh::mul(x, x), x4 = glh::mul(x2, x2), x3 = glh::mul(x, x2);
  return glh::mul(x3, x4);
}
static void mds(u64* s) {
  u64 o[12];
  for (int r = 0; r < 12; r++) {
    u64 acc = r == 0 ? glh::mul(kMdsDiag0, s[0]) : 0;
    for (int i = 0; i < 12; i++) acc = glh::add(acc, glh::mul(kMdsCirc[i], s[(i + r) % 12]));
    o[r] = acc;
  }
  memcpy(s, o, sizeof o);
}
static void naive(const PoseidonTablesHost& T, u64* s) {
  for (int r = 0; r < 30; r++) {
    for (int i = 0; i < 12; i++) s[i] = glh::add(s[i], T.rc[12 * r + i]);
    bool full = r < 4 || r >= 26;
    for (int i = 0; i < (full ? 12 : 1); i++) s[i] = sbox(s[i]);
    mds(s);
  }
}
// the structure of poseidon_permute() in poseidon.cuh, with the folded constants taken from the double table
static void hybrid(const PoseidonTablesHost& T, const double (*next)[2][12], u64* s) {
  const int D = T.dense, R = 22 - D;
  auto folded = [&](int L, int i) { return (u64)next[L][0][i] | ((u64)next[L][1][i] << 32); };
  for (int half = 0; half < 2; half++) {
    for (int i = 0; i < 12; i++) s[i] = glh::add(s[i], T.rc[12 * 26 * half + i]);
    int nrounds = half ? 4 : 4 + D, layer0 = half ? 4 + D : 0;
    for (int r = 0; r < nrounds; r++) {
      for (int i = 0; i < (r < 4 ? 12 : 1); i++) s[i] = sbox(s[i]);
      mds(s);
      for (int i = 0; i < 12; i++) s[i] = glh::add(s[i], folded(layer0 + r, i) % GL_P);
    }
    if (half) break;
    u64 o[12];
    o[0] = s[0];
    for (int c = 1; c < 12; c++) {
      u64 acc = 0;
      for (int r = 1; r < 12; r++) acc = glh::add(acc, glh::mul(s[r], T.h_init[(r - 1) * 11 + (c - 1)]));
      o[c] = acc;
    }
    memcpy(s, o, sizeof o);
    for (int r = 0; r < R; r++) {
      u64 s0 = glh::add(sbox(s[0]), T.h_rc[r]);
      u64 d = glh::mul(s0, 25);
      for (int i = 1; i < 12; i++) d = glh::add(d, glh::mul(s[i], T.h_w_hat[r * 11 + i - 1]));
      for (int i = 1; i < 12; i++) s[i] = glh::add(s[i], glh::mul(s0, T.h_v[r * 11 + i - 1]));
      s[0] = d;
    }
  }
}

// the structure of poseidon_permute_coop() in poseidon.cuh: four full rounds, the 22 partial rounds as the linear
// recurrence of build_linear_tables (two accumulators per lane, x_1 from the small MDS row), four full rounds
static void linearised(const PoseidonTablesHost& T, u64* s) {
  for (int i = 0; i < 12; i++) s[i] = glh::add(s[i], T.rc[i]);
  for (int r = 0; r < 4; r++) {
    for (int i = 0; i < 12; i++) s[i] = sbox(s[i]);
    mds(s);
    for (int i = 0; i < 12; i++) s[i] = glh::add(s[i], T.rc[12 * (r + 1) + i]);
  }
  u64 acc[2][16];
  for (int slot = 0; slot < 2; slot++)
    for (int lane = 0; lane < 16; lane++) {
      u64 a = T.lin_c[slot][lane];
      for (int i = 1; i < 12; i++) a = glh::add(a, glh::mul(T.lin_p[slot][i - 1][lane], s[i]));
      acc[slot][lane] = a;
    }
  u64 x1 = T.rc[60];
  for (int i = 1; i < 12; i++) x1 = glh::add(x1, glh::mul(kMdsCirc[i], s[i]));
  u64 x = s[0];
  for (int k = 0; k < 22; k++) {
    u64 y = sbox(x);
    for (int slot = 0; slot < 2; slot++)
      for (int lane = 0; lane < 16; lane++)
        acc[slot][lane] = glh::add(acc[slot][lane], glh::mul(T.lin_coef[k][slot][lane], y));
    if (k == 0) x = glh::add(x1, glh::mul(kMdsCirc[0] + kMdsDiag0, y));
    else if (k < 21) x = k <= 16 ? acc[0][k - 1] : acc[1][k - 5];
  }
  for (int i = 0; i < 12; i++) s[i] = acc[1][i];
  for (int r = 26; r < 30; r++) {
    for (int i = 0; i < 12; i++) s[i] = sbox(s[i]);
    mds(s);
    if (r < 29)
      for (int i = 0; i < 12; i++) s[i] = glh::add(s[i], T.rc[12 * (r + 1) + i]);
  }
}

int main() {
  static double next[QPZK_MDS_LAYERS_MAX][2][12], split[QPZK_MDS_LAYERS_MAX][2][12];
  u64 x = 0x9E3779B97F4A7C15ULL;
  for (int D = 0; D <= 21; D++) {
    PoseidonTablesHost* T = new PoseidonTablesHost();
    build_poseidon_tables(T, D);
    if (D == 0 && (memcmp(T->h_init, T->fast_init, sizeof T->h_init) || memcmp(T->h_v, T->fast_v, sizeof T->h_v) ||
                   memcmp(T->h_w_hat, T->fast_w_hat, sizeof T->h_w_hat) || memcmp(T->h_rc, T->fast_rc, sizeof T->h_rc))) {
      printf("FAIL: D = 0 tables differ from the all-sparse tables\n");
      return 1;
    }
    poseidon_next_rc_f64(*T, next, false);
    poseidon_next_rc_f64(*T, split, true);
    for (int L = 0; L < QPZK_MDS_LAYERS_MAX; L++)  // the split layout must recombine to the plain one
      for (int k = 0; k < 2; k++)
        for (int r = 0; r < 6; r++)
          if (split[L][k][r] + split[L][k][r + 6] != next[L][k][r] || split[L][k][r] - split[L][k][r + 6] != next[L][k][r + 6]) {
            printf("FAIL: split constants, D=%d layer %d\n", D, L);
            return 1;
          }
    for (int t = 0; t < 16; t++) {
      u64 a[12], b[12];
      for (int i = 0; i < 12; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        a[i] = b[i] = t == 0 ? 0 : x % GL_P;
      }
      naive(*T, a);
      if (D == 0) {
        u64 c[12];
        memcpy(c, b, sizeof c);
        linearised(*T, c);
        if (memcmp(a, c, sizeof a)) {
          printf("FAIL: the linearised partial rounds differ from the textbook permutation\n");
          return 1;
        }
      }
      hybrid(*T, next, b);
      if (t == 0 && D == 0 && (a[0] != 0x3c18a9786cb0b359ULL || a[11] != 0x1792b1c4342109d7ULL)) {
        printf("FAIL: permutation KAT\n");
        return 1;
      }
      if (memcmp(a, b, sizeof a)) {
        printf("FAIL: D=%d differs from the textbook permutation\n", D);
        return 1;
      }
    }
    delete T;
  }
  double h[12];
  poseidon_mds_half_f64(h);
  for (int i = 0; i < 12; i++)
    if (h[i] != (double)(long long)h[i]) {
      printf("FAIL: halved MDS constant %d is not an integer\n", i);
      return 1;
    }
  printf("ok\n");
  return 0;
}
