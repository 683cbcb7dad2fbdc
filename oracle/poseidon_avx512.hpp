// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// Eight Poseidon permutations at a time in AVX-512 (one u64 lane per permutation), for the CPU baseline:
// qp-plonky2's prover hashes Merkle leaves with hand-vectorised Poseidon / packed-Goldilocks code selected by
// `target-feature`, so a scalar port would flatter the GPU arm. The batched form hashes
// 8 independent leaves (or 8 sibling pairs) per call and must produce exactly what `poseidon()` does.
// Compiled for AVX-512 regardless of the build flags and selected at run time (the oracle is built in one
// container and travels to another box).
#pragma once
#include <immintrin.h>

#include "poseidon.hpp"

namespace orc {

#define ORC_AVX512 __attribute__((target("avx512f,avx512dq,avx512vl"), always_inline)) static inline
#define ORC_AVX512_FN __attribute__((target("avx512f,avx512dq,avx512vl")))

typedef __m512i v8;

static inline bool have_avx512() {
  static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") &&
                         __builtin_cpu_supports("avx512vl");
  return ok;
}

ORC_AVX512 v8 v8_set1(u64 x) { return _mm512_set1_epi64((long long)x); }
// a + b for canonical a, b -> canonical
ORC_AVX512 v8 v8_add(v8 a, v8 b) {
  const v8 p = v8_set1(P);
  v8 s = _mm512_add_epi64(a, b);
  __mmask8 over = _mm512_cmplt_epu64_mask(s, a) | _mm512_cmpge_epu64_mask(s, p);
  return _mm512_mask_sub_epi64(s, over, s, p);
}
// (lo, hi) of a 128-bit value -> canonical; the reduction of gl.hpp reduce128, lane-wise
ORC_AVX512 v8 v8_reduce128(v8 lo, v8 hi) {
  const v8 eps = v8_set1(EPS), p = v8_set1(P);
  v8 hi_hi = _mm512_srli_epi64(hi, 32), hi_lo = _mm512_and_si512(hi, eps);
  v8 t0 = _mm512_sub_epi64(lo, hi_hi);
  __mmask8 borrow = _mm512_cmplt_epu64_mask(lo, hi_hi);
  t0 = _mm512_mask_sub_epi64(t0, borrow, t0, eps);
  v8 t1 = _mm512_sub_epi64(_mm512_slli_epi64(hi_lo, 32), hi_lo);  // hi_lo * EPS
  v8 r = _mm512_add_epi64(t0, t1);
  __mmask8 carry = _mm512_cmplt_epu64_mask(r, t0);
  r = _mm512_mask_add_epi64(r, carry, r, eps);
  __mmask8 ge = _mm512_cmpge_epu64_mask(r, p);
  return _mm512_mask_sub_epi64(r, ge, r, p);
}
// 64 x 64 -> 128 from four 32 x 32 -> 64 products (vpmuludq reads the low halves of each lane)
ORC_AVX512 void v8_mul_wide(v8 a, v8 b, v8& lo, v8& hi) {
  const v8 eps = v8_set1(EPS);
  v8 ah = _mm512_srli_epi64(a, 32), bh = _mm512_srli_epi64(b, 32);
  v8 ll = _mm512_mul_epu32(a, b), lh = _mm512_mul_epu32(a, bh), hl = _mm512_mul_epu32(ah, b), hh = _mm512_mul_epu32(ah, bh);
  v8 mid = _mm512_add_epi64(lh, _mm512_srli_epi64(ll, 32));                  // < 2^64
  v8 mid2 = _mm512_add_epi64(hl, _mm512_and_si512(mid, eps));                // < 2^64
  lo = _mm512_or_si512(_mm512_and_si512(ll, eps), _mm512_slli_epi64(mid2, 32));
  hi = _mm512_add_epi64(hh, _mm512_add_epi64(_mm512_srli_epi64(mid, 32), _mm512_srli_epi64(mid2, 32)));
}
ORC_AVX512 v8 v8_mul(v8 a, v8 b) {
  v8 lo, hi;
  v8_mul_wide(a, b, lo, hi);
  return v8_reduce128(lo, hi);
}
ORC_AVX512 v8 v8_sbox(v8 x) {
  v8 x2 = v8_mul(x, x), x4 = v8_mul(x2, x2), x3 = v8_mul(x, x2);
  return v8_mul(x3, x4);
}
// MDS entries are < 2^6: the 32-bit halves of the state accumulate in u64 without overflow (as mds_layer)
ORC_AVX512 void v8_mds_layer(v8* s) {
  const v8 eps = v8_set1(EPS);
  v8 lo[24], hi[24];
  for (int i = 0; i < 12; i++) {
    lo[i] = lo[i + 12] = _mm512_and_si512(s[i], eps);
    hi[i] = hi[i + 12] = _mm512_srli_epi64(s[i], 32);
  }
  for (int r = 0; r < 12; r++) {
    v8 al = _mm512_setzero_si512(), ah = _mm512_setzero_si512();
    for (int i = 0; i < 12; i++) {
      const v8 c = v8_set1(MDS_CIRC[i]);
      al = _mm512_add_epi64(al, _mm512_mul_epu32(lo[i + r], c));
      ah = _mm512_add_epi64(ah, _mm512_mul_epu32(hi[i + r], c));
    }
    if (MDS_DIAG[r]) {
      const v8 c = v8_set1(MDS_DIAG[r]);
      al = _mm512_add_epi64(al, _mm512_mul_epu32(lo[r], c));
      ah = _mm512_add_epi64(ah, _mm512_mul_epu32(hi[r], c));
    }
    // al + (ah << 32) as (lo, hi): ah < 2^44
    v8 sh = _mm512_slli_epi64(ah, 32);
    v8 l = _mm512_add_epi64(al, sh);
    __mmask8 c1 = _mm512_cmplt_epu64_mask(l, al);
    v8 h = _mm512_srli_epi64(ah, 32);
    h = _mm512_mask_add_epi64(h, c1, h, v8_set1(1));
    s[r] = v8_reduce128(l, h);
  }
}
// sum_i a[i] * b[i] (b: constants), n <= 12 terms: products reduced one by one (canonical), then added
ORC_AVX512 v8 v8_dot_const(const v8* a, const u64* b, int n) {
  v8 acc = _mm512_setzero_si512();
  for (int i = 0; i < n; i++) acc = v8_add(acc, v8_mul(a[i], v8_set1(b[i])));
  return acc;
}

// Eight permutations: s[i] holds state word i of every lane. Same round structure as poseidon().
ORC_AVX512_FN static void poseidon_x8(v8* s) {
  const PoseidonTables& T = tables();
  int round = 0;
  for (int r = 0; r < HALF_N_FULL; r++, round++) {
    for (int i = 0; i < 12; i++) s[i] = v8_sbox(v8_add(s[i], v8_set1(T.rc[12 * round + i])));
    v8_mds_layer(s);
  }
  for (int i = 0; i < 12; i++) s[i] = v8_add(s[i], v8_set1(T.fast_first[i]));
  {
    v8 o[12];
    o[0] = s[0];
    for (int c = 1; c < 12; c++) o[c] = v8_dot_const(&s[1], T.fast_init_t[c - 1], 11);
    for (int i = 0; i < 12; i++) s[i] = o[i];
  }
  for (int r = 0; r < N_PARTIAL; r++) {
    s[0] = v8_sbox(s[0]);
    if (r < N_PARTIAL - 1) s[0] = v8_add(s[0], v8_set1(T.fast_rc[r]));
    v8 d = v8_add(v8_mul(s[0], v8_set1(MDS_CIRC[0] + MDS_DIAG[0])), v8_dot_const(&s[1], T.fast_w_hat[r], 11));
    const v8 s0 = s[0];
    for (int i = 1; i < 12; i++) s[i] = v8_add(s[i], v8_mul(s0, v8_set1(T.fast_v[r][i - 1])));
    s[0] = d;
  }
  round += N_PARTIAL;
  for (int r = 0; r < HALF_N_FULL; r++, round++) {
    for (int i = 0; i < 12; i++) s[i] = v8_sbox(v8_add(s[i], v8_set1(T.rc[12 * round + i])));
    v8_mds_layer(s);
  }
}

// hash_or_noop of 8 rows (row j at rows + j * stride, `len` canonical words each) -> out[j]
ORC_AVX512_FN static void hash_or_noop_x8(const u64* rows, size_t stride, size_t len, Hash* out) {
  if (len <= 4) {
    for (int j = 0; j < 8; j++) out[j] = hash_or_noop(rows + j * stride, len);
    return;
  }
  const __m512i idx = _mm512_mullo_epi64(_mm512_set_epi64(7, 6, 5, 4, 3, 2, 1, 0), _mm512_set1_epi64((long long)stride));
  v8 s[12];
  for (int i = 0; i < 12; i++) s[i] = _mm512_setzero_si512();
  for (size_t off = 0; off < len; off += SPONGE_RATE) {
    const size_t take = len - off < (size_t)SPONGE_RATE ? len - off : SPONGE_RATE;
    for (size_t i = 0; i < take; i++) s[i] = _mm512_i64gather_epi64(idx, (const long long*)(rows + off + i), 8);  // overwrite mode
    poseidon_x8(s);
  }
  alignas(64) u64 tmp[4][8];
  for (int i = 0; i < 4; i++) _mm512_store_si512((void*)tmp[i], s[i]);
  for (int j = 0; j < 8; j++)
    for (int i = 0; i < 4; i++) out[j].e[i] = tmp[i][j];
}
// two_to_one of 8 sibling pairs: out[j] = H(in[2j] | in[2j+1])
ORC_AVX512_FN static void two_to_one_x8(const Hash* in, Hash* out) {
  alignas(64) u64 tmp[8][8];
  for (int j = 0; j < 8; j++)
    for (int i = 0; i < 8; i++) tmp[i][j] = in[2 * j + (i >> 2)].e[i & 3];
  v8 s[12];
  for (int i = 0; i < 8; i++) s[i] = _mm512_load_si512((const void*)tmp[i]);
  for (int i = 8; i < 12; i++) s[i] = _mm512_setzero_si512();
  poseidon_x8(s);
  alignas(64) u64 o[4][8];
  for (int i = 0; i < 4; i++) _mm512_store_si512((void*)o[i], s[i]);
  for (int j = 0; j < 8; j++)
    for (int i = 0; i < 4; i++) out[j].e[i] = o[i][j];
}

// Word 7 of the permutation of `st` with candidate w[j] written at position pos, for eight candidates.
ORC_AVX512_FN static void pow_responses_x8_avx(const State& st, size_t pos, const u64* w, u64* resp) {
  v8 s[12];
  for (int i = 0; i < 12; i++) s[i] = v8_set1(st[i]);
  s[pos] = _mm512_loadu_si512((const void*)w);
  poseidon_x8(s);
  _mm512_storeu_si512((void*)resp, s[7]);
}
static inline void pow_responses_x8(const State& st, size_t pos, const u64* w, u64* resp) {
  if (have_avx512()) return pow_responses_x8_avx(st, pos, w, resp);
  for (int j = 0; j < 8; j++) {
    State s2 = st;
    s2[pos] = w[j];
    poseidon(s2);
    resp[j] = s2[7];
  }
}

}  // namespace orc
