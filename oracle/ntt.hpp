// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// Reference transforms with plonky2's conventions:
//   PolynomialValues::ifft        values on H=<w_n>, natural order  -> coefficients, natural order
//   PolynomialCoeffs::lde(r)      zero-pad to n*2^r coefficients
//   PolynomialCoeffs::coset_fft(s) evaluations on s*<w_N>, natural order
// These are the bodies behind `PolynomialBatch::from_values/from_coeffs`, reached from
// /root/reference/wormhole/prover/src/lib.rs:233-237 (prove) and
// /root/reference/wormhole/circuit/src/circuit.rs:98-108 (build). Plain iterative radix-2; a naive
// O(n^2) DFT is kept as an independent cross-check for small n.
#pragma once
#include <mutex>
#include <vector>

#include "gl.hpp"
#include "poseidon_avx512.hpp"

namespace orc {

// Twiddles of one butterfly stage, w_m^j for j < m/2 (m = 2^s): they depend on the stage only, so they are
// built once per process.
static inline const std::vector<u64>& stage_twiddles(unsigned s) {
  static std::vector<u64> tabs[40];
  static std::once_flag once[40];
  std::call_once(once[s], [s] {
    size_t half = (size_t)1 << (s - 1);
    std::vector<u64> tw(half);
    u64 wm = root_of_unity(s);
    tw[0] = 1;
    for (size_t j = 1; j < half; j++) tw[j] = mul(tw[j - 1], wm);
    tabs[s] = std::move(tw);
  });
  return tabs[s];
}

// One stage over the whole array, eight butterflies per instruction (poseidon_avx512.hpp arithmetic); half >= 8.
ORC_AVX512_FN static void fft_stage_x8(u64* a, size_t n, size_t half, const u64* tw) {
  const v8 p = v8_set1(P);
  for (size_t base = 0; base < n; base += 2 * half)
    for (size_t j = 0; j < half; j += 8) {
      v8 u = _mm512_loadu_si512((const void*)(a + base + j));
      v8 t = v8_mul(_mm512_loadu_si512((const void*)(tw + j)), _mm512_loadu_si512((const void*)(a + base + j + half)));
      _mm512_storeu_si512((void*)(a + base + j), v8_add(u, t));
      v8 d = _mm512_sub_epi64(u, t);
      __mmask8 borrow = _mm512_cmplt_epu64_mask(u, t);
      _mm512_storeu_si512((void*)(a + base + j + half), _mm512_mask_add_epi64(d, borrow, d, p));
    }
}

// In-place forward DFT: out[i] = sum_j a[j] * w^(i*j), natural order in and out.
static inline void fft_inplace(std::vector<u64>& a) {
  size_t n = a.size();
  if (n <= 1) return;
  unsigned k = log2_strict(n);
  for (size_t i = 0; i < n; i++) {
    size_t j = bitrev(i, k);
    if (i < j) std::swap(a[i], a[j]);
  }
  const bool vec = have_avx512();
  for (unsigned s = 1; s <= k; s++) {
    size_t m = (size_t)1 << s, half = m >> 1;
    const std::vector<u64>& tw = stage_twiddles(s);
    if (vec && half >= 8) {
      fft_stage_x8(a.data(), n, half, tw.data());
      continue;
    }
    for (size_t base = 0; base < n; base += m)
      for (size_t j = 0; j < half; j++) {
        u64 t = mul(tw[j], a[base + j + half]);
        u64 u = a[base + j];
        a[base + j] = add(u, t);
        a[base + j + half] = sub(u, t);
      }
  }
}

static inline std::vector<u64> fft(std::vector<u64> coeffs) {
  fft_inplace(coeffs);
  return coeffs;
}

// values -> coefficients: (1/n) * DFT with w^-1, i.e. forward DFT then reverse [1..n) and scale.
static inline std::vector<u64> ifft(std::vector<u64> values) {
  size_t n = values.size();
  fft_inplace(values);
  u64 ninv = inv((u64)n % P);
  std::vector<u64> out(n);
  for (size_t i = 0; i < n; i++) out[i] = mul(values[(n - i) % n], ninv);
  return out;
}

static inline std::vector<u64> coset_fft(std::vector<u64> coeffs, u64 shift) {
  u64 s = 1;
  for (size_t i = 0; i < coeffs.size(); i++) {
    coeffs[i] = mul(coeffs[i], s);
    s = mul(s, shift);
  }
  fft_inplace(coeffs);
  return coeffs;
}

static inline std::vector<u64> coset_ifft(std::vector<u64> values, u64 shift) {
  std::vector<u64> c = ifft(std::move(values));
  u64 si = inv(shift), s = 1;
  for (size_t i = 0; i < c.size(); i++) {
    c[i] = mul(c[i], s);
    s = mul(s, si);
  }
  return c;
}

static inline std::vector<u64> lde(const std::vector<u64>& coeffs, unsigned rate_bits) {
  std::vector<u64> out(coeffs.size() << rate_bits, 0);
  for (size_t i = 0; i < coeffs.size(); i++) out[i] = coeffs[i];
  return out;
}

// Independent cross-check: naive evaluation of the polynomial at shift * w_n^i.
static inline std::vector<u64> naive_coset_eval(const std::vector<u64>& coeffs, size_t npoints,
                                                u64 shift) {
  u64 w = root_of_unity(log2_strict(npoints));
  std::vector<u64> out(npoints);
  u64 x = shift;
  for (size_t i = 0; i < npoints; i++) {
    u64 acc = 0;
    for (size_t j = coeffs.size(); j-- > 0;) acc = add(mul(acc, x), coeffs[j]);
    out[i] = acc;
    x = mul(x, w);
  }
  return out;
}

// Extension-field polynomial helpers (coefficients as E2).
static inline E2 eval_poly_e2(const std::vector<E2>& c, E2 x) {
  E2 acc = e2(0);
  for (size_t j = c.size(); j-- > 0;) acc = acc * x + c[j];
  return acc;
}
static inline E2 eval_base_poly_at_e2(const u64* c, size_t n, E2 x) {
  E2 acc = e2(0);
  for (size_t j = n; j-- > 0;) acc = acc * x + e2(c[j]);
  return acc;
}

}  // namespace orc
