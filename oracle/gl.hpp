// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under qp-zk-circuits-rm_b200/ may include,
// link or call this. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, as the checker.
//
// Goldilocks field F_p, p = 2^64 - 2^32 + 1, and its quadratic extension F_p[X]/(X^2 - 7).
//
// Restates the arithmetic of the un-vendored dependency qp-plonky2-field 1.1.1
// (pinned at /root/reference/Cargo.lock:514-515; aliased as `F = GoldilocksField`, `D = 2` at
// /root/reference/common/src/circuit.rs:10-12). Constants per SURVEY.md App. A.1.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>

namespace orc {

typedef uint64_t u64;
typedef unsigned __int128 u128;

static const u64 P = 0xFFFFFFFF00000001ULL;
static const u64 EPS = 0xFFFFFFFFULL;  // 2^32 - 1 == 2^64 mod p
// MULTIPLICATIVE_GROUP_GENERATOR (also coset_shift()) and POWER_OF_TWO_GENERATOR.
static const u64 GEN = 14293326489335486720ULL;
static const u64 ROOT_2_32 = 7277203076849721926ULL;

static inline u64 canon(u64 a) { return a >= P ? a - P : a; }

// Branch-free on purpose: the carries are data-dependent coin flips and mispredict badly.
static inline u64 add(u64 a, u64 b) {  // a, b canonical
  u64 s = a + b;
  u64 over = (u64)(s < a) | (u64)(s >= P);
  return s - (P & (0 - over));
}
static inline u64 sub(u64 a, u64 b) {
  u64 d = a - b;
  return d + (P & (0 - (u64)(a < b)));
}
static inline u64 neg(u64 a) { return a ? P - a : 0; }
static inline u64 reduce128_slow(u128 x) { return (u64)(x % P); }  // definition, kept as cross-check
// 2^64 == 2^32 - 1, 2^96 == -1 (mod p): x = lo + 2^64*hi_lo + 2^96*hi_hi == lo - hi_hi + EPS*hi_lo.
static inline u64 reduce128(u128 x) {
  u64 lo = (u64)x, hi = (u64)(x >> 64);
  u64 hi_hi = hi >> 32, hi_lo = hi & EPS;
  u64 t0, r;
  u64 borrow = __builtin_sub_overflow(lo, hi_hi, &t0);
  t0 -= EPS & (0 - borrow);
  u64 t1 = (hi_lo << 32) - hi_lo;  // hi_lo * EPS
  u64 carry = __builtin_add_overflow(t0, t1, &r);
  r += EPS & (0 - carry);
  return r - (P & (0 - (u64)(r >= P)));
}
static inline u64 mul(u64 a, u64 b) { return reduce128((u128)a * b); }
static inline u64 sqr(u64 a) { return mul(a, a); }

static inline u64 pow(u64 b, u64 e) {
  u64 r = 1;
  while (e) {
    if (e & 1) r = mul(r, b);
    b = mul(b, b);
    e >>= 1;
  }
  return r;
}
static inline u64 inv(u64 a) { return pow(a, P - 2); }

// primitive_root_of_unity(bits) = POWER_OF_TWO_GENERATOR^(2^(32-bits))
static inline u64 root_of_unity(unsigned bits) {
  u64 r = ROOT_2_32;
  for (unsigned i = bits; i < 32; i++) r = sqr(r);
  return r;
}

static inline unsigned log2_strict(size_t n) {
  unsigned k = 0;
  while (((size_t)1 << k) < n) k++;
  return k;
}
static inline size_t bitrev(size_t x, unsigned bits) {
  if (bits == 0) return 0;
  u64 v = (u64)x;
  v = ((v >> 1) & 0x5555555555555555ULL) | ((v & 0x5555555555555555ULL) << 1);
  v = ((v >> 2) & 0x3333333333333333ULL) | ((v & 0x3333333333333333ULL) << 2);
  v = ((v >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((v & 0x0F0F0F0F0F0F0F0FULL) << 4);
  v = __builtin_bswap64(v);
  return (size_t)(v >> (64 - bits));
}

// Quadratic extension, W = 7.
struct E2 {
  u64 a, b;  // a + b X
};
static inline E2 e2(u64 a, u64 b = 0) { return E2{a, b}; }
static inline bool operator==(E2 x, E2 y) { return x.a == y.a && x.b == y.b; }
static inline bool operator!=(E2 x, E2 y) { return !(x == y); }
static inline E2 operator+(E2 x, E2 y) { return E2{add(x.a, y.a), add(x.b, y.b)}; }
static inline E2 operator-(E2 x, E2 y) { return E2{sub(x.a, y.a), sub(x.b, y.b)}; }
static inline E2 operator*(E2 x, E2 y) {
  return E2{add(mul(x.a, y.a), mul(7, mul(x.b, y.b))), add(mul(x.a, y.b), mul(x.b, y.a))};
}
static inline E2 scale(E2 x, u64 s) { return E2{mul(x.a, s), mul(x.b, s)}; }
static inline E2 e2inv(E2 x) {
  u64 d = sub(sqr(x.a), mul(7, sqr(x.b)));
  u64 di = inv(d);
  return E2{mul(x.a, di), mul(neg(x.b), di)};
}
static inline E2 e2pow(E2 b, u64 e) {
  E2 r = e2(1);
  while (e) {
    if (e & 1) r = r * b;
    b = b * b;
    e >>= 1;
  }
  return r;
}

}  // namespace orc
