// Goldilocks field (p = 2^64 - 2^32 + 1) device arithmetic for sm_100a.
//
// Replaces `GoldilocksField` / `QuadraticExtension<GoldilocksField>` of qp-plonky2-field 1.1.1
// (pinned /root/reference/Cargo.lock:514-515; aliased `F`, `D = 2` at
// /root/reference/common/src/circuit.rs:10-12).
//
// Representation: a field element is ANY u64 (values >= p are allowed in flight); `gl_canon`
// produces the canonical representative and is applied wherever a value leaves the device or is
// compared. Products are 64x64->128 (mul.lo/mul.hi.u64 -> IMAD.WIDE.U32 chains) followed by the
// Goldilocks fold 2^64 = 2^32 - 1, 2^96 = -1 (mod p).
#pragma once
#include <cstdint>

typedef uint64_t u64;
typedef uint32_t u32;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL
#define GL_GEN 14293326489335486720ULL       // MULTIPLICATIVE_GROUP_GENERATOR == coset shift
#define GL_ROOT_2_32 7277203076849721926ULL  // POWER_OF_TWO_GENERATOR

#define GL_DEV __device__ __forceinline__
#define GL_HD __host__ __device__ __forceinline__

GL_DEV u64 gl_canon(u64 a) { return a >= GL_P ? a - GL_P : a; }

// x = lo + 2^64*hi  ->  lo - hi_hi + hi_lo*(2^32-1)   (result is any-u64 representative)
GL_DEV u64 gl_reduce128(u64 lo, u64 hi) {
  u32 hi_hi = (u32)(hi >> 32), hi_lo = (u32)hi;
  u64 t0 = lo - (u64)hi_hi;
  if (lo < (u64)hi_hi) t0 -= GL_EPS;          // borrow: -2^64 == -EPS
  u64 t1 = (u64)hi_lo * (u64)0xFFFFFFFFu;     // mul.wide.u32
  u64 r = t0 + t1;
  if (r < t1) r += GL_EPS;                    // carry: +2^64 == +EPS (cannot overflow again)
  return r;
}

GL_DEV u64 gl_mul(u64 a, u64 b) { return gl_reduce128(a * b, __umul64hi(a, b)); }
GL_DEV u64 gl_sqr(u64 a) { return gl_mul(a, a); }

// a*b + c, one reduction. a*b + c < 2^128 always.
GL_DEV u64 gl_mad(u64 a, u64 b, u64 c) {
  u64 lo = a * b, hi = __umul64hi(a, b);
  lo += c;
  hi += (lo < c);
  return gl_reduce128(lo, hi);
}

// General add: both operands may be non-canonical.
GL_DEV u64 gl_add(u64 a, u64 b) {
  u64 s = a + b;
  if (s < a) {            // wrapped: +EPS, which itself can wrap once more only if s >= p
    u64 t = s + GL_EPS;
    s = (t < s) ? t + GL_EPS : t;
  }
  return s;
}
// Add where b is canonical (< p): a single correction suffices.
GL_DEV u64 gl_add_c(u64 a, u64 b_canon) {
  u64 s = a + b_canon;
  if (s < a) s += GL_EPS;
  return s;
}
// General subtract.
GL_DEV u64 gl_sub(u64 a, u64 b) {
  u64 d = a - b;
  if (a < b) {            // wrapped: -EPS, can wrap again only if d < EPS
    u64 t = d - GL_EPS;
    d = (d < GL_EPS) ? t - GL_EPS : t;
  }
  return d;
}
GL_DEV u64 gl_neg(u64 a) {
  u64 c = gl_canon(a);
  return c ? GL_P - c : 0;
}

GL_DEV u64 gl_pow(u64 b, u64 e) {
  u64 r = 1;
  while (e) {
    if (e & 1) r = gl_mul(r, b);
    b = gl_sqr(b);
    e >>= 1;
  }
  return r;
}
GL_DEV u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }

// ---- quadratic extension F_p[X]/(X^2 - 7) ----
struct gl2 {
  u64 a, b;
};
GL_DEV gl2 gl2_make(u64 a, u64 b) { gl2 r; r.a = a; r.b = b; return r; }
GL_DEV gl2 gl2_add(gl2 x, gl2 y) { return gl2_make(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
GL_DEV gl2 gl2_sub(gl2 x, gl2 y) { return gl2_make(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
GL_DEV gl2 gl2_mul(gl2 x, gl2 y) {
  u64 bb = gl_mul(x.b, y.b);
  u64 a = gl_mad(x.a, y.a, gl_mul(bb, 7));
  u64 b = gl_mad(x.a, y.b, gl_mul(x.b, y.a));
  return gl2_make(a, b);
}
GL_DEV gl2 gl2_scale(gl2 x, u64 s) { return gl2_make(gl_mul(x.a, s), gl_mul(x.b, s)); }
GL_DEV gl2 gl2_canon(gl2 x) { return gl2_make(gl_canon(x.a), gl_canon(x.b)); }
GL_DEV gl2 gl2_inv(gl2 x) {
  u64 d = gl_sub(gl_sqr(x.a), gl_mul(7, gl_sqr(x.b)));
  u64 di = gl_inv(d);
  return gl2_make(gl_mul(x.a, di), gl_mul(gl_neg(x.b), di));
}

// ---- host-side reference arithmetic for table generation (never on the data path) ----
namespace glh {
typedef unsigned __int128 u128;
static inline u64 mul(u64 a, u64 b) { return (u64)(((u128)a * b) % GL_P); }
static inline u64 add(u64 a, u64 b) { return (u64)(((u128)a + b) % GL_P); }
static inline u64 sub(u64 a, u64 b) { return (u64)(((u128)a + GL_P - b) % GL_P); }
static inline u64 pow(u64 b, u64 e) {
  u64 r = 1;
  while (e) {
    if (e & 1) r = mul(r, b);
    b = mul(b, b);
    e >>= 1;
  }
  return r;
}
static inline u64 inv(u64 a) { return pow(a, GL_P - 2); }
static inline u64 root_of_unity(unsigned bits) {  // POWER_OF_TWO_GENERATOR^(2^(32-bits))
  u64 r = GL_ROOT_2_32;
  for (unsigned i = bits; i < 32; i++) r = mul(r, r);
  return r;
}
}  // namespace glh
