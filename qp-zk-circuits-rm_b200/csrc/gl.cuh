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

// x = r0 + 2^32 r1 + 2^64 r2 + 2^96 r3  ->  (r1:r0) - r3 + r2*(2^32-1)   (any-u64 representative)
// Carry-chain PTX, no compares or selects. NOTE on flags: CC.CF is the hardware carry, i.e. after a
// sub chain it is NOT-borrow. `subc m,0,0` after a SUB chain therefore yields the borrow mask
// (0 / 0xffffffff), but after an ADD chain it would yield the INVERTED carry mask - so add chains
// read the carry with `addc c,0,0` (0/1), negate it into a mask and add that.
GL_DEV u64 gl_reduce4(u32 r0, u32 r1, u32 r2, u32 r3) {
  asm("{\n\t"
      ".reg .u32 m, tl, th;\n\t"
      ".reg .u64 t;\n\t"
      "sub.cc.u32 %0, %0, %3;\n\t"     // (r1:r0) -= r3
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.u32 m, 0, 0;\n\t"          // borrow mask: -2^64 == -EPS
      "sub.cc.u32 %0, %0, m;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "mul.wide.u32 t, %2, 0xffffffff;\n\t"  // r2 * EPS
      "mov.b64 {tl, th}, t;\n\t"
      "add.cc.u32 %0, %0, tl;\n\t"
      "addc.cc.u32 %1, %1, th;\n\t"
      "addc.u32 m, 0, 0;\n\t"          // carry (0/1): +2^64 == +EPS (cannot carry again)
      "sub.u32 m, 0, m;\n\t"           // 0 / 0xffffffff == c*EPS (low word)
      "add.cc.u32 %0, %0, m;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "}"
      : "+r"(r0), "+r"(r1)
      : "r"(r2), "r"(r3));
  return ((u64)r1 << 32) | r0;
}
GL_DEV u64 gl_reduce128(u64 lo, u64 hi) {
  return gl_reduce4((u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32));
}

// 64x64 -> 128 schoolbook product as four 32-bit limbs (mad.lo.cc / madc.hi.cc chains map to
// IMAD.WIDE.U32 with carry-out on sm_100a).
GL_DEV void gl_mul_wide(u64 a, u64 b, u32& r0, u32& r1, u32& r2, u32& r3) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  asm("{\n\t"
      "mul.lo.u32 %0, %4, %6;\n\t"
      "mul.hi.u32 %1, %4, %6;\n\t"
      "mul.lo.u32 %2, %5, %7;\n\t"
      "mul.hi.u32 %3, %5, %7;\n\t"
      "mad.lo.cc.u32 %1, %4, %7, %1;\n\t"
      "madc.hi.cc.u32 %2, %4, %7, %2;\n\t"
      "addc.u32 %3, %3, 0;\n\t"
      "mad.lo.cc.u32 %1, %5, %6, %1;\n\t"
      "madc.hi.cc.u32 %2, %5, %6, %2;\n\t"
      "addc.u32 %3, %3, 0;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}

GL_DEV u64 gl_mul(u64 a, u64 b) {
  u32 r0, r1, r2, r3;
  gl_mul_wide(a, b, r0, r1, r2, r3);
  return gl_reduce4(r0, r1, r2, r3);
}
GL_DEV u64 gl_sqr(u64 a) { return gl_mul(a, a); }

// a*b + c, one reduction. a*b + c < 2^128 always.
GL_DEV u64 gl_mad(u64 a, u64 b, u64 c) {
  u32 r0, r1, r2, r3;
  gl_mul_wide(a, b, r0, r1, r2, r3);
  u32 c0 = (u32)c, c1 = (u32)(c >> 32);
  asm("add.cc.u32 %0, %0, %4;\n\t"
      "addc.cc.u32 %1, %1, %5;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t"
      "addc.u32 %3, %3, 0;"
      : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3)
      : "r"(c0), "r"(c1));
  return gl_reduce4(r0, r1, r2, r3);
}

// General add: both operands may be non-canonical. a + b = s + c*2^64, 2^64 == EPS; the first
// correction can wrap once more (only when s >= p), hence two corrections.
GL_DEV u64 gl_add(u64 a, u64 b) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  asm("{\n\t.reg .u32 c, m;\n\t.reg .u64 t;\n\t"
      "add.cc.u32 %0, %0, %2;\n\t"
      "addc.cc.u32 %1, %1, %3;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "sub.u32 m, 0, c;\n\t"           // 0 / 0xffffffff == low word of c*EPS
      "add.cc.u32 %0, %0, m;\n\t"
      "addc.cc.u32 %1, %1, 0;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "sub.u32 c, 0, c;\n\t"
      "add.cc.u32 %0, %0, c;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "}"
      : "+r"(a0), "+r"(a1)
      : "r"(b0), "r"(b1));
  return ((u64)a1 << 32) | a0;
}
// Add where b is canonical (< p): a single correction suffices.
GL_DEV u64 gl_add_c(u64 a, u64 b_canon) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b_canon, b1 = (u32)(b_canon >> 32);
  asm("{\n\t.reg .u32 c;\n\t.reg .u64 t;\n\t"
      "add.cc.u32 %0, %0, %2;\n\t"
      "addc.cc.u32 %1, %1, %3;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "sub.u32 c, 0, c;\n\t"
      "add.cc.u32 %0, %0, c;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "}"
      : "+r"(a0), "+r"(a1)
      : "r"(b0), "r"(b1));
  return ((u64)a1 << 32) | a0;
}
// General subtract: a - b = d - c*2^64; the correction -EPS can wrap once more (only if d < EPS).
GL_DEV u64 gl_sub(u64 a, u64 b) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  asm("{\n\t.reg .u32 m;\n\t"
      "sub.cc.u32 %0, %0, %2;\n\t"
      "subc.cc.u32 %1, %1, %3;\n\t"
      "subc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, m;\n\t"
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, m;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "}"
      : "+r"(a0), "+r"(a1)
      : "r"(b0), "r"(b1));
  return ((u64)a1 << 32) | a0;
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
