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

// -DQPZK_CHECKED (make libqpzk_checked.so): every tile, level, wire and table index computed on the device is
// range-checked and a violation traps the kernel (the context then reports an error on its next call). This is the
// stand-in for compute-sanitizer, which the GPU pool does not allow; tests/test_gpu_checked_build.py runs the
// small-case suite against the checked library.
#ifdef QPZK_CHECKED
#include <cstdio>
#define QPZK_CHECK(cond)                                                                              \
  do {                                                                                                \
    if (!(cond)) {                                                                                    \
      printf("QPZK_CHECK failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, \
             (int)threadIdx.x);                                                                       \
      __trap();                                                                                       \
    }                                                                                                 \
  } while (0)
#else
#define QPZK_CHECK(cond) ((void)0)
#endif

#define GL_DEV __device__ __forceinline__
#ifndef GL_FOLD_ALU
#define GL_FOLD_ALU 1
#endif
#define GL_HD __host__ __device__ __forceinline__

GL_DEV u64 gl_canon(u64 a) { return a >= GL_P ? a - GL_P : a; }

// The fold  lo + r2*EPS - h  (mod p)  for lo any u64, r2 < 2^32, h <= 2^63: the Goldilocks reduction
// 2^64 == EPS, 2^96 == -1 applied to x = lo + 2^64*r2 + 2^96*h. Returns some u64 representative.
//
// Written with a signed 128-bit intermediate ON PURPOSE. PTX only has two-input adds with a single carry
// flag, so the same fold as add.cc/subc chains needs a mask-and-correct sequence per wrap: 12 SASS
// instructions. From the __int128 expression ptxas emits three-input IADD3 with TWO carry-outs, sums both
// carries into one k in {-1, 0, 1} (`IADD3.X k = 0 - 1 + P + P'`) and a single correction by k*EPS
// follows: 6-7 instructions. The single correction never wraps again: if k = 1 then t mod 2^64 <=
// 2^64 - 2^33, and if k = -1 then t mod 2^64 >= 2^64 - h.
// (A note for anyone writing carry chains in PTX here: CC.CF is the hardware carry, so after a SUB chain
// `subc m,0,0` yields the borrow mask, but after an ADD chain it yields the INVERTED carry mask.)
//
// Two spellings. They differ only in which pipe pays for r2*EPS and k*EPS:
//  gl_fold_alu: (r2 << 32) - r2 and (k << 32) - k as adds. Default: Poseidon saturates the multiplier
//               (FMA-heavy) pipe (ncu: 88 % with the multiplier spelling, ALU at 43 %).
//  gl_fold_mul: two IMAD.WIDE. Measured on the NTT kernels, which saturate the ALU pipe instead
//               (ALU 76 %): no difference there - ptxas rewrites multiplications by 2^32 - 1 into
//               shift/subtract itself - so it is kept only as the -DGL_FOLD_ALU=0 experiment.
GL_DEV u64 gl_fold_alu(u64 lo, u32 r2, u64 h) {
  __int128 t = (__int128)(unsigned __int128)lo - (__int128)(unsigned __int128)((u64)r2 + h) +
               (__int128)(unsigned __int128)((u64)r2 << 32);
  u64 tl = (u64)t;
  u32 k = (u32)(u64)(t >> 64);  // 0, 1 or 0xffffffff
  u32 tl0 = (u32)tl, tl1 = (u32)(tl >> 32);
  u32 sx = (u32)((int32_t)k >> 31);
  asm("sub.cc.u32 %0, %0, %2; subc.u32 %1, %1, %3; add.u32 %1, %1, %2;" : "+r"(tl0), "+r"(tl1) : "r"(k), "r"(sx));
  return ((u64)tl1 << 32) | tl0;
}
GL_DEV u64 gl_fold_mul(u64 lo, u32 r2, u64 h) {
  __int128 t = (__int128)(unsigned __int128)lo + (__int128)((u64)r2 * GL_EPS) - (__int128)(unsigned __int128)h;
  u64 tl = (u64)t;
  u32 k = (u32)(u64)(t >> 64);  // 0, 1 or 0xffffffff
  u64 r = (u64)k * GL_EPS + tl;  // k = -1: adds 2^64 - 2^33 + 1, i.e. -EPS - 2^32 (mod 2^64) ...
  u32 rl = (u32)r, rh = (u32)(r >> 32);
  asm("mad.hi.u32 %0, %1, 2, %0;" : "+r"(rh) : "r"(k));  // ... so give the 2^32 back: hi += k >> 31
  return ((u64)rh << 32) | rl;
}
// Third spelling: r2*EPS rides on ONE multiply-add whose addend is lo (IMAD.WIDE with carry-out), -h is a
// two-word subtract, the carry and the borrow meet in k = C - B and the usual single correction follows:
// 7-8 instructions, one of them on the multiplier pipe.
GL_DEV u64 gl_fold_mad(u64 lo, u32 r2, u64 h) {
  u32 t0 = (u32)lo, t1 = (u32)(lo >> 32), h0 = (u32)h, h1 = (u32)(h >> 32), k, sx;
  asm("{\n\t.reg .u32 c;\n\t"
      "mad.lo.cc.u32 %0, %4, 0xffffffff, %0;\n\t"
      "madc.hi.cc.u32 %1, %4, 0xffffffff, %1;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, %5;\n\t"
      "subc.cc.u32 %1, %1, %6;\n\t"
      "subc.u32 %2, c, 0;\n\t"           // k = carry - borrow: 0, 1 or 0xffffffff
      "shr.s32 %3, %2, 31;\n\t"
      "sub.cc.u32 %0, %0, %2;\n\t"       // + k*EPS = + (k << 32) - k
      "subc.u32 %1, %1, %3;\n\t"
      "add.u32 %1, %1, %2;\n\t"
      "}"
      : "+r"(t0), "+r"(t1), "=&r"(k), "=&r"(sx)
      : "r"(r2), "r"(h0), "r"(h1));
  return ((u64)t1 << 32) | t0;
}
GL_DEV u64 gl_fold(u64 lo, u32 r2, u64 h) {
#if GL_FOLD_ALU == 1
  return gl_fold_alu(lo, r2, h);
#elif GL_FOLD_ALU == 2
  return gl_fold_mad(lo, r2, h);
#else
  return gl_fold_mul(lo, r2, h);
#endif
}
// x = r0 + 2^32 r1 + 2^64 r2 + 2^96 r3 (+ 2^128 r4, r4 < 2^31)  ->  (r1:r0) + r2*EPS - (r4:r3)
GL_DEV u64 gl_reduce4(u32 r0, u32 r1, u32 r2, u32 r3) { return gl_fold(((u64)r1 << 32) | r0, r2, r3); }
GL_DEV u64 gl_reduce5(u32 r0, u32 r1, u32 r2, u32 r3, u32 r4) {
  return gl_fold(((u64)r1 << 32) | r0, r2, ((u64)r4 << 32) | r3);
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

// The same products spelled for a kernel whose limiter is the multiplier (FMA-heavy) pipe - Poseidon. On
// sm_100a every IMAD.* form issues to that pipe: IMAD.WIDE / IMAD.HI hold it for 4 cycles per warp, IMAD,
// IMAD.X and IMAD.MOV for 2. From the mad.lo.cc / madc.hi.cc chains of gl_mul_wide ptxas emits, per product,
// 3 IMAD.WIDE + IMAD + IMAD.HI (the a1*b1 product split in two because its halves are not a register pair)
// + IMAD.MOV + 2 IMAD.X = 24 pipe cycles, and it computes a0*a1 twice for a square (26 cycles). Here the
// partial products are four (three for a square) independent mul.wide - IMAD.WIDE with a zero addend - and
// every addition is an add.cc / addc chain that ptxas keeps on the ALU pipe as IADD3 / IADD3.X: 16 + 4 and
// 12 + 2 pipe cycles with the same (mul) or a smaller (sqr: 18 against 20) instruction count.
GL_DEV void gl_mul_wide_alu(u64 a, u64 b, u32& r0, u32& r1, u32& r2, u32& r3) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  asm("{\n\t.reg .b64 p, q, s, t;\n\t.reg .b32 p1, q0, q1, s0, s1, t0, t1, m2;\n\t"
      "mul.wide.u32 p, %4, %6;\n\t"
      "mul.wide.u32 q, %4, %7;\n\t"
      "mul.wide.u32 s, %5, %6;\n\t"
      "mul.wide.u32 t, %5, %7;\n\t"
      "mov.b64 {%0, p1}, p;\n\t"
      "mov.b64 {q0, q1}, q;\n\t"
      "mov.b64 {s0, s1}, s;\n\t"
      "mov.b64 {t0, t1}, t;\n\t"
      "add.cc.u32 q0, q0, s0;\n\t"
      "addc.cc.u32 q1, q1, s1;\n\t"
      "addc.u32 m2, 0, 0;\n\t"
      "add.cc.u32 %1, p1, q0;\n\t"
      "addc.cc.u32 %2, t0, q1;\n\t"
      "addc.u32 %3, t1, m2;\n\t"
      "}"
      : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}
GL_DEV void gl_sqr_wide_alu(u64 a, u32& r0, u32& r1, u32& r2, u32& r3) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32);
  asm("{\n\t.reg .b64 p, q, t;\n\t.reg .b32 p1, q0, q1, t0, t1, m0, m1, m2;\n\t"
      "mul.wide.u32 p, %4, %4;\n\t"
      "mul.wide.u32 q, %4, %5;\n\t"
      "mul.wide.u32 t, %5, %5;\n\t"
      "mov.b64 {%0, p1}, p;\n\t"
      "mov.b64 {q0, q1}, q;\n\t"
      "mov.b64 {t0, t1}, t;\n\t"
      "shl.b32 m0, q0, 1;\n\t"              // 2*a0*a1 as 65 bits
      "shf.l.wrap.b32 m1, q0, q1, 1;\n\t"
      "shr.u32 m2, q1, 31;\n\t"
      "add.cc.u32 %1, p1, m0;\n\t"
      "addc.cc.u32 %2, t0, m1;\n\t"
      "addc.u32 %3, t1, m2;\n\t"
      "}"
      : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
      : "r"(a0), "r"(a1));
}
GL_DEV u64 gl_mul_alu(u64 a, u64 b) {
  u32 r0, r1, r2, r3;
  gl_mul_wide_alu(a, b, r0, r1, r2, r3);
  return gl_reduce4(r0, r1, r2, r3);
}
GL_DEV u64 gl_sqr_alu(u64 a) {
  u32 r0, r1, r2, r3;
  gl_sqr_wide_alu(a, r0, r1, r2, r3);
  return gl_reduce4(r0, r1, r2, r3);
}

// a*b + c, one reduction (a*b + c < 2^128 always). The addend rides on the first partial product:
// (r1:r0) = a0*b0 + c is one IMAD.WIDE with carry-out, (r3:r2) = a1*b1 + carry one IMAD.WIDE.X.
GL_DEV u64 gl_mad(u64 a, u64 b, u64 c) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  u32 r0 = (u32)c, r1 = (u32)(c >> 32), r2, r3;
  asm("{\n\t"
      "mad.lo.cc.u32 %0, %4, %6, %0;\n\t"
      "madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
      "madc.lo.cc.u32 %2, %5, %7, 0;\n\t"
      "madc.hi.u32 %3, %5, %7, 0;\n\t"
      "mad.lo.cc.u32 %1, %4, %7, %1;\n\t"
      "madc.hi.cc.u32 %2, %4, %7, %2;\n\t"
      "addc.u32 %3, %3, 0;\n\t"
      "mad.lo.cc.u32 %1, %5, %6, %1;\n\t"
      "madc.hi.cc.u32 %2, %5, %6, %2;\n\t"
      "addc.u32 %3, %3, 0;\n\t"
      "}"
      : "+r"(r0), "+r"(r1), "=&r"(r2), "=&r"(r3)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
  return gl_reduce4(r0, r1, r2, r3);
}

// a + b = s + c*2^64 with 2^64 == EPS. The corrections are multiply-adds by the carry (one IMAD.WIDE
// with carry-out each) instead of mask-and-add chains.
// General add: both operands may be non-canonical; the first correction can wrap once more (only
// when s >= p), hence two corrections.
GL_DEV u64 gl_add(u64 a, u64 b) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  asm("{\n\t.reg .u32 c;\n\t"
      "add.cc.u32 %0, %0, %2;\n\t"
      "addc.cc.u32 %1, %1, %3;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "mad.lo.cc.u32 %0, c, 0xffffffff, %0;\n\t"
      "madc.hi.cc.u32 %1, c, 0xffffffff, %1;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "mad.lo.cc.u32 %0, c, 0xffffffff, %0;\n\t"
      "madc.hi.u32 %1, c, 0xffffffff, %1;\n\t"
      "}"
      : "+r"(a0), "+r"(a1)
      : "r"(b0), "r"(b1));
  return ((u64)a1 << 32) | a0;
}
// Add where b is canonical (< p): a single correction suffices.
GL_DEV u64 gl_add_c(u64 a, u64 b_canon) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b_canon, b1 = (u32)(b_canon >> 32);
  asm("{\n\t.reg .u32 c;\n\t"
      "add.cc.u32 %0, %0, %2;\n\t"
      "addc.cc.u32 %1, %1, %3;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "mad.lo.cc.u32 %0, c, 0xffffffff, %0;\n\t"
      "madc.hi.u32 %1, c, 0xffffffff, %1;\n\t"
      "}"
      : "+r"(a0), "+r"(a1)
      : "r"(b0), "r"(b1));
  return ((u64)a1 << 32) | a0;
}
// gl_add_c with both corrections on the ALU pipe (no IMAD / IMAD.HI by 2^32 - 1).
GL_DEV u64 gl_add_c_alu(u64 a, u64 b_canon) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b_canon, b1 = (u32)(b_canon >> 32);
  asm("{\n\t.reg .u32 c;\n\t"
      "add.cc.u32 %0, %0, %2;\n\t"
      "addc.cc.u32 %1, %1, %3;\n\t"
      "addc.u32 c, 0, 0;\n\t"            // carry: + 2^64 == + EPS = + 2^32 - 1: low word - c, high word + c - borrow
      "sub.cc.u32 %0, %0, c;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "add.u32 %1, %1, c;\n\t"
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
