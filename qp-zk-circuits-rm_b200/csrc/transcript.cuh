// Device-resident Fiat-Shamir transcript for sm_100a: the duplex-sponge `Challenger` of qp-plonky2 1.1.1
// (iop/challenger.rs, un-vendored; driven by `prove()` reached from
// /root/reference/wormhole/prover/src/lib.rs:233-237) kept in HBM next to the challenges it produces, so
// that a whole proof is ONE stream-ordered enqueue: every kernel that needs a challenge reads it from
// `TranscriptDev`, nothing travels to the host between stages and the proving thread waits once, at the end.
//
// Semantics (pinned by the shipped proofs through tests/test_oracle_verifier.py):
//   observe(x): the output buffer is dropped; x joins the input buffer; 8 buffered inputs trigger a duplex.
//   duplex    : state[0..len) = inputs (OVERWRITE mode), permute, the 8 outputs are state[0..8).
//   get()     : duplex first if inputs are pending or no output is left; outputs are popped from the END.
// Inputs overwrite their state word as soon as they are observed (nothing reads the state between an
// observation and the duplex that follows it), so the sponge is 12 words plus two counters.
//
// One 16-lane group runs the sponge (poseidon_permute_coop: one state word per lane).
#pragma once
#include "poseidon.cuh"

namespace qpzk {

#define QPZK_MAX_CHALLENGES 4
#define QPZK_MAX_FRI_ROUNDS 16
#define QPZK_MAX_QUERIES 128

struct Challenges {
  u64 beta[QPZK_MAX_CHALLENGES], gamma[QPZK_MAX_CHALLENGES], alpha[QPZK_MAX_CHALLENGES];
};

// Everything Fiat-Shamir produces for one proof, in the order the prover needs it. Lives in device memory;
// copied back once with the proof pieces (the parity tests read the challenges from it).
struct TranscriptDev {
  u64 state[12];
  u32 in_len, out_len;
  u64 pi_hash[4];
  Challenges ch;
  u64 zeta[2], zeta_next[2];
  u64 fri_alpha[2];
  u64 fri_beta[QPZK_MAX_FRI_ROUNDS][2];
  u64 pow_witness, pow_resp;
  u64 xidx[QPZK_MAX_QUERIES];
};
// word offsets into TranscriptDev for the squeeze destination of k_transcript_step
#define QPZK_TR_OFF(field) ((u32)(offsetof(TranscriptDev, field) / 8))

struct TranscriptInit {  // the sponge after the host has observed circuit_digest | H(public_inputs)
  u64 state[12];
  u64 pi_hash[4];
};

__global__ void __launch_bounds__(32) k_transcript_init(TranscriptDev* __restrict__ T, TranscriptInit init) {
  const u32 t = threadIdx.x;
  if (t < 12) T->state[t] = init.state[t];
  if (t < 4) T->pi_hash[t] = init.pi_hash[t];
  if (t == 0) {
    T->in_len = 0;
    T->out_len = 8;  // eight observations = one duplex: its outputs are available
    T->pow_witness = ~0ull;
  }
}

// Per-lane view of the sponge (lane < 12 holds state[lane]); in_len / out_len are uniform.
struct CoopSponge {
  u64 s;
  u32 in_len, out_len;
};
GL_DEV void sponge_duplex(CoopSponge& sp, u32 lane, u64* xch) {
  sp.s = gl_canon(poseidon_permute_coop(sp.s, lane, xch, 0xffffu));
  sp.in_len = 0;
  sp.out_len = 8;
}
// src_mode 0: element i at src[i]; 1: extension coefficients stored SoA [2][m] observed interleaved
// (element i = src[(i & 1) * m + (i >> 1)]). Observed values are also written, canonical, to `copy` (may be null).
GL_DEV void sponge_observe(CoopSponge& sp, u32 lane, u64* xch, const u64* __restrict__ src, u32 n, u32 src_mode, u64 m,
                           u64* __restrict__ copy) {
  u32 i = 0;
  u64 pre = 0;         // this lane's element of the chunk that follows a permutation, requested BEFORE the
  bool have_pre = false;  // permutation so that its global-memory latency hides behind the ~6.5 us of arithmetic
  auto fetch = [&](u32 e) { return src_mode ? src[(u64)(e & 1) * m + (e >> 1)] : src[e]; };
  while (i < n) {
    const u32 room = 8 - sp.in_len, take = n - i < room ? n - i : room;
    if (lane >= sp.in_len && lane < sp.in_len + take) {
      const u32 e = i + lane - sp.in_len;
      u64 v = have_pre ? pre : fetch(e);
      v = gl_canon(v);
      sp.s = v;
      if (copy) copy[e] = v;
    }
    sp.in_len += take;
    sp.out_len = 0;
    i += take;
    have_pre = false;
    if (sp.in_len == 8) {
      if (i < n) {  // after the permutation in_len is 0: lane l takes element i + l
        if (lane < 8 && i + lane < n) pre = fetch(i + lane);
        have_pre = true;
      }
      sponge_duplex(sp, lane, xch);
    }
  }
}
GL_DEV u64 sponge_get(CoopSponge& sp, u32 lane, u64* xch) {
  if (sp.in_len != 0 || sp.out_len == 0) sponge_duplex(sp, lane, xch);
  sp.out_len--;
  const u32 lo = __shfl_sync(0xffffu, (u32)sp.s, sp.out_len, 16), hi = __shfl_sync(0xffffu, (u32)(sp.s >> 32), sp.out_len, 16);
  return ((u64)hi << 32) | lo;
}

// One transcript step: observe n elements, then squeeze nsq1 challenges into the words
// ((u64*)T)[dst1 + j] and nsq2 more into ((u64*)T)[dst2 + j] (betas then gammas share one step).
// post: 0 nothing; 1: also T->zeta_next = squeezed (zeta) * aux (aux = w_n, the generator of the trace
// subgroup); 2: squeezed words j >= 1 are masked with aux (query indices after the PoW response).
__global__ void __launch_bounds__(16)
k_transcript_step(TranscriptDev* __restrict__ T, const u64* __restrict__ src, u32 n, u32 src_mode, u64 m,
                  u64* __restrict__ copy, u32 nsq1, u32 dst1, u32 nsq2, u32 dst2, u32 post, u64 aux) {
  __shared__ u64 xch[COOP_XCH_WORDS];
  const u32 lane = threadIdx.x;
  CoopSponge sp;
  sp.s = lane < 12 ? T->state[lane] : 0;
  sp.in_len = T->in_len;
  sp.out_len = T->out_len;
  sponge_observe(sp, lane, xch, src, n, src_mode, m, copy);
  u64* words = reinterpret_cast<u64*>(T);
  for (u32 j = 0; j < nsq1 + nsq2; j++) {
    u64 v = sponge_get(sp, lane, xch);  // every lane holds the value
    if (post == 2 && j >= 1) v &= aux;
    QPZK_CHECK((j < nsq1 ? dst1 + j : dst2 + (j - nsq1)) < sizeof(TranscriptDev) / 8);
    if (lane == 0) words[j < nsq1 ? dst1 + j : dst2 + (j - nsq1)] = v;
    if (post == 1 && j < 2 && lane == 1) T->zeta_next[j] = gl_canon(gl_mul(v, aux));
  }
  if (lane < 12) T->state[lane] = sp.s;
  if (lane == 0) {
    T->in_len = sp.in_len;
    T->out_len = sp.out_len;
  }
}

// alpha_c^t for t < stride: the table the quotient kernel reduces constraint terms with (prover.cuh AlphaAcc).
__global__ void __launch_bounds__(256) k_alpha_powers(const TranscriptDev* __restrict__ T, u32 stride, u64* __restrict__ apw) {
  const u32 c = blockIdx.x;
  const u64 a = T->ch.alpha[c];
  for (u32 t = threadIdx.x; t < stride; t += blockDim.x) apw[c * stride + t] = gl_canon(gl_pow(a, t));
}

// ---- blinding salts drawn on the device ----
// plonky2 draws the SALT_SIZE blinding columns of a hiding oracle from the OS RNG (`F::rand()`), which makes
// reference proofs irreproducible and, for a GPU prover fed from the host, costs 12.6 MB of PCIe traffic per
// wormhole proof (4 columns x 2^17 x 8 B x 3 oracles). With QPZK_PROVE_SEEDED_SALTS the caller passes a 32-byte
// seed instead and the salts come from ChaCha8 keyed by it: block b of oracle o is the ChaCha8 block with key =
// seed, counter = b, nonce = (o, 0); its sixteen 32-bit words are eight salts (little-endian pairs), salt j of
// an oracle is element (column j / N, natural row j % N), values >= p wrap by p. The generator is a stream
// cipher, so the seeded mode is as hiding as the seed is secret; the parity tests restate it on the host.
struct SaltSeed {
  u32 key[8];
};
GL_DEV u32 rotl32(u32 x, int n) { return (x << n) | (x >> (32 - n)); }
__global__ void __launch_bounds__(128) k_salts_chacha8(SaltSeed seed, u32 oracle, u64 count /* salts, multiple of 8 */,
                                                       u64* __restrict__ out) {
  const u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (b * 8 >= count) return;
  u32 s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
#pragma unroll
  for (int i = 0; i < 8; i++) s[4 + i] = seed.key[i];
  s[12] = (u32)b;
  s[13] = (u32)(b >> 32);
  s[14] = oracle;
  s[15] = 0;
  u32 x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = s[i];
#define QPZK_QR(a, b_, c, d)                    \
  x[a] += x[b_]; x[d] = rotl32(x[d] ^ x[a], 16); \
  x[c] += x[d]; x[b_] = rotl32(x[b_] ^ x[c], 12); \
  x[a] += x[b_]; x[d] = rotl32(x[d] ^ x[a], 8);  \
  x[c] += x[d]; x[b_] = rotl32(x[b_] ^ x[c], 7);
#pragma unroll
  for (int r = 0; r < 4; r++) {  // 8 rounds = 4 double rounds
    QPZK_QR(0, 4, 8, 12) QPZK_QR(1, 5, 9, 13) QPZK_QR(2, 6, 10, 14) QPZK_QR(3, 7, 11, 15)
    QPZK_QR(0, 5, 10, 15) QPZK_QR(1, 6, 11, 12) QPZK_QR(2, 7, 8, 13) QPZK_QR(3, 4, 9, 14)
  }
#undef QPZK_QR
  u64 v[8];
#pragma unroll
  for (int e = 0; e < 8; e++) v[e] = gl_canon(((u64)(x[2 * e + 1] + s[2 * e + 1]) << 32) | (u64)(x[2 * e] + s[2 * e]));
  ulonglong2* o = reinterpret_cast<ulonglong2*>(out + b * 8);
#pragma unroll
  for (int e = 0; e < 4; e++) o[e] = make_ulonglong2(v[2 * e], v[2 * e + 1]);
}

// Small host values into device words without a host-to-device copy (a pageable cudaMemcpyAsync
// synchronises the stream): the per-stage hooks place caller-provided challenges this way.
struct Words16 {
  u64 w[16];
};
__global__ void k_set_words(u64* __restrict__ dst, Words16 v, u32 n) {
  if (threadIdx.x < n) dst[threadIdx.x] = v.w[threadIdx.x];
}

}  // namespace qpzk
