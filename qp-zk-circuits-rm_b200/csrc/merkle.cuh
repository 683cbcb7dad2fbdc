// Poseidon leaf hashing and Merkle levels for sm_100a.
//
// Replaces `MerkleTree::new` / `fill_digests_buf` / `fill_subtree` / `MerkleTree::prove` of
// qp-plonky2 1.1.1 (hash/merkle_tree.rs, un-vendored), reached from every commit under
// /root/reference/wormhole/prover/src/lib.rs:233-237 and
// /root/reference/wormhole/circuit/src/circuit.rs:98-108.
//
// Device layout (B200-first, not plonky2's): leaves stay COLUMN-major ([width][N], the NTT's
// natural output), so the row "transpose" of the reference disappears - a thread that hashes row
// L reads lde[c][L] and the warp's loads are unit-stride. Digests are stored level by level
// (level 0 = leaf digests, level l has N >> l entries, the last level is the cap); plonky2's
// interleaved `digests` vector is produced only on export.
#pragma once
#include "poseidon.cuh"

namespace qpzk {

// hash_or_noop of one row per thread. Element (row, c) lives at src[row*row_stride + c*col_stride].
#ifndef QPZK_LEAF_MINB
#define QPZK_LEAF_MINB 5
#endif
__global__ void __launch_bounds__(128, QPZK_LEAF_MINB)
k_leaf_hash(const u64* __restrict__ src, u64 row_stride, u64 col_stride, u32 width, u64 nrows,
            u64* __restrict__ digests) {
  u64 row = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const u64* p = src + row * row_stride;
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = 0;
  if (width <= 4) {  // hash_or_noop: short rows are copied, zero padded
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (i < width) s[i] = gl_canon(p[i * col_stride]);
  } else {
    for (u32 off = 0; off < width; off += 8) {
      u32 len = width - off < 8 ? width - off : 8;
#pragma unroll
      for (int i = 0; i < 8; i++)
        if (i < len) s[i] = __ldg(p + (u64)(off + i) * col_stride);  // overwrite-mode absorb
      poseidon_permute(s);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) s[i] = gl_canon(s[i]);
  }
  ulonglong2* o = reinterpret_cast<ulonglong2*>(digests + row * 4);
  o[0] = make_ulonglong2(s[0], s[1]);
  o[1] = make_ulonglong2(s[2], s[3]);
}

// One Merkle level: out[t] = two_to_one(in[2t], in[2t+1]).
__global__ void __launch_bounds__(128, QPZK_LEAF_MINB)
k_merkle_level(const u64* __restrict__ in, u64* __restrict__ out, u64 nout) {
  u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nout) return;
  const ulonglong2* q = reinterpret_cast<const ulonglong2*>(in + t * 8);
  ulonglong2 a = q[0], b = q[1], c = q[2], d = q[3];
  u64 s[12] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y, 0, 0, 0, 0};
  poseidon_permute(s);
  ulonglong2* o = reinterpret_cast<ulonglong2*>(out + t * 4);
  o[0] = make_ulonglong2(gl_canon(s[0]), gl_canon(s[1]));
  o[1] = make_ulonglong2(gl_canon(s[2]), gl_canon(s[3]));
}

// ---- low-latency path (16 lanes per row / node) for trees and tree tops too small to fill the machine ----
// ONE launch takes a subtree from its first level to the cap: a 16-lane group computes one node, then
// bumps the arrival counter of its parent; the group that arrives second finds both children published
// and goes on to compute the parent, the first one retires. There is no grid-wide barrier and no spinning,
// so nothing has to be co-resident, and the dependent chain is exactly one permutation per level (the
// per-level launches this replaces added a launch gap per level: 16 % of the GPU time of a proof sat in
// ~1,300 such launches, profiles/r1_bench_launch_list_summary.txt). Counters are reset by the group that
// consumes them, so one small per-context array serves every tree built on the context's stream.
#ifndef QPZK_COOP_THREADS
#define QPZK_COOP_THREADS 128
#endif
#define QPZK_COOP_GROUPS (QPZK_COOP_THREADS / 16)
#define QPZK_CLIMB_MAX_START 8192  // most nodes a climb may start from (sizes the counter array: 2x this)

// FROM_LEAVES: group g hashes leaf row (first + g) [hash_or_noop] = node (0, first + g).
// otherwise  : group g computes node (l0 + 1, first + g) from its two children at level l0.
// Then it climbs while it is the second arrival, up to level `top` = log_n - cap_height.
// `levels`: level l at word offset (2N - (2N >> l)) * 4. counters: [2 * count] u32, all zero on entry and on exit.
template <bool FROM_LEAVES>
__global__ void __launch_bounds__(QPZK_COOP_THREADS)
k_tree_climb(const u64* __restrict__ src, u64 row_stride, u64 col_stride, u32 width, u64* __restrict__ levels,
             u32 log_n, u32 cap_height, u32 l0, u64 first, u64 count, u32* __restrict__ counters) {
  __shared__ u64 xch[QPZK_COOP_GROUPS][COOP_XCH_WORDS];
  const u32 g = threadIdx.x >> 4, lane = threadIdx.x & 15;
  const u32 mask = 0xffffu << (threadIdx.x & 16);  // this group's half of the warp
  const u64 gi = (u64)blockIdx.x * QPZK_COOP_GROUPS + g;
  if (gi >= count) return;  // whole groups retire together
  const u64 twoN = (u64)2 << log_n;
  const u32 top = log_n - cap_height;
  u32 l = FROM_LEAVES ? 0 : l0 + 1;
  const u32 l_start = l;
  u64 t = first + gi;
  u64 s = 0;
  if (FROM_LEAVES) {
    const u64* p = src + t * row_stride;
    if (width <= 4) {  // hash_or_noop: short rows are copied, zero padded
      if (lane < width) s = p[lane * col_stride];
    } else {
      u64 nxt = lane < 8 && lane < width ? __ldg(p + (u64)lane * col_stride) : 0;
      for (u32 off = 0; off < width; off += 8) {
        if (lane < 8 && off + lane < width) s = nxt;  // overwrite-mode absorb
        if (lane < 8 && off + 8 + lane < width) nxt = __ldg(p + (u64)(off + 8 + lane) * col_stride);  // one chunk ahead
        s = poseidon_permute_coop(s, lane, xch[g], mask);
      }
    }
  } else {
    const u64* in = levels + (twoN - (twoN >> l0) + 2 * t) * 4;
    s = lane < 8 ? in[lane] : 0;
    s = poseidon_permute_coop(s, lane, xch[g], mask);
  }
  for (;;) {
    u64* out = levels + (twoN - (twoN >> l) + t) * 4;
    QPZK_CHECK(l <= log_n && t < (((u64)1 << log_n) >> l));
    if (lane < 4) out[lane] = gl_canon(s);
#ifdef CLIMB_TL  // scripts/exp/climb_exp.cu: latest completion time per level
    if (lane == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      atomicMax(&g_tl[l], now);
    }
#endif
    if (l >= top) break;
    // publish, then pair up with the sibling subtree
    __threadfence();
    __syncwarp(mask);
    const u32 d = l + 1 - l_start;                                   // parent's depth above the start level
    const u64 ci = 2 * count - ((2 * count) >> d) + ((t >> 1) - (first >> d));
    u32 old = 0;
    QPZK_CHECK(ci < 2 * count && ci < 2 * QPZK_CLIMB_MAX_START);
    if (lane == 0) old = atomicAdd(&counters[ci], 1u);
    QPZK_CHECK(lane != 0 || old < 2);
    old = __shfl_sync(mask, old, 0, 16);
    if (old == 0) break;                                             // the sibling will take the parent
    if (lane == 0) counters[ci] = 0;
    __threadfence();
    l++;
    t >>= 1;
    const u64* in = levels + (twoN - (twoN >> (l - 1)) + 2 * t) * 4;
    s = lane < 8 ? __ldcg(in + lane) : 0;                            // L2: the sibling was written by another SM
#ifdef CLIMB_TL
    unsigned long long tp0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tp0));
#endif
    s = poseidon_permute_coop(s, lane, xch[g], mask);
#ifdef CLIMB_TL
    if (lane == 0) {
      unsigned long long tp1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tp1));
      atomicMax(&g_pmax[l], tp1 - tp0);
      atomicAdd(&g_psum[l], tp1 - tp0);
    }
#endif
  }
}

// Generic batched sponge / permutation entry points (KATs, host API).
__global__ void __launch_bounds__(128) k_permute(u64* states, u64 n) {
  u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = states[t * 12 + i];
  poseidon_permute(s);
#pragma unroll
  for (int i = 0; i < 12; i++) states[t * 12 + i] = gl_canon(s[i]);
}

__global__ void __launch_bounds__(128)
k_hash_no_pad(const u64* __restrict__ in, u64 n, u32 len, u64* __restrict__ out) {
  u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const u64* p = in + t * len;
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = 0;
  for (u32 off = 0; off < len; off += 8) {
    u32 l = len - off < 8 ? len - off : 8;
#pragma unroll
    for (int i = 0; i < 8; i++)
      if (i < l) s[i] = p[off + i];
    poseidon_permute(s);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) out[t * 4 + i] = gl_canon(s[i]);
}

// plonky2 `digests` layout: per cap subtree with L layers, the sibling pair q of layer i sits at
// 2*((q << (i+1)) + (1 << i) - 1) + {0,1}. `levels` = level-major buffer described above.
__global__ void k_export_digests(const u64* __restrict__ levels, u32 log_n, u32 cap_height,
                                 u64* __restrict__ out) {
  u32 L = log_n - cap_height;
  u64 total = ((u64)2 << log_n) - ((u64)2 << cap_height);  // digests below the cap
  u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  // locate level i: level i starts at offset 2N - (2N >> i)
  u64 twoN = (u64)2 << log_n;
  u32 i = 0;
  while (g >= twoN - (twoN >> (i + 1))) i++;
  u64 j = g - (twoN - (twoN >> i));           // index within level i
  u32 per_sub_bits = L - i;                   // log2(digests of this level per subtree)
  u64 sub = j >> per_sub_bits, jj = j & (((u64)1 << per_sub_bits) - 1);
  u64 pq = jj >> 1, parity = jj & 1;
  u64 idx = 2 * ((pq << (i + 1)) + ((u64)1 << i) - 1) + parity;
  u64 sub_len = ((u64)2 << L) - 2;
  const ulonglong2* s = reinterpret_cast<const ulonglong2*>(levels + g * 4);
  ulonglong2* d = reinterpret_cast<ulonglong2*>(out + (sub * sub_len + idx) * 4);
  d[0] = s[0];
  d[1] = s[1];
}

// MerkleTree::prove gather: siblings[l] = level_l[(leaf >> l) ^ 1] for l < L.
__global__ void k_gather_path(const u64* __restrict__ levels, u32 log_n, u32 cap_height, u64 leaf,
                              u64* __restrict__ out) {
  u32 L = log_n - cap_height;
  u32 l = threadIdx.x >> 2, e = threadIdx.x & 3;
  if (l >= L) return;
  u64 twoN = (u64)2 << log_n;
  u64 off = twoN - (twoN >> l);
  out[l * 4 + e] = levels[(off + ((leaf >> l) ^ 1)) * 4 + e];
}

}  // namespace qpzk
