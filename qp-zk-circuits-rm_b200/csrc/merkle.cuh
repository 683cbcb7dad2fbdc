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

// ---- low-latency variants (16 lanes per row / node) for trees too small to fill the machine ----
#define QPZK_COOP_THREADS 128
#define QPZK_COOP_GROUPS (QPZK_COOP_THREADS / 16)
__global__ void __launch_bounds__(QPZK_COOP_THREADS)
k_leaf_hash_coop(const u64* __restrict__ src, u64 row_stride, u64 col_stride, u32 width, u64 nrows,
                 u64* __restrict__ digests) {
  __shared__ u64 xch[QPZK_COOP_GROUPS][24];
  const u32 g = threadIdx.x >> 4, lane = threadIdx.x & 15;
  u64 row = (u64)blockIdx.x * QPZK_COOP_GROUPS + g;
  const bool live = row < nrows;
  if (!live) row = nrows - 1;  // keep the whole warp in the exchange; results are discarded
  const u64* p = src + row * row_stride;
  u64 s = 0;
  if (width <= 4) {  // hash_or_noop: short rows are copied, zero padded
    if (lane < width) s = p[lane * col_stride];
  } else {
    for (u32 off = 0; off < width; off += 8) {
      if (lane < 8 && off + lane < width) s = __ldg(p + (u64)(off + lane) * col_stride);  // overwrite-mode absorb
      s = poseidon_permute_coop(s, lane, xch[g]);
    }
  }
  if (live && lane < 4) digests[row * 4 + lane] = gl_canon(s);
}

__global__ void __launch_bounds__(QPZK_COOP_THREADS)
k_merkle_level_coop(const u64* __restrict__ in, u64* __restrict__ out, u64 nout) {
  __shared__ u64 xch[QPZK_COOP_GROUPS][24];
  const u32 g = threadIdx.x >> 4, lane = threadIdx.x & 15;
  u64 t = (u64)blockIdx.x * QPZK_COOP_GROUPS + g;
  const bool live = t < nout;
  if (!live) t = nout - 1;
  u64 s = lane < 8 ? in[t * 8 + lane] : 0;
  s = poseidon_permute_coop(s, lane, xch[g]);
  if (live && lane < 4) out[t * 4 + lane] = gl_canon(s);
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
