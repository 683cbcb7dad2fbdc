// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// `MerkleTree::new(leaves, cap_height)`, `MerkleTree::prove`, `verify_merkle_proof_to_cap` and
// `PolynomialBatch::from_values / from_coeffs / get_lde_values` of qp-plonky2 1.1.1
// (/root/reference/Cargo.lock:489-490), restated per SURVEY.md §8(a) H4-H7 and App. A.5.
// Reference call sites that land here: /root/reference/wormhole/prover/src/lib.rs:233-237,
// /root/reference/wormhole/circuit/src/circuit.rs:98-108,
// /root/reference/wormhole/aggregator/src/circuits/tree.rs:127,136, /root/reference/voting/src/lib.rs:355-356.
#pragma once
#include <atomic>
#include <functional>
#include <thread>
#include <vector>

#include "ntt.hpp"
#include "poseidon.hpp"
#include "poseidon_avx512.hpp"

namespace orc {

struct MerkleTree {
  size_t nleaves = 0, leaf_len = 0;
  unsigned cap_height = 0;
  std::vector<u64> leaves;     // row-major [nleaves][leaf_len]
  std::vector<Hash> digests;   // plonky2's interleaved layout, 2*(nleaves - 2^cap_height)
  std::vector<Hash> cap;       // 2^cap_height
  const u64* leaf(size_t i) const { return &leaves[i * leaf_len]; }
};

// Recursive fill: buffer halves = left/right child subtrees; the left child's digest is stored in
// the LAST slot of the left half and the right child's in the FIRST slot of the right half.
static inline Hash fill_subtree(Hash* buf, size_t buflen, const u64* leaves, size_t nleaves,
                                size_t leaf_len) {
  if (buflen == 0) return hash_or_noop(leaves, leaf_len);
  size_t half = buflen / 2;
  Hash l = fill_subtree(buf, half - 1, leaves, nleaves / 2, leaf_len);
  Hash r = fill_subtree(buf + half + 1, half - 1, leaves + (nleaves / 2) * leaf_len, nleaves / 2,
                        leaf_len);
  buf[half - 1] = l;
  buf[half] = r;
  return two_to_one(l, r);
}

static inline void parallel_for(size_t n, unsigned threads, const std::function<void(size_t)>& f);

// The same subtree level by level, eight permutations per call (poseidon_avx512.hpp): leaf digests, then each
// level's nodes, every digest stored where the recursive rule above puts it - sibling pair q of level i at
// 2 * ((q << (i + 1)) + (1 << i) - 1) + {0, 1}. Used when the CPU has AVX-512 (run-time check); results are
// identical to fill_subtree.
static inline Hash fill_subtree_x8(Hash* buf, size_t buflen, const u64* leaves, size_t nleaves, size_t leaf_len) {
  if (nleaves < 8 || !have_avx512()) return fill_subtree(buf, buflen, leaves, nleaves, leaf_len);
  std::vector<Hash> cur(nleaves), next;
  for (size_t j = 0; j < nleaves; j += 8) hash_or_noop_x8(leaves + j * leaf_len, leaf_len, leaf_len, &cur[j]);
  for (unsigned i = 0; cur.size() > 1; i++) {
    const size_t count = cur.size();
    for (size_t j = 0; j < count; j++) buf[2 * (((j >> 1) << (i + 1)) + ((size_t)1 << i) - 1) + (j & 1)] = cur[j];
    next.resize(count / 2);
    size_t j = 0;
    for (; j + 8 <= count / 2; j += 8) two_to_one_x8(&cur[2 * j], &next[j]);
    for (; j < count / 2; j++) next[j] = two_to_one(cur[2 * j], cur[2 * j + 1]);
    cur.swap(next);
  }
  (void)buflen;
  return cur[0];
}

static inline MerkleTree merkle_new(std::vector<u64> leaves, size_t nleaves, size_t leaf_len,
                                    unsigned cap_height, unsigned threads = 1) {
  MerkleTree t;
  t.nleaves = nleaves;
  t.leaf_len = leaf_len;
  t.cap_height = cap_height;
  t.leaves = std::move(leaves);
  size_t ncap = (size_t)1 << cap_height;
  t.digests.resize(2 * (nleaves - ncap));
  t.cap.resize(ncap);
  size_t sub_leaves = nleaves >> cap_height, sub_digests = t.digests.size() >> cap_height;
  // split each cap subtree further so that `threads` workers have something to do
  unsigned split = 0;
  while (((size_t)1 << (cap_height + split)) < 4 * (size_t)threads && (sub_leaves >> split) > 1)
    split++;
  if (threads <= 1) split = 0;
  size_t parts = (size_t)1 << split;
  if (split == 0) {
    parallel_for(ncap, threads, [&](size_t s) {
      t.cap[s] = fill_subtree_x8(t.digests.data() + s * sub_digests, sub_digests,
                                 t.leaves.data() + s * sub_leaves * leaf_len, sub_leaves, leaf_len);
    });
    return t;
  }
  // Work on sub-subtrees in parallel, then finish the top `split` levels of each cap subtree
  // serially with the same layout rule.
  struct Job { Hash* buf; size_t buflen; const u64* lv; size_t nl; Hash out; };
  std::vector<Job> jobs;
  std::vector<std::vector<size_t>> tops(ncap);
  std::function<void(size_t, Hash*, size_t, const u64*, size_t, unsigned)> plan =
      [&](size_t s, Hash* buf, size_t buflen, const u64* lv, size_t nl, unsigned depth) {
        if (depth == split) {
          tops[s].push_back(jobs.size());
          jobs.push_back(Job{buf, buflen, lv, nl, Hash{}});
          return;
        }
        size_t half = buflen / 2;
        plan(s, buf, half - 1, lv, nl / 2, depth + 1);
        plan(s, buf + half + 1, half - 1, lv + (nl / 2) * leaf_len, nl / 2, depth + 1);
      };
  for (size_t s = 0; s < ncap; s++)
    plan(s, t.digests.data() + s * sub_digests, sub_digests,
         t.leaves.data() + s * sub_leaves * leaf_len, sub_leaves, 0);
  parallel_for(jobs.size(), threads, [&](size_t j) {
    jobs[j].out = fill_subtree_x8(jobs[j].buf, jobs[j].buflen, jobs[j].lv, jobs[j].nl, leaf_len);
  });
  std::function<Hash(size_t, size_t&, Hash*, size_t, unsigned)> finish =
      [&](size_t s, size_t& next, Hash* buf, size_t buflen, unsigned depth) -> Hash {
    if (depth == split) return jobs[tops[s][next++]].out;
    size_t half = buflen / 2;
    Hash l = finish(s, next, buf, half - 1, depth + 1);
    Hash r = finish(s, next, buf + half + 1, half - 1, depth + 1);
    buf[half - 1] = l;
    buf[half] = r;
    return two_to_one(l, r);
  };
  (void)parts;
  for (size_t s = 0; s < ncap; s++) {
    size_t next = 0;
    t.cap[s] = finish(s, next, t.digests.data() + s * sub_digests, sub_digests, 0);
  }
  return t;
}

// MerkleTree::prove: siblings bottom-up, read out of the interleaved digest buffer.
static inline std::vector<Hash> merkle_prove(const MerkleTree& t, size_t leaf_index) {
  unsigned num_layers = log2_strict(t.nleaves) - t.cap_height;
  size_t tree_len = t.digests.size() >> t.cap_height;
  const Hash* tree = t.digests.data() + tree_len * (leaf_index >> num_layers);
  size_t pair_index = leaf_index & (((size_t)1 << num_layers) - 1);
  std::vector<Hash> sib(num_layers);
  for (unsigned i = 0; i < num_layers; i++) {
    size_t parity = pair_index & 1;
    pair_index >>= 1;
    size_t siblings_index = (pair_index << (i + 1)) + ((size_t)1 << i) - 1;
    sib[i] = tree[2 * siblings_index + (1 - parity)];
  }
  return sib;
}

// verify_merkle_proof_to_cap
static inline bool merkle_verify(const u64* leaf, size_t leaf_len, size_t leaf_index,
                                 const Hash* cap, const Hash* siblings, size_t nsib) {
  Hash h = hash_or_noop(leaf, leaf_len);
  size_t idx = leaf_index;
  for (size_t i = 0; i < nsib; i++) {
    h = (idx & 1) ? two_to_one(siblings[i], h) : two_to_one(h, siblings[i]);
    idx >>= 1;
  }
  return h == cap[idx];
}

static inline void parallel_for(size_t n, unsigned threads, const std::function<void(size_t)>& f) {
  if (threads <= 1 || n <= 1) {
    for (size_t i = 0; i < n; i++) f(i);
    return;
  }
  std::vector<std::thread> th;
  std::atomic<size_t> next{0};
  for (unsigned t = 0; t < threads; t++)
    th.emplace_back([&] {
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= n) break;
        f(i);
      }
    });
  for (auto& x : th) x.join();
}

// PolynomialBatch. Column-major polynomials, row-major bit-reversed leaves.
struct PolyBatch {
  unsigned degree_log = 0, rate_bits = 0, salt_cols = 0;
  size_t ncols = 0;
  std::vector<std::vector<u64>> coeffs;  // [ncols][n]
  MerkleTree tree;                       // leaves [n << rate_bits][ncols + salt_cols]
};

// salts: NULL, or [salt_cols][n << rate_bits] column-major LDE-domain values in NATURAL order
// (the reference draws them from an OS RNG; injected here so comparisons are reproducible).
static inline PolyBatch batch_from_coeffs(std::vector<std::vector<u64>> coeffs, unsigned rate_bits,
                                          unsigned cap_height, const u64* salts,
                                          unsigned salt_cols, unsigned threads = 1) {
  PolyBatch b;
  b.ncols = coeffs.size();
  size_t n = coeffs.empty() ? 0 : coeffs[0].size();
  b.degree_log = log2_strict(n);
  b.rate_bits = rate_bits;
  b.salt_cols = salts ? salt_cols : 0;
  size_t N = n << rate_bits;
  unsigned lb = b.degree_log + rate_bits;
  size_t width = b.ncols + b.salt_cols;
  std::vector<u64> leaves(N * width);
  parallel_for(b.ncols, threads, [&](size_t c) {
    std::vector<u64> v = coset_fft(lde(coeffs[c], rate_bits), GEN);
    for (size_t i = 0; i < N; i++) leaves[bitrev(i, lb) * width + c] = v[i];
  });
  parallel_for(b.salt_cols, threads, [&](size_t s) {
    for (size_t i = 0; i < N; i++) leaves[bitrev(i, lb) * width + b.ncols + s] = salts[s * N + i];
  });
  b.coeffs = std::move(coeffs);
  b.tree = merkle_new(std::move(leaves), N, width, cap_height, threads);
  return b;
}

static inline PolyBatch batch_from_values(const std::vector<std::vector<u64>>& values,
                                          unsigned rate_bits, unsigned cap_height,
                                          const u64* salts, unsigned salt_cols,
                                          unsigned threads = 1) {
  std::vector<std::vector<u64>> coeffs(values.size());
  parallel_for(values.size(), threads, [&](size_t c) { coeffs[c] = ifft(values[c]); });
  return batch_from_coeffs(std::move(coeffs), rate_bits, cap_height, salts, salt_cols, threads);
}

// get_lde_values(index, step) = leaves[rev(index*step)][.. len - salt]
static inline const u64* batch_lde_row(const PolyBatch& b, size_t index, size_t step) {
  unsigned lb = b.degree_log + b.rate_bits;
  return b.tree.leaf(bitrev(index * step, lb));
}

}  // namespace orc
