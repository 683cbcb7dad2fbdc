// libqpzk: C ABI (include/qpzk.h) over the sm_100a kernels. Unity build: the kernels share the
// Poseidon tables in __constant__ memory, so everything is one translation unit.
//
// Host-side orchestration that replaces the bodies of `PolynomialBatch::from_values /
// from_coeffs`, `MerkleTree::new / prove` (qp-plonky2 1.1.1 fri/oracle.rs, hash/merkle_tree.rs),
// reached from /root/reference/wormhole/prover/src/lib.rs:233-237 and
// /root/reference/wormhole/circuit/src/circuit.rs:98-108. There is no CPU fallback anywhere in
// this file: every entry point either runs the CUDA path or returns an error.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sched.h>
#include <map>
#include <memory>
#include <tuple>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/qpzk.h"
#include "gl.cuh"
#include "merkle.cuh"
#include "ntt.cuh"
#include "poseidon.cuh"
#include "poseidon_tables.hpp"
#include "poseidon_upload.cuh"
#include "prover.cuh"

using namespace qpzk;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(x)                                                                               \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess)                                                                  \
      return fail(e_ == cudaErrorMemoryAllocation ? QPZK_ERR_OOM : QPZK_ERR_CUDA,           \
                  std::string(#x) + ": " + cudaGetErrorString(e_));                         \
  } while (0)
#define QP(x)             \
  do {                    \
    int r_ = (x);         \
    if (r_ != QPZK_OK) return r_; \
  } while (0)
// Nothing may unwind across the C ABI (include/qpzk.h): every extern "C" body that can allocate runs inside
// this guard and turns an exception into a status code.
template <class F>
static int guarded(F&& f) noexcept {
  try {
    return f();
  } catch (const std::bad_alloc&) {
    return fail(QPZK_ERR_OOM, "out of host memory");
  } catch (const std::exception& e) {
    return fail(QPZK_ERR_CUDA, std::string("internal error: ") + e.what());
  } catch (...) {
    return fail(QPZK_ERR_CUDA, "internal error");
  }
}

static inline u64 glh_bitrev(u64 x, u32 bits) {
  u64 r = 0;
  for (u32 i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}

struct TabKey {
  int k;
  bool inverse;
  bool operator<(const TabKey& o) const { return k != o.k ? k < o.k : inverse < o.inverse; }
};

struct qpzk_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool blocking_sync = false;      // QPZK_CTX_BLOCKING_SYNC: sleep instead of spinning while the device works
  bool yield_sync = false;         // QPZK_CTX_YIELD_SYNC: poll + sched_yield()
  cudaEvent_t sync_ev = nullptr;
  cudaEvent_t ev[QPZK_NUM_STAGES + 1];
  float stage_ms[QPZK_NUM_STAGES] = {0};
  uint64_t launches = 0;
  int sm_count = 148;
  std::map<TabKey, std::pair<u64*, u64*>> root_tabs;   // (lo, hi)
  std::map<std::tuple<int, int, u64>, u64*> coset_pm;    // (k, r, shift) -> pm[2^r][2^k]
  std::map<std::pair<u64, int>, std::pair<u64*, u64*>> pow_tabs;  // (base, k) -> two-level base^e table
  std::map<std::tuple<int, int, bool>, u64*> tw_mats;    // (k, a, inverse) -> twiddle matrix [2^a][2^(k-a)]
  u64* scratch_path = nullptr;                           // small device scratch for openings
  u32* climb_counters = nullptr;                         // k_tree_climb arrival counters (all zero between launches)
  u64 coop_max = 1024;                                   // most permutations per step for the 16-lane kernels
  cudaMemPool_t pool = nullptr;                          // the library's own stream-ordered pool on this device
  // uploads from pageable host memory (h2d_copy): two pinned staging buffers filled by h2d_threads host threads
  void* h2d_arena = nullptr;       // [thread][2] chunks of kH2DChunk bytes
  cudaEvent_t h2d_ev[32] = {};     // one per chunk: the copy engine has read it
  int h2d_threads = 4;
  size_t h2d_min = (size_t)32 << 20;
};

static void dev_free(qpzk_ctx* c, void* p);

struct qpzk_batch {
  qpzk_ctx* ctx = nullptr;
  uint32_t ncols = 0, salt_cols = 0, degree_bits = 0, rate_bits = 0, cap_height = 0;
  u64* coeffs = nullptr;   // [ncols][n]
  // LDE values, column-major in bit-reversed row order: element (column c, leaf L) at lde[c * lde_stride + L].
  // A whole batch holds [ncols+salt_cols][N]. A multi-GPU shard allocates only its own leaves [leaf0, leaf1):
  // lde_stride = leaf1 - leaf0 and `lde` is the allocation shifted back by leaf0, so kernels keep addressing
  // by absolute leaf index and there is no foreign row to serve by mistake.
  u64* lde_alloc = nullptr;
  u64* lde = nullptr;
  u64 lde_stride = 0;
  u64* levels = nullptr;   // digest levels, 2N*4 u64 (a shard's foreign digests and cap entries are zero)
  u64 leaf0 = 0, leaf1 = 0;  // the leaves this batch holds: everything, or one rank's shard of whole cap subtrees
  uint32_t log_N() const { return degree_bits + rate_bits; }
  uint32_t width() const { return ncols + salt_cols; }
  bool sharded() const { return leaf0 != 0 || leaf1 != ((u64)1 << log_N()); }
  bool owns(u64 leaf) const { return leaf >= leaf0 && leaf < leaf1; }
  ~qpzk_batch() {
    if (!ctx) return;
    dev_free(ctx, coeffs);
    dev_free(ctx, lde_alloc);
    dev_free(ctx, levels);
  }
};

struct qpzk_tree {
  qpzk_ctx* ctx = nullptr;
  uint32_t log_n = 0, cap_height = 0, leaf_len = 0;
  u64* levels = nullptr;
  ~qpzk_tree() {
    if (ctx) dev_free(ctx, levels);
  }
};

// ------------------------------------------------------------------------------------------
// Wait for the context's stream. cudaStreamSynchronize spins on a host core, which is the lowest
// latency but oversubscribes the host once contexts x processes exceed the cores (measured: 8 ranks x 6
// proving threads on 32 cores lost 10 % of throughput); a context created with
// QPZK_CTX_BLOCKING_SYNC waits on a blocking-sync event instead.
static cudaError_t ctx_wait(qpzk_ctx* c) {
  if (!c->blocking_sync && !c->yield_sync) return cudaStreamSynchronize(c->stream);
  cudaError_t e = cudaEventRecord(c->sync_ev, c->stream);
  if (e != cudaSuccess) return e;
  if (c->blocking_sync) return cudaEventSynchronize(c->sync_ev);
  while ((e = cudaEventQuery(c->sync_ev)) == cudaErrorNotReady) sched_yield();
  return e;
}

static int dev_alloc(qpzk_ctx* c, size_t bytes, u64** out) {
  void* p = nullptr;
  CU(cudaMallocFromPoolAsync(&p, bytes ? bytes : 8, c->pool, c->stream));
  *out = (u64*)p;
  return QPZK_OK;
}
static void dev_free(qpzk_ctx* c, void* p) {
  if (p) cudaFreeAsync(p, c->stream);
}
struct DevBuf {  // scoped stream-ordered allocation
  qpzk_ctx* c;
  u64* p = nullptr;
  explicit DevBuf(qpzk_ctx* c_) : c(c_) {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { dev_free(c, p); }
  int alloc(size_t bytes) { return dev_alloc(c, bytes, &p); }
};

static int get_root_tab(qpzk_ctx* c, int k, bool inverse, RootTab* out) {
  TabKey key{k, inverse};
  auto it = c->root_tabs.find(key);
  int lk = (k + 1) / 2;
  if (it == c->root_tabs.end()) {
    u64 root = glh::root_of_unity(k);
    if (inverse) root = glh::inv(root);
    u64 *lo, *hi;
    QP(dev_alloc(c, sizeof(u64) << lk, &lo));
    QP(dev_alloc(c, sizeof(u64) << (k - lk), &hi));
    u64 cnt = (u64)1 << lk;
    k_build_root_tab<<<(unsigned)((cnt + 127) / 128), 128, 0, c->stream>>>(lo, hi, root, k, lk);
    c->launches++;
    CU(cudaGetLastError());
    it = c->root_tabs.emplace(key, std::make_pair(lo, hi)).first;
  }
  out->lo = it->second.first;
  out->hi = it->second.second;
  out->k = k;
  out->lk = lk;
  return QPZK_OK;
}

// Two-level table of base^e, e < 2^k, for an arbitrary base (coset shift removal).
static int get_pow_tab(qpzk_ctx* c, u64 base, int k, RootTab* out) {
  auto key = std::make_pair(base, k);
  auto it = c->pow_tabs.find(key);
  int lk = (k + 1) / 2;
  if (it == c->pow_tabs.end()) {
    u64 *lo, *hi;
    QP(dev_alloc(c, sizeof(u64) << lk, &lo));
    QP(dev_alloc(c, sizeof(u64) << (k - lk), &hi));
    u64 cnt = (u64)1 << lk;
    k_build_root_tab<<<(unsigned)((cnt + 127) / 128), 128, 0, c->stream>>>(lo, hi, base, k, lk);
    c->launches++;
    CU(cudaGetLastError());
    it = c->pow_tabs.emplace(key, std::make_pair(lo, hi)).first;
  }
  out->lo = it->second.first;
  out->hi = it->second.second;
  out->k = k;
  out->lk = lk;
  return QPZK_OK;
}

// M[k1][j] = root^(j*k1) (k_build_tw_matrix): the twiddles between the first 2^a-point stage of a
// 2^k-point transform and the rest.
static int get_tw_matrix(qpzk_ctx* c, int k, int a, bool inverse, const u64** out) {
  auto key = std::make_tuple(k, a, inverse);
  auto it = c->tw_mats.find(key);
  if (it == c->tw_mats.end()) {
    RootTab tab;
    QP(get_root_tab(c, k, inverse, &tab));
    u64* m;
    u64 cnt = (u64)1 << k;
    QP(dev_alloc(c, cnt * sizeof(u64), &m));
    k_build_tw_matrix<<<(unsigned)((cnt + 255) / 256), 256, 0, c->stream>>>(m, tab, k, a);
    c->launches++;
    CU(cudaGetLastError());
    it = c->tw_mats.emplace(key, m).first;
  }
  *out = it->second;
  return QPZK_OK;
}

static int get_coset_pm(qpzk_ctx* c, int k, int r, u64 shift, const u64** out) {
  auto key = std::make_tuple(k, r, shift);
  auto it = c->coset_pm.find(key);
  if (it == c->coset_pm.end()) {
    u64* pm;
    u64 cnt = (u64)1 << (k + r);
    QP(dev_alloc(c, cnt * sizeof(u64), &pm));
    u64 wN = glh::root_of_unity(k + r);
    k_build_coset_pm<<<(unsigned)((cnt + 255) / 256), 256, 0, c->stream>>>(pm, shift, wN, k, r);
    c->launches++;
    CU(cudaGetLastError());
    it = c->coset_pm.emplace(key, pm).first;
  }
  *out = it->second;
  return QPZK_OK;
}

// Batched n-point transforms. flavour LDE: src natural -> dst DIF order per coset (ncosets = 2^r,
// coset pre-multipliers applied). flavour IFFT: src natural values -> dst natural coefficients.
// Up to 2^14 points (every wormhole / voting / aggregation-node proof) one CTA transforms a whole
// (column, coset) in shared memory; above that, two passes.
static const int kSmallMaxLog = 14;
static const int kSmallTableLog = 12;  // up to here the full twiddle table fits next to the tile
static const int kSmallSmemBytes = 148 * 1024;
static const int kClusterMaxLog = 16;  // 8 CTAs x 2 rows of 2^12 points
static int ntt_cluster_min() {  // QPZK_NTT_CLUSTER_MIN=13|14: also take 2^13 / 2^14-point transforms (experiment)
  static const int v = [] {
    const char* e = getenv("QPZK_NTT_CLUSTER_MIN");
    int x = e ? atoi(e) : 15;
    return x < 9 ? 9 : x;
  }();
  return v;
}
static bool ntt_cluster_enabled() {    // QPZK_NTT_CLUSTER=0: the two-pass kernels (A/B measurements)
  static const bool on = [] {
    const char* e = getenv("QPZK_NTT_CLUSTER");
    return !(e && e[0] == '0');
  }();
  return on;
}

// (Tried and dropped: running the two passes of a long transform per group of columns small enough for
// pass A's output to stay in the 126 MB L2 until pass B overwrites it in place. The kernels are
// integer-issue-bound, not HBM-bound, so the saved DRAM round trip bought nothing and the extra launch
// boundaries cost 11 % (48 MB groups) to 47 % (16 MB groups) of the LDE time at 2^16 x 135.)

template <bool NATURAL_OUT>
static int launch_small(qpzk_ctx* c, const u64* src, u64 src_stride, u64* dst, u64 dst_stride, const u64* pm,
                        RootTab tab, u32 ncols, u32 ncosets, int k, int r, u64 scale, u32 blk0) {
  const u64* tw1 = nullptr;
  if (k > kSmallTableLog) QP(get_tw_matrix(c, k, 4, NATURAL_OUT, &tw1));
  u32 n = 1u << k;
  size_t smem = ((size_t)tile_pitch(n) + (tw1 ? n >> 4 : n)) * 8;
  if (k == 14)  // one CTA per SM: 1024 threads (64 registers, ~250 B of spills) keep 32 warps resident; 4 % faster than 512
    k_ntt_small<NATURAL_OUT, 1024><<<dim3(ncols, ncosets), 1024, smem, c->stream>>>(src, src_stride, dst, dst_stride, pm,
                                                                                  tab, tw1, k, r, scale, blk0);
  else
    k_ntt_small<NATURAL_OUT, 256><<<dim3(ncols, ncosets), 256, smem, c->stream>>>(src, src_stride, dst, dst_stride, pm,
                                                                                  tab, tw1, k, r, scale, blk0);
  c->launches++;
  CU(cudaGetLastError());
  return QPZK_OK;
}

// blk0 / nblk: which of the 2^r leaf blocks (n bit-reversed leaves each, block b = coset rev_r(b)) to
// evaluate; the full commit passes (0, 2^r), a multi-GPU shard its own range.
static int launch_lde_shift(qpzk_ctx* c, const u64* coeffs, u64 src_stride, u64* lde, u64 dst_stride,
                            u32 ncols, int k, int r, u64 shift, u32 blk0 = 0, u32 nblk = 0) {
  if (ncols == 0) return QPZK_OK;
  if (nblk == 0) nblk = (1u << r) - blk0;
  RootTab tab;
  QP(get_root_tab(c, k, false, &tab));
  const u64* pm;
  QP(get_coset_pm(c, k, r, shift, &pm));
  u32 ncosets = nblk;
  if (k <= kSmallMaxLog && k < ntt_cluster_min()) return launch_small<false>(c, coeffs, src_stride, lde, dst_stride, pm, tab, ncols, ncosets, k, r, 1, blk0);
  if (k > 20) return fail(QPZK_ERR_UNSUPPORTED, "degree_bits > 20 not supported");
  if (k <= kClusterMaxLog && ntt_cluster_enabled()) {  // 2^15, 2^16 points: one pass, the tile spread over a cluster
    const int a1 = k == 15 ? 3 : 4, lb = k - a1;       // n = 2^a1 rows of 2^lb points: radix-8 / radix-16 across the cluster
    RootTab tab_b;
    QP(get_root_tab(c, lb, false, &tab_b));
    const u64 *tw1, *twc;
    QP(get_tw_matrix(c, lb, 4, false, &tw1));
    QP(get_tw_matrix(c, k, a1, false, &twc));
    const u32 B = 1u << lb;
    const size_t smem = ((size_t)tile_pitch(B) * ((1u << a1) / QPZK_NTT_CLUSTER) + (B >> 4)) * 8;
    const dim3 grid(QPZK_NTT_CLUSTER, ncosets, ncols);
    if (a1 == 4)
      k_ntt_cluster<4, 256><<<grid, 256, smem, c->stream>>>(coeffs, src_stride, lde, dst_stride, pm, tab_b, tw1, twc, k, r, blk0);
    else
      k_ntt_cluster<3, 256><<<grid, 256, smem, c->stream>>>(coeffs, src_stride, lde, dst_stride, pm, tab_b, tw1, twc, k, r, blk0);
    c->launches++;
    CU(cudaGetLastError());
    return QPZK_OK;
  }
  int a = (k + 1) / 2;
  if (a > 8) a = 8;
  int b = k - a;
  const u64* twm;
  QP(get_tw_matrix(c, k, a, false, &twm));
  u32 cols_log = 4, cols = 16;
  size_t smem_a = ((size_t)(1u << a) * cols + (1u << a)) * 8;
  int rows_log = 12 - b;  // kTileElems / n2
  if (rows_log < 0) rows_log = 0;
  if (rows_log > a) rows_log = a;
  u32 rows = 1u << rows_log;
  size_t smem_b = ((size_t)rows * tile_pitch(1u << b) + (1u << b)) * 8;
  k_ntt_pass_a<true><<<dim3((1u << b) / cols, ncosets, ncols), 256, smem_a, c->stream>>>(
      coeffs, src_stride, lde, dst_stride, pm, tab, twm, k, a, r, cols_log, blk0);
  k_ntt_pass_b_rows<<<dim3((1u << a) / rows, ncols, ncosets), 256, smem_b, c->stream>>>(lde, dst_stride, tab, k, a, r,
                                                                                       (u32)rows_log, blk0);
  c->launches += 2;
  CU(cudaGetLastError());
  return QPZK_OK;
}

static int launch_lde(qpzk_ctx* c, const u64* coeffs, u64 src_stride, u64* lde, u64 dst_stride, u32 ncols,
                      int k, int r, u32 blk0 = 0, u32 nblk = 0) {
  return launch_lde_shift(c, coeffs, src_stride, lde, dst_stride, ncols, k, r, GL_GEN, blk0, nblk);
}

static int launch_ifft(qpzk_ctx* c, const u64* values, u64 src_stride, u64* coeffs, u64 dst_stride,
                       u32 ncols, int k) {
  if (ncols == 0) return QPZK_OK;
  RootTab tab;
  QP(get_root_tab(c, k, true, &tab));
  u64 ninv = glh::inv(((u64)1 << k) % GL_P);
  if (k <= kSmallMaxLog) return launch_small<true>(c, values, src_stride, coeffs, dst_stride, nullptr, tab, ncols, 1, k, 0, ninv, 0);
  if (k > 22) return fail(QPZK_ERR_UNSUPPORTED, "ifft: more than 2^22 points not supported");
  int a = (k + 1) / 2, b = k - a;
  const u64* twm;
  QP(get_tw_matrix(c, k, a, true, &twm));
  // tiles of 4096 elements: [2^a][cols] in pass A, [rc][2^b] in pass B
  u32 cols_log = a <= 8 ? 4 : 12 - a, cols = 1u << cols_log;
  u32 rc_log = b <= 8 ? 4 : 12 - b, rc = 1u << rc_log;
  u64* tmp;
  QP(dev_alloc(c, (size_t)ncols << (k + 3), &tmp));
  size_t smem_a = ((size_t)(1u << a) * cols + (1u << a)) * 8;
  size_t smem_b = ((size_t)rc * tile_pitch(1u << b) + (1u << b)) * 8;
  k_ntt_pass_a<false><<<dim3((1u << b) / cols, 1, ncols), 256, smem_a, c->stream>>>(
      values, src_stride, tmp, (u64)1 << k, nullptr, tab, twm, k, a, 0, cols_log, 0);
  k_ntt_pass_b_transpose<<<dim3((1u << a) / rc, ncols), 256, smem_b, c->stream>>>(tmp, (u64)1 << k, coeffs, dst_stride,
                                                                                  tab, k, a, rc_log, ninv);
  c->launches += 2;
  dev_free(c, tmp);
  CU(cudaGetLastError());
  return QPZK_OK;
}

// Leaf digests + all levels down to the cap. Element (row, col) at src[row*rs + col*cs].
// Above `coop_max` independent permutations a step runs one thread per permutation (throughput); at or below
// it the rest of the tree - leaves included, if there are that few - is ONE k_tree_climb launch on the 16-lane
// path. The 16-lane kernels spend ~3.8x the lane-instructions per permutation and a scheduler saturates at about
// two of their warps (scripts/exp/coop_lat.cu: 10.2 us per step at one warp per scheduler, 21.3 us at 3.5), and in
// a climb the LATER sibling goes on with the parent, so the groups that survive are the ones on the busiest SMs:
// started from 4096 nodes, levels 2-5 take 27 us each instead of 10.7 (scripts/exp/climb_exp.cu). Starting the
// climb at 1024 nodes and giving the levels above it to k_merkle_level is both faster and cheaper (2^14 ZK
// proof, 8 streams; QPZK_COOP_MAX, read once per process, overrides):
//   4096 -> 206.8 proofs/s, 8.14 ms single proof, 3.08 ms voting-sized proof
//   2048 -> 211.6 / 7.96 / 3.00      1024 -> 215.3 / 7.94 / 2.97      512 -> 215.5 / 8.01 / 3.08
// [leaf0, leaf0 + nleaves) restricts the work to a range of whole cap subtrees (multi-GPU shard); the
// default is the whole tree.
static u64 coop_max_from_env() {
  static const u64 v = [] {
    const char* e = getenv("QPZK_COOP_MAX");
    u64 x = e ? strtoull(e, nullptr, 10) : 1024;
    return x > QPZK_CLIMB_MAX_START ? (u64)QPZK_CLIMB_MAX_START : x;
  }();
  return v;
}
static int build_tree(qpzk_ctx* c, const u64* src, u64 rs, u64 cs, u32 width, u32 log_n, u32 cap_height,
                      u64* levels, cudaEvent_t after_leaves, u64 leaf0 = 0, u64 nleaves = 0) {
  const u64 N = (u64)1 << log_n;
  if (nleaves == 0) nleaves = N - leaf0;
  const u32 top = log_n - cap_height;
  auto groups = [](u64 count) { return (unsigned)((count + QPZK_COOP_GROUPS - 1) / QPZK_COOP_GROUPS); };
  if (nleaves <= c->coop_max) {
    k_tree_climb<true><<<groups(nleaves), QPZK_COOP_THREADS, 0, c->stream>>>(src, rs, cs, width, levels, log_n, cap_height, 0,
                                                                            leaf0, nleaves, c->climb_counters);
    c->launches++;
    CU(cudaGetLastError());
    if (after_leaves) CU(cudaEventRecord(after_leaves, c->stream));
    return QPZK_OK;
  }
  // Leaf rows up to QPZK_COOP_LEAF_MAX (default 4096) take the 16-lane sponge as well - the climb kernel told to stop at
  // the leaves (cap_height = log_n) - although the levels above them start climbing only at coop_max nodes: 17 dependent
  // permutations of a 135-column row cost 17 x 25 us on a thread each and 17 x 21 us on 16 lanes at 4096 rows (a
  // voting-sized proof: 2.84 -> 2.74 ms); at 8192 rows the lanes are slower than the threads.
  static const u64 coop_leaf_max = [] {
    const char* e = getenv("QPZK_COOP_LEAF_MAX");
    u64 x = e ? strtoull(e, nullptr, 10) : 4096;
    return x > QPZK_CLIMB_MAX_START ? (u64)QPZK_CLIMB_MAX_START : x;
  }();
  if (nleaves <= coop_leaf_max && width > 4)
    k_tree_climb<true><<<groups(nleaves), QPZK_COOP_THREADS, 0, c->stream>>>(src, rs, cs, width, levels, log_n, log_n, 0, leaf0,
                                                                            nleaves, c->climb_counters);
  else
    k_leaf_hash<<<(unsigned)((nleaves + 127) / 128), 128, 0, c->stream>>>(src + leaf0 * rs, rs, cs, width, nleaves,
                                                                         levels + leaf0 * 4);
  c->launches++;
  CU(cudaGetLastError());
  if (after_leaves) CU(cudaEventRecord(after_leaves, c->stream));
  const u64 twoN = 2 * N;
  for (u32 l = 0; l < top; l++) {
    const u64 nout = nleaves >> (l + 1), first = leaf0 >> (l + 1);
    if (nout <= c->coop_max) {  // the rest of the way in one launch
      k_tree_climb<false><<<groups(nout), QPZK_COOP_THREADS, 0, c->stream>>>(nullptr, 0, 0, 0, levels, log_n, cap_height, l, first,
                                                                           nout, c->climb_counters);
      c->launches++;
      break;
    }
    const u64* in = levels + (twoN - (twoN >> l) + 2 * first) * 4;
    u64* out = levels + (twoN - (twoN >> (l + 1)) + first) * 4;
    k_merkle_level<<<(unsigned)((nout + 127) / 128), 128, 0, c->stream>>>(in, out, nout);
    c->launches++;
  }
  CU(cudaGetLastError());
  return QPZK_OK;
}

static const u64* cap_ptr(const u64* levels, u32 log_n, u32 cap_height) {
  u64 twoN = (u64)2 << log_n;
  return levels + (twoN - (twoN >> (log_n - cap_height))) * 4;
}

// ---- integer multiply-add peak (the Poseidon roofline denominators) ----
// KIND 0: IMAD (32-bit mad.lo, x = x*a + b chains). KIND 1: IMAD.WIDE.U32 with a 64-bit accumulator, the
// instruction the field multiply and the MDS layer are made of: acc[i] += x[(i+j)&7] * c[j], with x
// refreshed from the accumulators every trip so that no product is loop-invariant (an earlier
// version multiplied two loop-invariant registers; ptxas hoisted the product and the "peak" it
// reported was an add rate). cuobjdump -sass shows 64 IMAD.WIDE.U32 per trip for KIND 1.
template <int KIND>
__global__ void __launch_bounds__(256) k_imad_peak(u64* out, int iters, u32 seed, u32 c0, u32 c1, u32 c2, u32 c3) {
  const u32 cc[8] = {c0, c1, c2, c3, c0 ^ 5u, c1 ^ 9u, c2 ^ 3u, c3 ^ 6u};
  u32 lo[8], hi[8], x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    lo[i] = seed + threadIdx.x * 8 + i;
    hi[i] = seed * 3 + blockIdx.x + i;
    x[i] = lo[i] ^ hi[i];
  }
  for (int it = 0; it < iters; it++) {
    if (KIND == 0) {
#pragma unroll
      for (int j = 0; j < 8; j++)
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(cc[j]), "r"(hi[i]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; j++)
#pragma unroll
        for (int i = 0; i < 8; i++)
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                       : "+r"(lo[i]), "+r"(hi[i])
                       : "r"(x[(i + j) & 7]), "r"(cc[j]));
#pragma unroll
      for (int i = 0; i < 8; i++) x[i] = lo[i] ^ hi[(i + 1) & 7];
    }
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += (((u64)hi[i] << 32) | lo[i]) + x[i];
  if (s == 0x123456789ULL) out[0] = s;  // keep the chains alive
}

// ------------------------------------------------------------------------------------------
extern "C" {

const char* qpzk_last_error(void) { return g_err.c_str(); }

// The library allocates from its own stream-ordered pool (one per device, shared by the contexts on it,
// never trimmed: a commit allocates and releases hundreds of MB per call) instead of reconfiguring the
// device's default pool under the host application.
static int device_pool(int device, cudaMemPool_t* out) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {nullptr};
  std::lock_guard<std::mutex> lk(mu);
  if (device >= 64) return fail(QPZK_ERR_UNSUPPORTED, "device index above 63");
  if (!pools[device]) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    CU(cudaMemPoolCreate(&pools[device], &props));
    uint64_t thresh = ~0ull;
    CU(cudaMemPoolSetAttribute(pools[device], cudaMemPoolAttrReleaseThreshold, &thresh));
  }
  *out = pools[device];
  return QPZK_OK;
}

// Poseidon tables -> __constant__ and kernel attributes, once per device (kernels of other contexts may be
// reading them).
static int device_init_once(int device) {
  static std::mutex mu;
  static bool device_ready[64] = {false};
  std::lock_guard<std::mutex> lk(mu);
  static PoseidonTablesHost* T = nullptr;
  if (!T) {
    T = new PoseidonTablesHost();
    build_poseidon_tables(T, PV_DENSE_PARTIAL);
  }
  if (device_ready[device]) return QPZK_OK;
  CU(poseidon_upload_tables(*T));
  // transforms of 2^12 points stage 48 KB + twiddles in shared memory: opt in above the 48 KB default
  const int kMaxSmem = 72 * 1024;
  CU(cudaFuncSetAttribute(k_ntt_small<true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallSmemBytes));
  CU(cudaFuncSetAttribute(k_ntt_small<false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallSmemBytes));
  CU(cudaFuncSetAttribute(k_ntt_small<true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallSmemBytes));
  CU(cudaFuncSetAttribute(k_ntt_small<false, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallSmemBytes));
  CU(cudaFuncSetAttribute(k_ntt_cluster<3, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallSmemBytes));
  CU(cudaFuncSetAttribute(k_ntt_cluster<4, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallSmemBytes));
  CU(cudaFuncSetAttribute(k_ntt_pass_a<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  CU(cudaFuncSetAttribute(k_ntt_pass_a<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  CU(cudaFuncSetAttribute(k_ntt_pass_b_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  CU(cudaFuncSetAttribute(k_ntt_pass_b_transpose, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  CU(cudaDeviceSynchronize());
  device_ready[device] = true;
  return QPZK_OK;
}

// ------------------------------------------------------------------------------------------
// Host -> device on the context's stream. Pinned or registered memory is one asynchronous copy. From PAGEABLE memory
// - what a Rust Vec<F> or a numpy array is - the driver stages through its own bounce buffer on the calling thread:
// 11-12 GB/s on the B200 boxes against 55 GB/s from pinned memory, and a 2^18-row witness with its salts is 484 MB
// (43 ms of a 121 ms proof). Above h2d_min bytes such a buffer therefore goes through pinned chunks of the
// context's own, filled by h2d_threads host threads while earlier chunks are on the wire: 28-31 GB/s
// (scripts/exp/h2d_bench.cu; below ~32 MB the driver's path is as fast).
// The caller's buffer has been read completely when this returns. QPZK_H2D_THREADS (0 or 1 = always the driver's
// path) and QPZK_H2D_MIN_MB override the defaults.
static const size_t kH2DChunk = (size_t)2 << 20;  // per thread, two of them
static const int kH2DMaxThreads = 16;
static void h2d_config_from_env(qpzk_ctx* c) {
  static const int threads = [] {
    const char* e = getenv("QPZK_H2D_THREADS");
    long v = e ? strtol(e, nullptr, 10) : 4;
    return (int)(v < 0 ? 0 : v > kH2DMaxThreads ? kH2DMaxThreads : v);
  }();
  static const size_t min_bytes = [] {
    const char* e = getenv("QPZK_H2D_MIN_MB");
    unsigned long long v = e ? strtoull(e, nullptr, 10) : 32;
    return (size_t)(v > 65536 ? 65536 : v) << 20;
  }();
  c->h2d_threads = threads;
  c->h2d_min = min_bytes;
}
// Every thread owns a contiguous part of the buffer and two staging chunks: fill one, hand it to the copy engine,
// fill the other. The threads share nothing but the stream (the copies of different threads may interleave in
// any order; whatever is enqueued after this call comes after all of them), so there is no handshake between
// them - an earlier version that had all threads fill ONE chunk together and meet per chunk showed
// 200-500 ms stalls when a spinning thread lost its core.
static int h2d_staged(qpzk_ctx* c, char* dst, const char* src, size_t bytes) {
  const int T = c->h2d_threads;
  if (!c->h2d_arena) CU(cudaHostAlloc(&c->h2d_arena, (size_t)2 * T * kH2DChunk, cudaHostAllocDefault));
  for (int i = 0; i < 2 * T; i++)
    if (!c->h2d_ev[i]) CU(cudaEventCreateWithFlags(&c->h2d_ev[i], cudaEventDisableTiming));
  const size_t per = (((bytes + T - 1) / T) + 4095) & ~(size_t)4095;
  static const bool trace = getenv("QPZK_H2D_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  std::atomic<int> first_err{(int)cudaSuccess};
  auto work = [&](int t) {
    const size_t beg = std::min(bytes, (size_t)t * per), end = std::min(bytes, beg + per);
    if (t && cudaSetDevice(c->device) != cudaSuccess) {
      int ok = (int)cudaSuccess;
      first_err.compare_exchange_strong(ok, (int)cudaGetLastError());
      return;
    }
    int j = 0;
    for (size_t off = beg; off < end; off += kH2DChunk, j ^= 1) {
      const size_t len = std::min(kH2DChunk, end - off);
      char* buf = static_cast<char*>(c->h2d_arena) + (size_t)(2 * t + j) * kH2DChunk;
      cudaEvent_t ev = c->h2d_ev[2 * t + j];
      cudaError_t e = cudaEventSynchronize(ev);  // the copy that last read this chunk has left it
      if (e == cudaSuccess) {
        memcpy(buf, src + off, len);
        e = cudaMemcpyAsync(dst + off, buf, len, cudaMemcpyHostToDevice, c->stream);
      }
      if (e == cudaSuccess) e = cudaEventRecord(ev, c->stream);
      if (e != cudaSuccess) {
        int ok = (int)cudaSuccess;
        first_err.compare_exchange_strong(ok, (int)e);
        return;
      }
    }
  };
  {
    struct Joiner {
      std::vector<std::thread> th;
      ~Joiner() {
        for (auto& t : th)
          if (t.joinable()) t.join();
      }
    } workers;
    workers.th.reserve(T - 1);
    for (int t = 1; t < T; t++) workers.th.emplace_back(work, t);
    work(0);
  }
  if (trace)
    fprintf(stderr, "[qpzk] staged upload: %zu bytes, %d threads, %.2f ms\n", bytes, T,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
  CU((cudaError_t)first_err.load());
  return QPZK_OK;
}
static int h2d_copy(qpzk_ctx* c, void* dst, const void* src, size_t bytes) {
  if (c->h2d_threads > 1 && bytes >= c->h2d_min) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, src);
    if (e != cudaSuccess) cudaGetLastError();  // older drivers report plain host memory as an error: clear it
    if (e != cudaSuccess || a.type == cudaMemoryTypeUnregistered)
      return h2d_staged(c, static_cast<char*>(dst), static_cast<const char*>(src), bytes);
  }
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return QPZK_OK;
}

static void ctx_release(qpzk_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) ctx_wait(c);
  for (auto& kv : c->root_tabs) {
    dev_free(c, kv.second.first);
    dev_free(c, kv.second.second);
  }
  for (auto& kv : c->coset_pm) dev_free(c, kv.second);
  for (auto& kv : c->tw_mats) dev_free(c, kv.second);
  for (auto& kv : c->pow_tabs) {
    dev_free(c, kv.second.first);
    dev_free(c, kv.second.second);
  }
  dev_free(c, c->scratch_path);
  dev_free(c, c->climb_counters);
  if (c->stream) ctx_wait(c);
  for (auto& e : c->ev)
    if (e) cudaEventDestroy(e);
  if (c->sync_ev) cudaEventDestroy(c->sync_ev);
  for (auto& e : c->h2d_ev)
    if (e) cudaEventDestroy(e);
  if (c->h2d_arena) cudaFreeHost(c->h2d_arena);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int qpzk_ctx_create(int device, uint32_t flags, qpzk_ctx** out) {
  return guarded([&]() -> int {
    if (!out) return fail(QPZK_ERR_BAD_ARG, "out is NULL");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev || device >= 64)
      return fail(QPZK_ERR_BAD_ARG, "no such CUDA device (there is no CPU fallback)");
    CU(cudaSetDevice(device));
    std::unique_ptr<qpzk_ctx, void (*)(qpzk_ctx*)> c(new qpzk_ctx(), ctx_release);
    c->device = device;
    for (auto& e : c->ev) e = nullptr;
    c->coop_max = coop_max_from_env();
    h2d_config_from_env(c.get());
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    QP(device_pool(device, &c->pool));
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->blocking_sync = (flags & QPZK_CTX_BLOCKING_SYNC) != 0;
    c->yield_sync = !c->blocking_sync && (flags & QPZK_CTX_YIELD_SYNC) != 0;
    CU(cudaEventCreateWithFlags(&c->sync_ev, cudaEventBlockingSync | cudaEventDisableTiming));
    for (auto& e : c->ev) CU(cudaEventCreate(&e));
    QP(device_init_once(device));
    QP(dev_alloc(c.get(), 64 * 4 * 8 + 4096 * 8, &c->scratch_path));
    u64* cnt = nullptr;
    QP(dev_alloc(c.get(), 2 * QPZK_CLIMB_MAX_START * sizeof(u32), &cnt));
    c->climb_counters = reinterpret_cast<u32*>(cnt);
    CU(cudaMemsetAsync(c->climb_counters, 0, 2 * QPZK_CLIMB_MAX_START * sizeof(u32), c->stream));
    CU(ctx_wait(c.get()));
    *out = c.release();
    return QPZK_OK;
  });
}

void qpzk_ctx_destroy(qpzk_ctx* c) { ctx_release(c); }

int qpzk_ctx_sync(qpzk_ctx* c) {
  if (!c) return fail(QPZK_ERR_BAD_ARG, "ctx is NULL");
  CU(ctx_wait(c));
  return QPZK_OK;
}
void* qpzk_ctx_stream(qpzk_ctx* c) { return c ? (void*)c->stream : nullptr; }
uint64_t qpzk_ctx_launch_count(const qpzk_ctx* c) { return c ? c->launches : 0; }
int qpzk_ctx_stage_ms(qpzk_ctx* c, float* out_ms) {
  if (!c || !out_ms) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  memcpy(out_ms, c->stage_ms, sizeof c->stage_ms);
  return QPZK_OK;
}

void* qpzk_poseidon_tables_host(void) {  // test hook: the generated tables (PoseidonTablesHost*)
  static PoseidonTablesHost* T = [] {
    PoseidonTablesHost* t = new (std::nothrow) PoseidonTablesHost();
    if (t) build_poseidon_tables(t);
    return t;
  }();
  return T;
}

int qpzk_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  CU(cudaHostAlloc(out, bytes ? bytes : 8, cudaHostAllocDefault));
  return QPZK_OK;
}
void qpzk_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
int qpzk_dev_alloc(qpzk_ctx* c, size_t bytes, void** out) {
  if (!c || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  CU(cudaSetDevice(c->device));
  u64* p;
  QP(dev_alloc(c, bytes, &p));
  *out = p;
  return QPZK_OK;
}
void qpzk_dev_free(qpzk_ctx* c, void* p) {
  if (c) dev_free(c, p);
}
int qpzk_memcpy_h2d(qpzk_ctx* c, void* dst, const void* src, size_t bytes) {
  if (!c || !dst || !src) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  CU(cudaSetDevice(c->device));
  QP(h2d_copy(c, dst, src, bytes));
  CU(ctx_wait(c));
  return QPZK_OK;
}
int qpzk_memcpy_d2h(qpzk_ctx* c, void* dst, const void* src, size_t bytes) {
  if (!c || !dst || !src) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(ctx_wait(c));
  return QPZK_OK;
}

// ---- hashing ----
int qpzk_poseidon_permute(qpzk_ctx* c, uint64_t* states, uint64_t n) {
  if (!c || (!states && n)) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  if (!n) return QPZK_OK;
  CU(cudaSetDevice(c->device));
  DevBuf d(c);
  QP(d.alloc(n * 96));
  CU(cudaMemcpyAsync(d.p, states, n * 96, cudaMemcpyHostToDevice, c->stream));
  k_permute<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(d.p, n);
  c->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(states, d.p, n * 96, cudaMemcpyDeviceToHost, c->stream));
  CU(ctx_wait(c));
  return QPZK_OK;
}

int qpzk_hash_no_pad(qpzk_ctx* c, const uint64_t* inputs, uint64_t n, uint32_t len, uint64_t* out) {
  if (!c || !inputs || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  if (!n) return QPZK_OK;
  if (len == 0) return fail(QPZK_ERR_BAD_ARG, "len must be > 0");
  CU(cudaSetDevice(c->device));
  DevBuf din(c), dout(c);
  QP(din.alloc(n * len * 8));
  QP(dout.alloc(n * 32));
  CU(cudaMemcpyAsync(din.p, inputs, n * len * 8, cudaMemcpyHostToDevice, c->stream));
  k_hash_no_pad<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(din.p, n, len, dout.p);
  c->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, dout.p, n * 32, cudaMemcpyDeviceToHost, c->stream));
  CU(ctx_wait(c));
  return QPZK_OK;
}

int qpzk_two_to_one(qpzk_ctx* c, const uint64_t* pairs, uint64_t n, uint64_t* out) {
  if (!c || !pairs || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  if (!n) return QPZK_OK;
  CU(cudaSetDevice(c->device));
  DevBuf din(c), dout(c);
  QP(din.alloc(n * 64));
  QP(dout.alloc(n * 32));
  CU(cudaMemcpyAsync(din.p, pairs, n * 64, cudaMemcpyHostToDevice, c->stream));
  k_merkle_level<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(din.p, dout.p, n);
  c->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, dout.p, n * 32, cudaMemcpyDeviceToHost, c->stream));
  CU(ctx_wait(c));
  return QPZK_OK;
}

// ---- MerkleTree ----
static bool is_pow2(uint64_t x) { return x && !(x & (x - 1)); }
static uint32_t ilog2(uint64_t x) {
  uint32_t k = 0;
  while (((uint64_t)1 << k) < x) k++;
  return k;
}

int qpzk_merkle_new(qpzk_ctx* c, const uint64_t* leaves, uint64_t nleaves, uint32_t leaf_len,
                    uint32_t cap_height, qpzk_tree** out) {
  return guarded([&]() -> int {
    if (!c || !leaves || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
    if (!is_pow2(nleaves) || nleaves > ((uint64_t)1 << 30)) return fail(QPZK_ERR_BAD_ARG, "nleaves must be a power of two, at most 2^30");
    uint32_t log_n = ilog2(nleaves);
    if (cap_height > log_n)
      return fail(QPZK_ERR_BAD_ARG, "cap_height must be at most log2(nleaves)");  // plonky2 asserts the same
    CU(cudaSetDevice(c->device));
    std::unique_ptr<qpzk_tree> t(new qpzk_tree());
    t->ctx = c;
    t->log_n = log_n;
    t->cap_height = cap_height;
    t->leaf_len = leaf_len;
    DevBuf dl(c);
    QP(dl.alloc(nleaves * (leaf_len ? leaf_len : 1) * 8));
    QP(dev_alloc(c, nleaves * 2 * 32, &t->levels));
    CU(cudaMemcpyAsync(dl.p, leaves, nleaves * leaf_len * 8, cudaMemcpyHostToDevice, c->stream));
    QP(build_tree(c, dl.p, leaf_len, 1, leaf_len, log_n, cap_height, t->levels, nullptr));
    CU(ctx_wait(c));
    *out = t.release();
    return QPZK_OK;
  });
}

int qpzk_tree_cap(const qpzk_tree* t, uint64_t* out) {
  if (!t || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  CU(cudaSetDevice(t->ctx->device));
  CU(cudaMemcpyAsync(out, cap_ptr(t->levels, t->log_n, t->cap_height), ((size_t)32) << t->cap_height,
                     cudaMemcpyDeviceToHost, t->ctx->stream));
  CU(ctx_wait(t->ctx));
  return QPZK_OK;
}

static int prove_from_levels(qpzk_ctx* c, const u64* levels, u32 log_n, u32 cap_height, u64 leaf,
                             uint64_t* siblings) {
  u32 L = log_n - cap_height;
  if (L == 0) return QPZK_OK;
  if (L > 64) return fail(QPZK_ERR_BAD_ARG, "tree too deep");
  k_gather_path<<<1, 256, 0, c->stream>>>(levels, log_n, cap_height, leaf, c->scratch_path);
  c->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(siblings, c->scratch_path, (size_t)L * 32, cudaMemcpyDeviceToHost, c->stream));
  CU(ctx_wait(c));
  return QPZK_OK;
}

int qpzk_tree_prove(const qpzk_tree* t, uint64_t leaf_index, uint64_t* siblings) {
  if (!t || !siblings) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  if (leaf_index >> t->log_n) return fail(QPZK_ERR_BAD_ARG, "leaf_index out of range");
  CU(cudaSetDevice(t->ctx->device));
  return prove_from_levels(t->ctx, t->levels, t->log_n, t->cap_height, leaf_index, siblings);
}

static int export_digests(qpzk_ctx* c, const u64* levels, u32 log_n, u32 cap_height, uint64_t* out) {
  u64 total = ((u64)2 << log_n) - ((u64)2 << cap_height);
  if (!total) return QPZK_OK;
  DevBuf d(c);
  QP(d.alloc(total * 32));
  k_export_digests<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(levels, log_n, cap_height, d.p);
  c->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, d.p, total * 32, cudaMemcpyDeviceToHost, c->stream));
  CU(ctx_wait(c));
  return QPZK_OK;
}

int qpzk_tree_digests(const qpzk_tree* t, uint64_t* out) {
  if (!t || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  CU(cudaSetDevice(t->ctx->device));
  return export_digests(t->ctx, t->levels, t->log_n, t->cap_height, out);
}

void qpzk_tree_free(qpzk_tree* t) {
  if (!t) return;
  cudaSetDevice(t->ctx->device);
  delete t;
}

}  // extern "C"

// ---- PolynomialBatch ----
// sync = false: everything is enqueued on the context's stream and the call returns without waiting (the
// prover pipeline); per-stage times are then not collected.
static int commit_impl(qpzk_ctx* c, const uint64_t* in, bool in_is_host, bool is_coeffs, uint32_t ncols,
                       uint32_t k, uint32_t r, uint32_t cap_height, const uint64_t* salts, bool salts_host,
                       uint32_t salt_cols, qpzk_batch** out, uint32_t sub_begin = 0, uint32_t sub_end = 0,
                       bool sync = true) {
  if (!c || !in || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  if (ncols == 0) return fail(QPZK_ERR_BAD_ARG, "ncols must be > 0");
  if (!salts) salt_cols = 0;
  if (k > 30 || r > 30 || k + r > 30) return fail(QPZK_ERR_UNSUPPORTED, "LDE domain too large");
  if (cap_height > k + r) return fail(QPZK_ERR_BAD_ARG, "cap_height must be at most degree_bits + rate_bits");
  CU(cudaSetDevice(c->device));
  const u64 n = (u64)1 << k, N = n << r;
  const u32 width = ncols + salt_cols;
  // shard = a range of cap subtrees = a range of leaves; it must consist of whole leaf blocks (cosets)
  if (sub_end == 0) sub_end = 1u << cap_height;
  if (sub_begin >= sub_end || sub_end > (1u << cap_height)) return fail(QPZK_ERR_BAD_ARG, "bad subtree range");
  const bool sharded = sub_begin != 0 || sub_end != (1u << cap_height);
  const u64 leaf0 = ((u64)sub_begin << (k + r)) >> cap_height, leaf1 = ((u64)sub_end << (k + r)) >> cap_height;
  if (sharded && ((leaf0 & (n - 1)) || (leaf1 & (n - 1))))
    return fail(QPZK_ERR_UNSUPPORTED, "shard must cover whole cosets: subtree range must be a multiple of 2^(cap_height - rate_bits)");
  const u32 blk0 = (u32)(leaf0 >> k), nblk = (u32)((leaf1 - leaf0) >> k);
  std::unique_ptr<qpzk_batch> b(new qpzk_batch());
  b->ctx = c;
  b->ncols = ncols;
  b->salt_cols = salt_cols;
  b->degree_bits = k;
  b->rate_bits = r;
  b->cap_height = cap_height;
  b->leaf0 = leaf0;
  b->leaf1 = leaf1;
  DevBuf staging(c), salt_staging(c);
  QP(dev_alloc(c, (size_t)ncols * n * 8, &b->coeffs));
  const u64 nown = leaf1 - leaf0;
  QP(dev_alloc(c, (size_t)width * nown * 8, &b->lde_alloc));
  b->lde = b->lde_alloc - leaf0;
  b->lde_stride = nown;
  QP(dev_alloc(c, (size_t)N * 2 * 32, &b->levels));

  cudaEvent_t* ev = c->ev;
  if (sync) CU(cudaEventRecord(ev[0], c->stream));
  const u64* src = in;
  if (in_is_host) {
    u64* dstp = b->coeffs;
    if (!is_coeffs) {
      QP(staging.alloc((size_t)ncols * n * 8));
      dstp = staging.p;
    }
    QP(h2d_copy(c, dstp, in, (size_t)ncols * n * 8));
    src = dstp;
  } else if (is_coeffs) {
    CU(cudaMemcpyAsync(b->coeffs, in, (size_t)ncols * n * 8, cudaMemcpyDeviceToDevice, c->stream));
    src = b->coeffs;
  }
  const u64* salt_src = salts;
  if (salt_cols && salts_host) {
    QP(salt_staging.alloc((size_t)salt_cols * N * 8));
    QP(h2d_copy(c, salt_staging.p, salts, (size_t)salt_cols * N * 8));
    salt_src = salt_staging.p;
  }
  if (sync) CU(cudaEventRecord(ev[1], c->stream));
  if (!is_coeffs) QP(launch_ifft(c, src, n, b->coeffs, n, ncols, (int)k));
  if (sync) CU(cudaEventRecord(ev[2], c->stream));
  // foreign digests and cap entries of a shard read as zero (the caller all-gathers the subtree roots)
  if (sharded) CU(cudaMemsetAsync(b->levels, 0, (size_t)N * 2 * 32, c->stream));
  QP(launch_lde(c, b->coeffs, n, b->lde, nown, ncols, (int)k, (int)r, blk0, nblk));
  if (salt_cols) {
    k_bitrev_rows<<<dim3((unsigned)((nown + 255) / 256), salt_cols), 256, 0, c->stream>>>(
        salt_src, b->lde + (size_t)ncols * nown, nown, leaf0, nown, (int)(k + r), salt_cols);
    c->launches++;
  }
  if (sync) CU(cudaEventRecord(ev[3], c->stream));
  QP(build_tree(c, b->lde, 1, nown, width, k + r, cap_height, b->levels, sync ? ev[4] : nullptr, leaf0, nown));
  if (sync) {
    CU(cudaEventRecord(ev[5], c->stream));
    CU(ctx_wait(c));
    for (int i = 0; i < 5; i++) cudaEventElapsedTime(&c->stage_ms[i], ev[i], ev[i + 1]);
    c->stage_ms[QPZK_STAGE_D2H] = 0;
  }
  *out = b.release();
  return QPZK_OK;
}

extern "C" {

int qpzk_batch_from_values(qpzk_ctx* c, const uint64_t* values, uint32_t ncols, uint32_t degree_bits,
                           uint32_t rate_bits, uint32_t cap_height, const uint64_t* salts, uint32_t salt_cols,
                           qpzk_batch** out) {
  return guarded([&] { return commit_impl(c, values, true, false, ncols, degree_bits, rate_bits, cap_height, salts, true, salt_cols, out); });
}
int qpzk_batch_from_coeffs(qpzk_ctx* c, const uint64_t* coeffs, uint32_t ncols, uint32_t degree_bits,
                           uint32_t rate_bits, uint32_t cap_height, const uint64_t* salts, uint32_t salt_cols,
                           qpzk_batch** out) {
  return guarded([&] { return commit_impl(c, coeffs, true, true, ncols, degree_bits, rate_bits, cap_height, salts, true, salt_cols, out); });
}
int qpzk_batch_from_values_dev(qpzk_ctx* c, const uint64_t* values, uint32_t ncols, uint32_t degree_bits,
                               uint32_t rate_bits, uint32_t cap_height, const uint64_t* salts,
                               uint32_t salt_cols, qpzk_batch** out) {
  return guarded([&] { return commit_impl(c, values, false, false, ncols, degree_bits, rate_bits, cap_height, salts, false, salt_cols, out); });
}
int qpzk_batch_from_coeffs_dev(qpzk_ctx* c, const uint64_t* coeffs, uint32_t ncols, uint32_t degree_bits,
                               uint32_t rate_bits, uint32_t cap_height, const uint64_t* salts,
                               uint32_t salt_cols, qpzk_batch** out) {
  return guarded([&] { return commit_impl(c, coeffs, false, true, ncols, degree_bits, rate_bits, cap_height, salts, false, salt_cols, out); });
}
int qpzk_batch_from_values_shard_dev(qpzk_ctx* c, const uint64_t* values, uint32_t ncols, uint32_t degree_bits,
                                     uint32_t rate_bits, uint32_t cap_height, const uint64_t* salts,
                                     uint32_t salt_cols, uint32_t subtree_begin, uint32_t subtree_end,
                                     qpzk_batch** out) {
  if (subtree_end == 0) return fail(QPZK_ERR_BAD_ARG, "empty subtree range");
  return guarded([&] {
    return commit_impl(c, values, false, false, ncols, degree_bits, rate_bits, cap_height, salts, false, salt_cols, out,
                       subtree_begin, subtree_end);
  });
}
int qpzk_batch_from_values_shard_dev_async(qpzk_ctx* c, const uint64_t* values, uint32_t ncols, uint32_t degree_bits,
                                           uint32_t rate_bits, uint32_t cap_height, const uint64_t* salts,
                                           uint32_t salt_cols, uint32_t subtree_begin, uint32_t subtree_end,
                                           qpzk_batch** out) {
  if (subtree_end == 0) return fail(QPZK_ERR_BAD_ARG, "empty subtree range");
  return guarded([&] {
    return commit_impl(c, values, false, false, ncols, degree_bits, rate_bits, cap_height, salts, false, salt_cols, out,
                       subtree_begin, subtree_end, false);
  });
}
int qpzk_batch_from_coeffs_shard_dev(qpzk_ctx* c, const uint64_t* coeffs, uint32_t ncols, uint32_t degree_bits,
                                     uint32_t rate_bits, uint32_t cap_height, const uint64_t* salts,
                                     uint32_t salt_cols, uint32_t subtree_begin, uint32_t subtree_end,
                                     qpzk_batch** out) {
  if (subtree_end == 0) return fail(QPZK_ERR_BAD_ARG, "empty subtree range");
  return guarded([&] {
    return commit_impl(c, coeffs, false, true, ncols, degree_bits, rate_bits, cap_height, salts, false, salt_cols, out,
                       subtree_begin, subtree_end);
  });
}
int qpzk_batch_set_cap(qpzk_batch* b, const uint64_t* cap) {
  if (!b || !cap) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  qpzk_ctx* c = b->ctx;
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(const_cast<u64*>(cap_ptr(b->levels, b->log_N(), b->cap_height)), cap, ((size_t)32) << b->cap_height,
                     cudaMemcpyHostToDevice, c->stream));
  CU(ctx_wait(c));
  return QPZK_OK;
}

int qpzk_batch_cap(const qpzk_batch* b, uint64_t* out) {
  if (!b || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  qpzk_ctx* c = b->ctx;
  CU(cudaSetDevice(c->device));
  CU(cudaEventRecord(c->ev[0], c->stream));
  CU(cudaMemcpyAsync(out, cap_ptr(b->levels, b->log_N(), b->cap_height), ((size_t)32) << b->cap_height,
                     cudaMemcpyDeviceToHost, c->stream));
  CU(cudaEventRecord(c->ev[1], c->stream));
  CU(ctx_wait(c));
  cudaEventElapsedTime(&c->stage_ms[QPZK_STAGE_D2H], c->ev[0], c->ev[1]);
  return QPZK_OK;
}
uint64_t* qpzk_batch_cap_dev(qpzk_batch* b) {
  return b ? const_cast<u64*>(cap_ptr(b->levels, b->log_N(), b->cap_height)) : nullptr;
}
int qpzk_batch_coeffs(const qpzk_batch* b, uint64_t* out) {
  if (!b || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  CU(cudaSetDevice(b->ctx->device));
  CU(cudaMemcpyAsync(out, b->coeffs, ((size_t)b->ncols << b->degree_bits) * 8, cudaMemcpyDeviceToHost,
                     b->ctx->stream));
  CU(ctx_wait(b->ctx));
  return QPZK_OK;
}
int qpzk_batch_get_lde_rows(const qpzk_batch* b, const uint32_t* idx, uint32_t nidx, uint32_t step,
                            uint64_t* out) {
  if (!b || !idx || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  if (!nidx) return QPZK_OK;
  qpzk_ctx* c = b->ctx;
  CU(cudaSetDevice(c->device));
  const u32 lb = b->log_N();
  const u64 N = (u64)1 << lb;
  for (uint32_t i = 0; i < nidx; i++) {
    const u64 nat = (u64)idx[i] * step;
    if (nat >= N) return fail(QPZK_ERR_BAD_ARG, "index*step out of range");
    if (lb && !b->owns(glh_bitrev(nat, lb))) return fail(QPZK_ERR_BAD_ARG, "row belongs to another rank's shard");
  }
  DevBuf didx(c), dout(c);
  QP(didx.alloc((size_t)nidx * 4));
  QP(dout.alloc((size_t)nidx * b->ncols * 8));
  CU(cudaMemcpyAsync(didx.p, idx, (size_t)nidx * 4, cudaMemcpyHostToDevice, c->stream));
  k_gather_rows<<<nidx, 128, 0, c->stream>>>(b->lde, b->lde_stride, (int)lb, (const u32*)didx.p, nidx, step, b->ncols, dout.p);
  c->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, dout.p, (size_t)nidx * b->ncols * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(ctx_wait(c));
  return QPZK_OK;
}
int qpzk_batch_open(const qpzk_batch* b, uint64_t leaf_index, uint64_t* leaf_out, uint64_t* siblings_out) {
  if (!b) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  if (leaf_index >> b->log_N()) return fail(QPZK_ERR_BAD_ARG, "leaf_index out of range");
  if (!b->owns(leaf_index)) return fail(QPZK_ERR_BAD_ARG, "leaf belongs to another rank's shard");
  qpzk_ctx* c = b->ctx;
  CU(cudaSetDevice(c->device));
  if (leaf_out) {
    if (b->width() > 4096) return fail(QPZK_ERR_UNSUPPORTED, "row too wide");
    u64* row = c->scratch_path + 64 * 4;
    k_gather_leaf<<<1, 128, 0, c->stream>>>(b->lde, b->lde_stride, leaf_index, b->width(), row);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(leaf_out, row, (size_t)b->width() * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(ctx_wait(c));
  }
  if (siblings_out) return prove_from_levels(c, b->levels, b->log_N(), b->cap_height, leaf_index, siblings_out);
  return QPZK_OK;
}
int qpzk_batch_export(const qpzk_batch* b, uint64_t* leaves, uint64_t* digests) {
  if (!b) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  if (b->sharded()) return fail(QPZK_ERR_UNSUPPORTED, "a shard holds only its own leaves: export the unsharded batch");
  qpzk_ctx* c = b->ctx;
  CU(cudaSetDevice(c->device));
  u64 N = (u64)1 << b->log_N();
  if (leaves) {
    DevBuf rows(c);
    QP(rows.alloc((size_t)N * b->width() * 8));
    dim3 grid((unsigned)((N + 31) / 32), (b->width() + 31) / 32);
    k_transpose_to_rows<<<grid, dim3(32, 8), 0, c->stream>>>(b->lde, rows.p, N, b->width());
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(leaves, rows.p, (size_t)N * b->width() * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(ctx_wait(c));
  }
  if (digests) return export_digests(c, b->levels, b->log_N(), b->cap_height, digests);
  return QPZK_OK;
}
// ---- PolynomialBatch <-> bytes ------------------------------------------------------------------------
// The layout of plonky2's `Write::write_polynomial_batch` / `Read::read_polynomial_batch`
// (qp-plonky2 util/serialization, un-vendored - restated, no fixture ships one: SURVEY 8(f).4): the
// `constants_sigmas_commitment` field of the serialized ProverOnlyCircuitData that
// /root/reference/wormhole/prover/src/lib.rs:105-187 (`new_from_files`) reads back, so that a prover started
// from files never recomputes the commit. All integers little-endian u64 unless noted:
//   polynomials.len(), then per polynomial: coeffs.len(), coeffs
//   merkle_tree: leaves.len(), then per leaf: len, elements | digests.len(), digests (4 words each, plonky2's
//                interleaved layout) | cap.height(), 2^height cap hashes
//   degree_log, rate_bits, blinding (1 byte)
static const u32 kSaltSize = 4;  // SALT_SIZE of a blinded oracle
static u64 batch_bytes_len(u32 ncols, u32 width, u32 k, u32 r, u32 h) {
  const u64 n = (u64)1 << k, N = n << r;
  return 8 + (u64)ncols * (8 + 8 * n) + 8 + N * (8 + 8 * (u64)width) + 8 + 32 * (2 * N - ((u64)2 << h)) + 8 + ((u64)32 << h) + 17;
}
// serialized leaves ([N] records of 1 + width words, the first one the length) -> column-major LDE
__global__ void k_rows_to_columns(const u64* __restrict__ rec, u64* __restrict__ lde, u64 N, u32 width) {
  __shared__ u64 tile[32][33];
  const u64 r0 = (u64)blockIdx.x * 32;
  const u32 c0 = blockIdx.y * 32;
  for (u32 rr = threadIdx.y; rr < 32; rr += blockDim.y) {
    const u32 cc = c0 + threadIdx.x;
    if (cc < width && r0 + rr < N) tile[rr][threadIdx.x] = rec[(r0 + rr) * (width + 1) + 1 + cc];
  }
  __syncthreads();
  for (u32 cc = threadIdx.y; cc < 32; cc += blockDim.y) {
    const u64 row = r0 + threadIdx.x;
    if (c0 + cc < width && row < N) lde[(u64)(c0 + cc) * N + row] = tile[threadIdx.x][cc];
  }
}
// plonky2 `digests` -> level-major (the inverse of k_export_digests: same index map, copy the other way)
__global__ void k_import_digests(const u64* __restrict__ in, u32 log_n, u32 cap_height, u64* __restrict__ levels) {
  const u32 L = log_n - cap_height;
  const u64 total = ((u64)2 << log_n) - ((u64)2 << cap_height);
  const u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const u64 twoN = (u64)2 << log_n;
  u32 i = 0;
  while (g >= twoN - (twoN >> (i + 1))) i++;
  const u64 j = g - (twoN - (twoN >> i));
  const u32 per_sub_bits = L - i;
  const u64 sub = j >> per_sub_bits, jj = j & (((u64)1 << per_sub_bits) - 1);
  const u64 idx = 2 * (((jj >> 1) << (i + 1)) + ((u64)1 << i) - 1) + (jj & 1);
  const u64 sub_len = ((u64)2 << L) - 2;
  const ulonglong2* s = reinterpret_cast<const ulonglong2*>(in + (sub * sub_len + idx) * 4);
  ulonglong2* d = reinterpret_cast<ulonglong2*>(levels + g * 4);
  d[0] = s[0];
  d[1] = s[1];
}
// counts words that differ or are not canonical
__global__ void k_count_mismatch(const u64* __restrict__ a, const u64* __restrict__ b, u64 n, unsigned long long* out) {
  u64 bad = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    bad += (a[i] != b[i]) || a[i] >= GL_P;
  if (bad) atomicAdd(out, (unsigned long long)bad);
}

int qpzk_batch_serialized_size(const qpzk_batch* b, uint64_t* nbytes) {
  if (!b || !nbytes) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  if (b->sharded()) return fail(QPZK_ERR_UNSUPPORTED, "a shard holds only its own leaves: serialise the unsharded batch");
  *nbytes = batch_bytes_len(b->ncols, b->width(), b->degree_bits, b->rate_bits, b->cap_height);
  return QPZK_OK;
}

int qpzk_batch_to_bytes(const qpzk_batch* b, uint8_t* out, uint64_t capacity) {
  return guarded([&]() -> int {
    uint64_t need = 0;
    QP(qpzk_batch_serialized_size(b, &need));
    if (!out || capacity < need) return fail(QPZK_ERR_BAD_ARG, "buffer too small for the serialised batch");
    qpzk_ctx* c = b->ctx;
    CU(cudaSetDevice(c->device));
    const u64 n = (u64)1 << b->degree_bits, N = (u64)1 << b->log_N();
    const u32 w = b->width();
    const u64 ndig = 2 * N - ((u64)2 << b->cap_height);
    uint8_t* p = out;
    auto put = [&](u64 v) { memcpy(p, &v, 8); p += 8; };
    put(b->ncols);
    for (u32 i = 0; i < b->ncols; i++) {
      put(n);
      CU(cudaMemcpyAsync(p, b->coeffs + (u64)i * n, n * 8, cudaMemcpyDeviceToHost, c->stream));
      p += n * 8;
    }
    put(N);
    {  // rows, each behind its length
      std::vector<u64> rows((size_t)N * w);
      QP(qpzk_batch_export(b, rows.data(), nullptr));
      for (u64 i = 0; i < N; i++) {
        put(w);
        memcpy(p, rows.data() + i * w, (size_t)w * 8);
        p += (size_t)w * 8;
      }
    }
    put(ndig);
    QP(export_digests(c, b->levels, b->log_N(), b->cap_height, reinterpret_cast<uint64_t*>(p)));  // waits for the stream
    p += ndig * 32;
    put(b->cap_height);
    CU(cudaMemcpyAsync(p, cap_ptr(b->levels, b->log_N(), b->cap_height), (size_t)32 << b->cap_height, cudaMemcpyDeviceToHost,
                       c->stream));
    p += (size_t)32 << b->cap_height;
    put(b->degree_bits);
    put(b->rate_bits);
    *p++ = b->salt_cols ? 1 : 0;
    CU(ctx_wait(c));
    if ((uint64_t)(p - out) != need) return fail(QPZK_ERR_CUDA, "internal error: serialised length");
    return QPZK_OK;
  });
}

// Nothing is recomputed: coefficients, leaves and digests go to the device as they are (what plonky2's
// `read_polynomial_batch` does on the host). The bytes are untrusted as far as SHAPES go - every length is
// checked against the others before anything is allocated - but, like plonky2, the digests are taken on
// trust unless QPZK_IMPORT_VERIFY asks for the check: LDE of the coefficients == the leaves' polynomial
// columns, Merkle tree of the leaves == the digests and the cap, every word canonical.
int qpzk_batch_from_bytes(qpzk_ctx* c, const uint8_t* bytes, uint64_t nbytes, uint32_t flags, qpzk_batch** out,
                          uint64_t* consumed) {
  return guarded([&]() -> int {
    if (!c || !bytes || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    const uint8_t* p = bytes;
    const uint8_t* const end = bytes + nbytes;
    bool short_read = false;
    auto get = [&]() -> u64 {
      if ((uint64_t)(end - p) < 8) { short_read = true; return 0; }
      u64 v;
      memcpy(&v, p, 8);
      p += 8;
      return v;
    };
    const char* trunc = "serialised PolynomialBatch: truncated";
    const u64 ncols64 = get();
    if (short_read) return fail(QPZK_ERR_BAD_ARG, trunc);
    if (ncols64 == 0 || ncols64 > 4096) return fail(QPZK_ERR_BAD_ARG, "serialised PolynomialBatch: bad polynomial count");
    const u32 ncols = (u32)ncols64;
    const u64 n = get();
    if (short_read) return fail(QPZK_ERR_BAD_ARG, trunc);
    if (!is_pow2(n) || n > ((u64)1 << 30)) return fail(QPZK_ERR_BAD_ARG, "serialised PolynomialBatch: polynomial length must be a power of two");
    const u32 k = ilog2(n);
    if ((uint64_t)(end - bytes) < 8 + (u64)ncols * (8 + 8 * n)) return fail(QPZK_ERR_BAD_ARG, trunc);
    const uint8_t* const coeffs0 = p;  // the first polynomial's coefficients; the others follow, each behind its length
    p += 8 * n;
    for (u32 i = 1; i < ncols; i++) {
      if (get() != n) return fail(QPZK_ERR_BAD_ARG, "serialised PolynomialBatch: polynomials differ in length");
      p += 8 * n;
    }
    const u64 N = get();
    if (short_read) return fail(QPZK_ERR_BAD_ARG, trunc);
    if (!is_pow2(N) || N < n || N > ((u64)1 << 30)) return fail(QPZK_ERR_BAD_ARG, "serialised PolynomialBatch: bad leaf count");
    const u32 r = ilog2(N) - k;
    const u64 w64 = get();
    if (short_read) return fail(QPZK_ERR_BAD_ARG, trunc);
    if (w64 != ncols && w64 != ncols + kSaltSize) return fail(QPZK_ERR_BAD_ARG, "serialised PolynomialBatch: leaf width must be the polynomial count (+ 4 salts)");
    const u32 w = (u32)w64;
    p -= 8;
    const uint8_t* const leaves0 = p;  // N records of (w + 1) words
    if ((uint64_t)(end - p) < N * 8 * (w + 1)) return fail(QPZK_ERR_BAD_ARG, trunc);
    for (u64 i = 0; i < N; i++) {
      u64 len;
      memcpy(&len, p + i * 8 * (w + 1), 8);
      if (len != w) return fail(QPZK_ERR_BAD_ARG, "serialised PolynomialBatch: leaves differ in length");
    }
    p += N * 8 * (w + 1);
    const u64 ndig = get();
    if (short_read) return fail(QPZK_ERR_BAD_ARG, trunc);
    if (ndig > 2 * N || (uint64_t)(end - p) < ndig * 32) return fail(QPZK_ERR_BAD_ARG, trunc);
    const uint8_t* const dig0 = p;
    p += ndig * 32;
    const u64 h64 = get();
    if (short_read) return fail(QPZK_ERR_BAD_ARG, trunc);
    if (h64 > k + r || ndig != 2 * N - ((u64)2 << h64)) return fail(QPZK_ERR_BAD_ARG, "serialised PolynomialBatch: digest count does not match the cap height");
    const u32 h = (u32)h64;
    if ((uint64_t)(end - p) < ((u64)32 << h) + 17) return fail(QPZK_ERR_BAD_ARG, trunc);
    const uint8_t* const cap0 = p;
    p += (u64)32 << h;
    const u64 degree_log = get(), rate_bits = get();
    const uint8_t blinding = *p++;
    if (degree_log != k || rate_bits != r || blinding > 1 || (blinding ? w != ncols + kSaltSize : w != ncols))
      return fail(QPZK_ERR_BAD_ARG, "serialised PolynomialBatch: degree_log / rate_bits / blinding contradict the vectors");

    std::unique_ptr<qpzk_batch> b(new qpzk_batch());
    b->ctx = c;
    b->ncols = ncols;
    b->salt_cols = w - ncols;
    b->degree_bits = k;
    b->rate_bits = r;
    b->cap_height = h;
    b->leaf0 = 0;
    b->leaf1 = N;
    b->lde_stride = N;
    QP(dev_alloc(c, (size_t)ncols * n * 8, &b->coeffs));
    QP(dev_alloc(c, (size_t)w * N * 8, &b->lde_alloc));
    b->lde = b->lde_alloc;
    QP(dev_alloc(c, (size_t)N * 2 * 32, &b->levels));
    // one strided copy drops the length word in front of every polynomial
    CU(cudaMemcpy2DAsync(b->coeffs, n * 8, coeffs0, (n + 1) * 8, n * 8, ncols, cudaMemcpyHostToDevice, c->stream));
    {
      DevBuf rec(c), dig(c);
      QP(rec.alloc((size_t)N * (w + 1) * 8));
      QP(h2d_copy(c, rec.p, leaves0, (size_t)N * (w + 1) * 8));
      k_rows_to_columns<<<dim3((unsigned)((N + 31) / 32), (w + 31) / 32), dim3(32, 8), 0, c->stream>>>(rec.p, b->lde, N, w);
      c->launches++;
      if (ndig) {
        QP(dig.alloc((size_t)ndig * 32));
        CU(cudaMemcpyAsync(dig.p, dig0, (size_t)ndig * 32, cudaMemcpyHostToDevice, c->stream));
        k_import_digests<<<(unsigned)((ndig + 255) / 256), 256, 0, c->stream>>>(dig.p, k + r, h, b->levels);
        c->launches++;
      }
      CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(const_cast<u64*>(cap_ptr(b->levels, k + r, h)), cap0, (size_t)32 << h, cudaMemcpyHostToDevice, c->stream));
    if (flags & QPZK_IMPORT_VERIFY) {
      DevBuf lde2(c), lev2(c), cnt(c);
      QP(lde2.alloc((size_t)ncols * N * 8));
      QP(lev2.alloc((size_t)N * 2 * 32));
      QP(cnt.alloc(8));
      CU(cudaMemsetAsync(cnt.p, 0, 8, c->stream));
      QP(launch_lde(c, b->coeffs, n, lde2.p, N, ncols, (int)k, (int)r));
      QP(build_tree(c, b->lde, 1, N, w, k + r, h, lev2.p, nullptr));
      k_count_mismatch<<<c->sm_count * 4, 256, 0, c->stream>>>(b->lde, lde2.p, (u64)ncols * N, (unsigned long long*)cnt.p);
      k_count_mismatch<<<c->sm_count * 4, 256, 0, c->stream>>>(b->levels, lev2.p, (2 * N - ((u64)1 << h)) * 4, (unsigned long long*)cnt.p);
      k_count_mismatch<<<c->sm_count * 4, 256, 0, c->stream>>>(b->coeffs, b->coeffs, (u64)ncols * n, (unsigned long long*)cnt.p);
      c->launches += 3;
      CU(cudaGetLastError());
      unsigned long long bad = 0;
      CU(cudaMemcpyAsync(&bad, cnt.p, 8, cudaMemcpyDeviceToHost, c->stream));
      CU(ctx_wait(c));
      if (bad) return fail(QPZK_ERR_BAD_ARG, "serialised PolynomialBatch: leaves / digests / cap are not the commitment of the polynomials");
    }
    CU(ctx_wait(c));  // the host bytes are borrowed for the call only
    if (consumed) *consumed = (uint64_t)(p - bytes);
    *out = b.release();
    return QPZK_OK;
  });
}

uint32_t qpzk_batch_ncols(const qpzk_batch* b) { return b ? b->ncols : 0; }
uint32_t qpzk_batch_width(const qpzk_batch* b) { return b ? b->width() : 0; }
uint32_t qpzk_batch_degree_bits(const qpzk_batch* b) { return b ? b->degree_bits : 0; }
void qpzk_batch_free(qpzk_batch* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  delete b;
}

int qpzk_measure_imad_peak(qpzk_ctx* c, int kind, double* out_ops_per_s) {
  if (!c || !out_ops_per_s) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  CU(cudaSetDevice(c->device));
  DevBuf d(c);
  QP(d.alloc(64));
  const int iters = 1024, threads = 256;
  const int blocks = c->sm_count * 8;
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CU(cudaEventRecord(c->ev[0], c->stream));
    if (kind == 0)
      k_imad_peak<0><<<blocks, threads, 0, c->stream>>>(d.p, iters, 12345u + rep, 17u, 15u, 41u, 16u);
    else
      k_imad_peak<1><<<blocks, threads, 0, c->stream>>>(d.p, iters, 12345u + rep, 17u, 15u, 41u, 16u);
    c->launches++;
    CU(cudaEventRecord(c->ev[1], c->stream));
    CU(ctx_wait(c));
    float ms;
    CU(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
    if (rep > 0 && ms < best) best = ms;
  }
  double ops = (double)blocks * threads * iters * 64;
  *out_ops_per_s = ops / (best * 1e-3);
  return QPZK_OK;
}

}  // extern "C"

#include "prover_host.inl"
