// Experiment harness (not product code): latency of ONE dependent chain of Poseidon permutations on the
// 16-lane cooperative path (the transcript / tree-top regime), against the naive host permutation.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../../qp-zk-circuits-rm_b200/csrc/poseidon.cuh"
#include "../../qp-zk-circuits-rm_b200/csrc/poseidon_upload.cuh"
using namespace qpzk;

__global__ void k_chain(u64* st, int nperm, int variant) {
  __shared__ u64 xch[2][COOP_XCH_WORDS];
  const u32 g = threadIdx.x >> 4, lane = threadIdx.x & 15;
  u64 s = lane < 12 ? st[lane] : 0;
  for (int i = 0; i < nperm; i++) s = poseidon_permute_coop(s, lane, xch[g]);
  if (g == 0 && lane < 12) st[lane] = gl_canon(s);
}
// Many resident groups (the tree-climb regime: 4096 leaves x 17 permutations): every 16-lane group runs its own
// chain. half_mask = 1: each half of a warp synchronises with its own 16-bit mask (k_tree_climb), 0: one
// full-warp mask. Measured on B200: no difference (the halves stay in lockstep either way); what matters is the
// number of resident warps per scheduler - 10.2 us per step at one warp per scheduler (148 blocks), 21.3 us at
// 3.5 (512 blocks = 4096 groups), 34.4 us at 7: the 16-lane permutation saturates a scheduler at about two warps.
__global__ void __launch_bounds__(128) k_chain_many(u64* st, int nperm, int half_mask) {
  __shared__ u64 xch[8][COOP_XCH_WORDS];
  const u32 g = threadIdx.x >> 4, lane = threadIdx.x & 15;
  const u32 mask = half_mask ? 0xffffu << (threadIdx.x & 16) : 0xffffffffu;
  u64 s = lane < 12 ? st[lane] + blockIdx.x * 8 + g : 0;
  for (int i = 0; i < nperm; i++) s = poseidon_permute_coop(s, lane, xch[g], mask);
  if (s == 0x123456789ULL) st[lane] = s;
}
__global__ void k_chain_thread(u64* st, int nperm) {
  u64 s[12];
  for (int i = 0; i < 12; i++) s[i] = st[i];
  for (int i = 0; i < nperm; i++) poseidon_permute(s);
  for (int i = 0; i < 12; i++) st[i] = gl_canon(s[i]);
}
static void host_permute(const PoseidonTablesHost& T, u64* s) {
  for (int r = 0; r < 30; r++) {
    for (int i = 0; i < 12; i++) s[i] = glh::add(s[i], T.rc[12 * r + i]);
    bool full = r < 4 || r >= 26;
    for (int i = 0; i < (full ? 12 : 1); i++) {
      u64 x = s[i], x2 = glh::mul(x, x), x4 = glh::mul(x2, x2), x3 = glh::mul(x, x2);
      s[i] = glh::mul(x3, x4);
    }
    u64 o[12];
    for (int rr = 0; rr < 12; rr++) {
      u64 acc = 0;
      for (int c = 0; c < 12; c++) {
        u64 m = kMdsCirc[((c - rr) % 12 + 12) % 12] + ((rr == c && rr == 0) ? kMdsDiag0 : 0);
        acc = glh::add(acc, glh::mul(m, s[c]));
      }
      o[rr] = acc;
    }
    for (int i = 0; i < 12; i++) s[i] = o[i];
  }
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e)); return 2; } } while (0)
int main(int argc, char** argv) {
  PoseidonTablesHost* T = new PoseidonTablesHost();
  build_poseidon_tables(T, PV_DENSE_PARTIAL);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { printf("no GPU\n"); return 0; }
  CK(poseidon_upload_tables(*T));
  const int nperm = 200;
  u64 h[12], want[12];
  for (int i = 0; i < 12; i++) want[i] = h[i] = 0x9E3779B97F4A7C15ULL * (i + 1) % GL_P;
  for (int i = 0; i < nperm; i++) host_permute(*T, want);
  u64* d;
  CK(cudaMalloc(&d, 96));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; mode++) {
    float best = 1e30f;
    u64 got[12];
    for (int it = 0; it < 5; it++) {
      CK(cudaMemcpy(d, h, 96, cudaMemcpyHostToDevice));
      cudaEventRecord(e0);
      if (mode == 0) k_chain<<<1, 32>>>(d, nperm, 0);
      else k_chain_thread<<<1, 32>>>(d, nperm);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    CK(cudaMemcpy(got, d, 96, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < 12; i++) bad += got[i] != want[i];
    printf("%s %s: %s  %.2f us per dependent permutation\n", argc > 1 ? argv[1] : "variant", mode == 0 ? "coop16" : "thread", bad ? "MISMATCH" : "ok",
           best * 1e3 / nperm);
  }
  for (int hm = 0; hm < 2; hm++)
    for (int blocks : {148, 512, 1024}) {
      float best = 1e30f;
      for (int it = 0; it < 3; it++) {
        cudaEventRecord(e0);
        k_chain_many<<<blocks, 128>>>(d, 17, hm);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      printf("%d blocks x 8 groups x 17 dependent permutations, %s mask: %.1f us (%.2f us per step)\n", blocks,
             hm ? "half-warp" : "full-warp", best * 1e3, best * 1e3 / 17);
    }
  return 0;
}
