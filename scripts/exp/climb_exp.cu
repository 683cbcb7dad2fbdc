// Experiment harness (not product code): k_tree_climb<false> alone on random digests, with a per-level timeline
// (latest completion time of each level, from %globaltimer) - where do the 194 us of a 4096-node climb go?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#ifndef CLIMB_TL
#define CLIMB_TL 1
#endif
__device__ unsigned long long g_tl[64];
__device__ unsigned long long g_t0;
__device__ unsigned long long g_pmax[64], g_psum[64];
#include "../../qp-zk-circuits-rm_b200/csrc/merkle.cuh"
#include "../../qp-zk-circuits-rm_b200/csrc/poseidon_upload.cuh"
using namespace qpzk;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e)); return 2; } } while (0)
__global__ void k_t0() { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_t0)); for (int i = 0; i < 64; i++) g_tl[i] = g_pmax[i] = g_psum[i] = 0; }
int main(int argc, char** argv) {
  int log_n = argc > 1 ? atoi(argv[1]) : 13;
  PoseidonTablesHost* T = new PoseidonTablesHost();
  build_poseidon_tables(T, PV_DENSE_PARTIAL);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { printf("no GPU\n"); return 0; }
  CK(poseidon_upload_tables(*T));
  const u64 N = 1ull << log_n;
  u64* levels; u32* counters;
  CK(cudaMalloc(&levels, N * 2 * 32));
  CK(cudaMalloc(&counters, 2 * QPZK_CLIMB_MAX_START * 4));
  CK(cudaMemset(counters, 0, 2 * QPZK_CLIMB_MAX_START * 4));
  std::vector<u64> h(N * 4);
  for (size_t i = 0; i < h.size(); i++) h[i] = (0x9E3779B97F4A7C15ULL * (i + 1)) % GL_P;
  CK(cudaMemcpy(levels, h.data(), N * 32, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const u64 nout = N / 2;
  for (int it = 0; it < 4; it++) {
    k_t0<<<1, 1>>>();
    cudaEventRecord(e0);
    k_tree_climb<false><<<(unsigned)((nout + QPZK_COOP_GROUPS - 1) / QPZK_COOP_GROUPS), QPZK_COOP_THREADS>>>(
        nullptr, 0, 0, 0, levels, log_n, 4, 0, 0, nout, counters);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long tl[64], t0;
    CK(cudaMemcpyFromSymbol(tl, g_tl, sizeof tl));
    CK(cudaMemcpyFromSymbol(&t0, g_t0, 8));
    printf("%s: %llu nodes, %d threads/CTA: %.1f us; level done at (us):", argc > 2 ? argv[2] : "", (unsigned long long)nout, QPZK_COOP_THREADS, ms * 1e3);
    for (int l = 1; l <= log_n - 4; l++) printf(" %.1f", tl[l] ? (tl[l] - t0) * 1e-3 : -1.0);
    printf("\n");
    unsigned long long pm[64], ps[64];
    CK(cudaMemcpyFromSymbol(pm, g_pmax, sizeof pm));
    CK(cudaMemcpyFromSymbol(ps, g_psum, sizeof ps));
    printf("   permutation time per level, max / mean (us):");
    for (int l = 2; l <= log_n - 4; l++) printf(" %.1f/%.1f", pm[l] * 1e-3, ps[l] * 1e-3 / (double)(N >> l));
    printf("\n");
  }
  return 0;
}
