// Experiment harness (not product code): single-warp dependent-chain latencies (cycles) of the building blocks of the
// 16-lane permutation, to see what a round's critical path is made of.
#include <cuda_runtime.h>
#include <cstdio>
#include "../../qp-zk-circuits-rm_b200/csrc/poseidon.cuh"
#include "../../qp-zk-circuits-rm_b200/csrc/poseidon_upload.cuh"
using namespace qpzk;

template <int OP>
__global__ void k_lat(u64* out, long long* cyc, int n, u64 seed) {
  __shared__ u64 xch[COOP_XCH_WORDS];
  const u32 lane = threadIdx.x & 15;
  u64 s = seed + threadIdx.x, b = seed * 3 + 1;
  if (lane < 12) { xch[lane] = s; xch[lane + 12] = s; }
  __syncwarp();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    if (OP == 0) s = gl_mul(s, b);
    if (OP == 1) s = sbox7(s);
    if (OP == 2) s = gl_mad(s, b, s);
    if (OP == 3) s = coop_group_sum(s, 0xffffffffu);
    if (OP == 4) s = coop_shfl64(s, 0, 0xffffffffu) + lane;
    if (OP == 5) s = gl_sqr(s);
    if (OP == 6) s = gl_add(s, b);
    if (OP == 7) {  // the exchange + circulant row of a full round, without the s-box
      u64* buf = xch + (i & 1) * 24;
      if (lane < 12) { buf[lane] = s; buf[lane + 12] = s; }
      __syncwarp();
      const u64* row = buf + (lane < 12 ? lane : 0);
      u32 al0 = 0, al1 = 0, ah0 = 0, ah1 = 0, bl0 = 0, bl1 = 0, bh0 = 0, bh1 = 0;
#pragma unroll
      for (int j = 0; j < 12; j += 2) {
        const u64 v0 = row[j], v1 = row[j + 1];
        mad_wide(al0, al1, (u32)v0, c_mds_circ[j]);
        mad_wide(ah0, ah1, (u32)(v0 >> 32), c_mds_circ[j]);
        mad_wide(bl0, bl1, (u32)v1, c_mds_circ[j + 1]);
        mad_wide(bh0, bh1, (u32)(v1 >> 32), c_mds_circ[j + 1]);
      }
      u64 lo = (((u64)al1 << 32) | al0) + (((u64)bl1 << 32) | bl0);
      u64 hi = (((u64)ah1 << 32) | ah0) + (((u64)bh1 << 32) | bh0);
      s = mds_combine((u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32));
    }
    if (OP == 8) s = gl_mul_alu(s, b);
    if (OP == 9) {  // two independent s-boxes: does a second chain ride for free?
      s = sbox7(s);
      b = sbox7(b);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  out[threadIdx.x] = s + b;
}
int main() {
  PoseidonTablesHost* T = new PoseidonTablesHost();
  build_poseidon_tables(T, PV_DENSE_PARTIAL);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { printf("no GPU\n"); return 0; }
  poseidon_upload_tables(*T);
  u64* d; long long* c;
  cudaMalloc(&d, 32 * 8); cudaMalloc(&c, 8);
  const char* names[] = {"gl_mul", "sbox7", "gl_mad", "group_sum", "shfl64", "gl_sqr", "gl_add", "exchange+row", "gl_mul_alu", "2 x sbox7"};
  const int n = 2000;
#define RUN(OP) { long long best = 1ll << 60, h; for (int it = 0; it < 3; it++) { k_lat<OP><<<1, 32>>>(d, c, n, 12345); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); if (h < best) best = h; } \
    printf("%-14s %7.1f cycles per dependent op\n", names[OP], (double)best / n); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
  return 0;
}
