// Microbenchmark (not product code): issue rates of the integer instructions the field arithmetic is
// made of, per SM per clock, on dependency chains the compiler cannot hoist. Check the SASS with
// cuobjdump before trusting a row.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
typedef unsigned int u32;

#define CHAINS 8
#define UNROLL 16

template <int KIND>
__global__ void __launch_bounds__(256) k(u64* out, int iters, u32 a, u32 b) {
  u32 lo[CHAINS], hi[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) { lo[i] = threadIdx.x * 7 + i; hi[i] = blockIdx.x + i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
#pragma unroll
      for (int i = 0; i < CHAINS; i++) {
        if (KIND == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(a), "r"(b));
        if (KIND == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(a), "r"(b));
        if (KIND == 2) { u64 t = (u64)hi[i] * a; lo[i] ^= (u32)t; hi[i] = (u32)(t >> 32); }
        if (KIND == 3) { u64 t = ((u64)hi[i] << 32) | lo[i]; t = (u64)hi[(i + 1) % CHAINS] * a + t; lo[i] = (u32)t; hi[i] = (u32)(t >> 32); }
        if (KIND == 4) asm volatile("mad.lo.cc.u32 %0, %0, %2, %0; madc.hi.cc.u32 %1, %0, %2, %1; addc.u32 %1, %1, 0;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a));
        if (KIND == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[i]) : "r"(hi[i]));
        if (KIND == 6) {  // wide-acc + independent add
          asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0,%1}, t;}" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a));
        }
        if (KIND == 7) {  // mad.lo + add on separate chains
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(a), "r"(b));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(hi[i]) : "r"(b));
        }
        if (KIND == 8) {  // wide (no acc) + 2 adds on other regs
          asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0,%1}, t;}" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a));
        }
        if (KIND == 9) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %2;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(b));
        if (KIND == 10) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(lo[i]) : "r"(hi[i]));
        if (KIND == 11) asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(lo[i]) : "r"(b));   // IMAD as an adder
      }
      if (KIND == 6 || KIND == 8) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
          // independent ALU work riding along: uses neither lo nor hi as destination of the wide chain
        }
      }
    }
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += ((u64)hi[i] << 32) ^ lo[i];
  if (s == 0x123456789ULL) out[0] = s;
}

template <int KIND>
static void run(const char* name, double ops_per_inner, u64* d, int sms) {
  const int iters = 512, threads = 256, blocks = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k<KIND><<<blocks, threads>>>(d, iters, 12345u + rep, 777u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  double ops = (double)blocks * threads * iters * UNROLL * CHAINS * ops_per_inner;
  printf("%-44s %8.3f ms  %7.2f T thread-ops/s  %6.1f per clk per SM (at 1.965 GHz)\n", name, best, ops / best / 1e9,
         ops / (best * 1e-3) / 1.965e9 / sms);
}

int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || !ndev) { printf("no GPU\n"); return 0; }
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  u64* d; cudaMalloc(&d, 64);
  int sms = p.multiProcessorCount;
  run<0>("IMAD (mad.lo.u32)", 1, d, sms);
  run<1>("IMAD.HI (mad.hi.u32)", 1, d, sms);
  run<2>("IMAD.WIDE no addend (mul.wide.u32)", 1, d, sms);
  run<3>("IMAD.WIDE 64-bit addend (mad.wide.u32)", 1, d, sms);
  run<4>("IMAD.WIDE carry-out + carry add (3 PTX)", 1, d, sms);
  run<5>("IADD3 (add.u32)", 1, d, sms);
  run<7>("IMAD + IADD3 pair (counted as 2)", 2, d, sms);
  run<9>("IADD3 + IADD3.X pair (counted as 2)", 2, d, sms);
  run<10>("SHF (funnel shift)", 1, d, sms);
  run<11>("IMAD x*1+b (IMAD as adder)", 1, d, sms);
  return 0;
}
