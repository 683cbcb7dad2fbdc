// Experiment harness (not product code): per-iteration times of the independent-worker staged upload, to find where the
// occasional 200-1200 ms first iterations of scripts/prof_big_upload.py come from.
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || !ndev) { printf("no GPU\n"); return 0; }
  const int T = argc > 1 ? atoi(argv[1]) : 4;
  const int fresh = argc > 2 ? atoi(argv[2]) : 0;   // 1: rewrite the source before every iteration; 2: new allocation every iteration
  const size_t bytes = (size_t)142 << 20, CH = (size_t)2 << 20;
  char* pg = (char*)malloc(bytes); memset(pg, 1, bytes);
  char* dev; cudaMalloc(&dev, bytes);
  char* arena; cudaHostAlloc(&arena, 2 * T * CH, cudaHostAllocDefault);
  std::vector<cudaEvent_t> ev(2 * T);
  for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  printf("T=%d fresh=%d:", T, fresh);
  for (int it = 0; it < 12; it++) {
    if (fresh == 1) memset(pg, it, bytes);
    if (fresh == 2) { free(pg); pg = (char*)malloc(bytes); memset(pg, it, bytes); }
    cudaStreamSynchronize(st);
    double t0 = now();
    const size_t per = (((bytes + T - 1) / T) + 4095) & ~(size_t)4095;
    auto work = [&](int t) {
      size_t beg = std::min(bytes, t * per), end = std::min(bytes, beg + per);
      if (t) cudaSetDevice(0);
      int j = 0;
      for (size_t off = beg; off < end; off += CH, j ^= 1) {
        size_t len = std::min(CH, end - off);
        char* buf = arena + (2 * t + j) * CH;
        cudaEventSynchronize(ev[2 * t + j]);
        memcpy(buf, pg + off, len);
        cudaMemcpyAsync(dev + off, buf, len, cudaMemcpyHostToDevice, st);
        cudaEventRecord(ev[2 * t + j], st);
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    double t1 = now();
    cudaStreamSynchronize(st);
    printf(" %.1f(%.1f)", (now() - t0) * 1e3, (t1 - t0) * 1e3);
    fflush(stdout);
  }
  printf(" ms\n");
  return 0;
}
