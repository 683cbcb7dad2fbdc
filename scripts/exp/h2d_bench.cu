// Experiment harness (not product code): host -> device bandwidth from pageable memory - the driver's own staging against
// a pipelined copy through two pinned buffers filled by several host threads.
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void par_memcpy(char* dst, const char* src, size_t n, int T) {
  if (T <= 1) { memcpy(dst, src, n); return; }
  std::vector<std::thread> th;
  size_t per = (n + T - 1) / T;
  for (int t = 0; t < T; t++) {
    size_t o = t * per; if (o >= n) break;
    size_t l = std::min(per, n - o);
    th.emplace_back([=] { memcpy(dst + o, src + o, l); });
  }
  for (auto& t : th) t.join();
}
int main(int argc, char** argv) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || !ndev) { printf("no GPU\n"); return 0; }
  printf("host threads: %u\n", std::thread::hardware_concurrency());
  for (size_t mb : {18, 71, 283}) {
    size_t bytes = mb << 20;
    char* pg = (char*)malloc(bytes); memset(pg, 1, bytes);
    char* pin; cudaMallocHost(&pin, bytes); memset(pin, 2, bytes);
    char* dev; cudaMalloc(&dev, bytes);
    cudaStream_t st; cudaStreamCreate(&st);
    auto time_it = [&](auto f) { double best = 1e9; for (int i = 0; i < 4; i++) { cudaStreamSynchronize(st); double t0 = now(); f(); cudaStreamSynchronize(st); best = std::min(best, now() - t0); } return best; };
    double tp = time_it([&] { cudaMemcpyAsync(dev, pg, bytes, cudaMemcpyHostToDevice, st); });
    double tn = time_it([&] { cudaMemcpyAsync(dev, pin, bytes, cudaMemcpyHostToDevice, st); });
    printf("%4zu MB: pageable %.2f ms (%.1f GB/s)  pinned %.2f ms (%.1f GB/s)\n", mb, tp * 1e3, bytes / tp / 1e9, tn * 1e3, bytes / tn / 1e9);
    for (size_t chunk_mb : {2, 4, 8}) for (int T : {1, 2, 4, 8}) {
      size_t chunk = chunk_mb << 20;
      char* stg[2]; cudaEvent_t ev[2];
      for (int i = 0; i < 2; i++) { cudaMallocHost(&stg[i], chunk); cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming); }
      // persistent workers: each copies its slice of the current chunk
      double ts = time_it([&] {
        int k = 0;
        for (size_t o = 0; o < bytes; o += chunk, k ^= 1) {
          size_t l = std::min(chunk, bytes - o);
          cudaEventSynchronize(ev[k]);
          par_memcpy(stg[k], pg + o, l, T);
          cudaMemcpyAsync(dev + o, stg[k], l, cudaMemcpyHostToDevice, st);
          cudaEventRecord(ev[k], st);
        }
      });
      printf("         staged chunk %zu MB x %d threads: %.2f ms (%.1f GB/s)\n", chunk_mb, T, ts * 1e3, bytes / ts / 1e9);
      for (int i = 0; i < 2; i++) { cudaFreeHost(stg[i]); cudaEventDestroy(ev[i]); }
    }
    free(pg); cudaFreeHost(pin); cudaFree(dev); cudaStreamDestroy(st);
  }
  return 0;
}
