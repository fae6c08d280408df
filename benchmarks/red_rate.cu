// red.global.add.f64 throughput vs address pattern (what bounds the atomic scatter strategies).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void red_add(double* p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
// pattern 0: 32 lanes -> 32 consecutive doubles; 1: runs of 3 consecutive doubles, runs scattered; 2: every lane scattered
__global__ void k(double* buf, long long n, int pattern, int iters) {
  unsigned long long s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761ull + 12345;
  const int lane = threadIdx.x & 31;
  for (int i = 0; i < iters; ++i) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    unsigned long long w = __shfl_sync(0xffffffffu, s, 0);  // warp-uniform random
    long long idx;
    if (pattern == 0) idx = (long long)((w >> 20) % (n - 64)) + lane;
    else if (pattern == 1) { unsigned long long g = __shfl_sync(0xffffffffu, s, (lane / 3) * 3); idx = (long long)((g >> 20) % (n - 8)) + lane % 3; }
    else idx = (long long)((s >> 20) % n);
    red_add(buf + idx, 1.0);
  }
}
int main() {
  const long long n = 1ll << 27;  // 1 GiB of doubles: far larger than L2
  double* buf; cudaMalloc(&buf, n * 8); cudaMemset(buf, 0, n * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("{");
  for (int pat = 0; pat < 3; ++pat)
    for (int small = 0; small < 2; ++small) {
      const long long nn = small ? (1ll << 22) : n;  // 32 MiB (L2 resident) or 1 GiB
      const int blocks = 148 * 8, threads = 256, iters = 2000;
      k<<<blocks, threads>>>(buf, nn, pat, 100); cudaDeviceSynchronize();
      cudaEventRecord(e0); k<<<blocks, threads>>>(buf, nn, pat, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("%s\"pattern%d_%s_Gred_per_s\": %.1f", (pat || small) ? ", " : "", pat, small ? "L2" : "HBM", (double)blocks * threads * iters / (ms * 1e-3) / 1e9);
    }
  printf("}\n");
  return 0;
}
