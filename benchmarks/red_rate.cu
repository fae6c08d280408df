// red.global.add.f64 and st.global.f64 throughput vs address pattern (what bounds the atomic scatter strategies).
// Index generation is kept to a handful of integer instructions (xorshift + mask) so that the memory pipe, not
// the ALU, is measured.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void red_add(double* p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
// pattern 0: 32 lanes -> 32 consecutive doubles; 1: runs of 3 consecutive doubles, runs scattered; 2: every lane scattered
template <int MODE, int PATTERN>  // MODE 0: red.add, 1: plain store
__global__ void k(double* buf, unsigned mask, int iters) {
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  const int lane = threadIdx.x & 31;
  if (PATTERN == 0) s = (blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * 2654435761u + 12345u;      // warp-uniform
  if (PATTERN == 1) s = (blockIdx.x * blockDim.x + threadIdx.x - lane + (lane / 3) * 3) * 2654435761u + 12345u;  // uniform per run of 3
  for (int i = 0; i < iters; ++i) {
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    unsigned idx;
    if (PATTERN == 0) idx = (s & mask & ~63u) + lane;
    else if (PATTERN == 1) idx = (s & mask & ~3u) + lane % 3;
    else idx = s & mask;
    if (MODE == 0) red_add(buf + idx, 1.0);
    else asm volatile("st.global.f64 [%0], %1;" ::"l"(buf + idx), "d"((double)i) : "memory");
  }
}
template <int MODE, int PATTERN>
float run(double* buf, unsigned mask, int iters) {
  const int blocks = 148 * 8, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, PATTERN><<<blocks, threads>>>(buf, mask, 100); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE, PATTERN><<<blocks, threads>>>(buf, mask, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return (float)((double)blocks * threads * iters / (ms * 1e-3) / 1e9);
}
int main() {
  const long long n = 1ll << 27;  // 1 GiB of doubles: far larger than L2
  double* buf; cudaMalloc(&buf, n * 8); cudaMemset(buf, 0, n * 8);
  const unsigned big = (1u << 27) - 1, small = (1u << 22) - 1;  // 1 GiB / 32 MiB (L2 resident)
  const int it = 2000;
  printf("{\"unit\": \"G ops/s\"");
#define ROW(MODE, PAT, name) printf(", \"%s_pattern%d_HBM\": %.1f, \"%s_pattern%d_L2\": %.1f", name, PAT, run<MODE, PAT>(buf, big, it), name, PAT, run<MODE, PAT>(buf, small, it));
  ROW(0, 0, "red") ROW(0, 1, "red") ROW(0, 2, "red") ROW(1, 0, "store") ROW(1, 1, "store") ROW(1, 2, "store")
  printf("}\n");
  return 0;
}
