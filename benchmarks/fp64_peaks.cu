// Register-resident FP64 microbenchmarks for the roofline denominators of the FP64-bound assembly kernels:
//   (1) DFMA peak (independent FMA chains), (2) DMMA mma.sync.m8n8k4.f64 peak, (3) shared-memory LDS.64/LDS.128 rate.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peaks fp64_peaks.cu ; prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  double fa = a + threadIdx.x * 1e-9, fb = b;
  for (int i = 0; i < iters; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(fa), "d"(fb));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(fa), "d"(fb));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(fa), "d"(fb));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(fa), "d"(fb));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

template <int VEC>
__global__ void lds_kernel(double* out, int iters) {
  __shared__ double s[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i;
  __syncthreads();
  double acc = 0;
  int idx = threadIdx.x * VEC;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (VEC == 1) acc += s[(idx + u * 512) & 4095];
      else { double2 v = *reinterpret_cast<double2*>(&s[(idx + u * 512) & 4094]); acc += v.x + v.y; }
    }
    idx = (idx + 8) & 4095;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, threads = 512, blocks = sms * 4, iters = 20000;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  float t1 = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  double dfma_tf = 2.0 * 8 * (double)iters * blocks * threads / (t1 * 1e-3) / 1e12;
  float t2 = time_ms([&] { dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  double dmma_tf = 2.0 * 4 * 256.0 * (double)iters * blocks * (threads / 32) / (t2 * 1e-3) / 1e12;
  float t3 = time_ms([&] { lds_kernel<1><<<blocks, threads>>>(out, 4000); });
  float t4 = time_ms([&] { lds_kernel<2><<<blocks, threads>>>(out, 4000); });
  double lds64 = 8.0 * 8 * 4000.0 * blocks * threads / (t3 * 1e-3) / 1e12, lds128 = 16.0 * 8 * 4000.0 * blocks * threads / (t4 * 1e-3) / 1e12;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, \"lds64_TBps\": %.2f, \"lds128_TBps\": %.2f}\n",
         p.name, sms, dfma_tf, dmma_tf, lds64, lds128);
  return 0;
}
