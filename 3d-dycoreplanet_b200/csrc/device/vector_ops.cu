// Device-resident vector algebra for the Krylov solvers that call the SpMVs: the Trilinos vector
// operations inside deal.II's SolverCG / SolverGMRES / SolverFGMRES (l2_norm, operator*, add, sadd, equ;
// e.g. /root/reference/include/core/boussinesq_model.tpp:1165, 1426-1440, include/linear_algebra/inverse_matrix.hpp:99).
// Dot products use a fixed two-stage tree, so results are bit-reproducible from run to run.
#include "dcp_internal.cuh"

namespace {
constexpr int DOT_BLOCKS = 592, DOT_THREADS = 256;

__global__ void __launch_bounds__(DOT_THREADS) dot_stage1(long long n, const double* __restrict__ x,
                                                          const double* __restrict__ y, double* __restrict__ partial) {
  __shared__ double s[DOT_THREADS / 32];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += x[i] * y[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < DOT_THREADS / 32; ++w) t += s[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(DOT_THREADS) dot_stage2(int nb, const double* __restrict__ partial, double* __restrict__ out) {
  __shared__ double s[DOT_THREADS];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = DOT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = s[0];
}
__global__ void axpy_kernel(long long n, double a, const double* __restrict__ x, double* __restrict__ y) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] += a * x[i];
}
__global__ void sadd_kernel(long long n, double s, double a, const double* __restrict__ x, double* __restrict__ y) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = s * y[i] + a * x[i];
}
__global__ void shift_kernel(long long n, double a, double* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += a;
}
__global__ void scale_kernel(long long n, double a, double* __restrict__ y) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] *= a;
}
// y += sign * (*a) * x with the scalar in device memory
__global__ void axpy_dev_kernel(long long n, const double* __restrict__ a_dev, double sign, const double* __restrict__ x, double* __restrict__ y) {
  const double a = sign * *a_dev;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] += a * x[i];
}
inline unsigned vgrid(dcp_ctx* ctx, long long n) {
  long long b = (n + 255) / 256, cap = (long long)ctx->sm_count * 8;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}
}  // namespace

extern "C" {

int dcp_vec_dot(dcp_ctx* ctx, int64_t n, const double* x_dev, const double* y_dev, double* result_host) {
  if (!ctx || !result_host || n < 0) return DCP_ERR_ARG;
  if (!ctx->dot_scratch) {
    DCP_CUDA(cudaMalloc((void**)&ctx->dot_scratch, sizeof(double) * (DOT_BLOCKS + 1)));
    DCP_CUDA(cudaMallocHost((void**)&ctx->dot_host, sizeof(double)));
  }
  if (n == 0) {
    *result_host = 0.0;
    return DCP_OK;
  }
  dot_stage1<<<DOT_BLOCKS, DOT_THREADS, 0, ctx->stream>>>(n, x_dev, y_dev, ctx->dot_scratch);
  dot_stage2<<<1, DOT_THREADS, 0, ctx->stream>>>(DOT_BLOCKS, ctx->dot_scratch, ctx->dot_scratch + DOT_BLOCKS);
  ctx->launches += 2;
  DCP_CUDA(cudaMemcpyAsync(ctx->dot_host, ctx->dot_scratch + DOT_BLOCKS, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  // the scatter-error counters of the assemblers travel with the result: in device-resident mode a dropped matrix entry
  // is reported by the next synchronising call instead of never (dcp.h, "deferred error reports")
  DCP_CUDA(cudaMemcpyAsync(ctx->h_err, ctx->d_err, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  *result_host = *ctx->dot_host;
  if (ctx->h_err[0] != 0 || ctx->h_err[1] != 0) return dcp_check_device_errors(ctx, "an earlier assembly call");
  return DCP_OK;
}

int dcp_vec_mgs(dcp_ctx* ctx, int64_t n, int k, const double* const* v_dev, double* w_dev, double* h_host) {
  if (!ctx || n < 0 || k < 0 || k > DCP_MGS_MAX || (k > 0 && !v_dev) || !w_dev || !h_host) return DCP_ERR_ARG;
  DCP_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->mgs_scalars) {
    DCP_CUDA(cudaMalloc((void**)&ctx->mgs_scalars, sizeof(double) * (DCP_MGS_MAX + 1 + DOT_BLOCKS)));
    DCP_CUDA(cudaMallocHost((void**)&ctx->mgs_host, sizeof(double) * (DCP_MGS_MAX + 1)));
  }
  if (n == 0) {
    for (int i = 0; i <= k; ++i) h_host[i] = 0.0;
    return DCP_OK;
  }
  double* sc = ctx->mgs_scalars;
  double* partial = sc + DCP_MGS_MAX + 1;
  for (int i = 0; i < k; ++i) {
    dot_stage1<<<DOT_BLOCKS, DOT_THREADS, 0, ctx->stream>>>(n, w_dev, v_dev[i], partial);
    dot_stage2<<<1, DOT_THREADS, 0, ctx->stream>>>(DOT_BLOCKS, partial, sc + i);
    axpy_dev_kernel<<<vgrid(ctx, n), 256, 0, ctx->stream>>>(n, sc + i, -1.0, v_dev[i], w_dev);
  }
  dot_stage1<<<DOT_BLOCKS, DOT_THREADS, 0, ctx->stream>>>(n, w_dev, w_dev, partial);
  dot_stage2<<<1, DOT_THREADS, 0, ctx->stream>>>(DOT_BLOCKS, partial, sc + k);
  ctx->launches += 3 * k + 2;
  DCP_CUDA(cudaMemcpyAsync(ctx->mgs_host, sc, sizeof(double) * (size_t)(k + 1), cudaMemcpyDeviceToHost, ctx->stream));
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i <= k; ++i) h_host[i] = ctx->mgs_host[i];
  return DCP_OK;
}

int dcp_vec_axpy(dcp_ctx* ctx, int64_t n, double a, const double* x_dev, double* y_dev) {
  if (!ctx || n < 0) return DCP_ERR_ARG;
  if (n == 0) return DCP_OK;
  axpy_kernel<<<vgrid(ctx, n), 256, 0, ctx->stream>>>(n, a, x_dev, y_dev);
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

int dcp_vec_sadd(dcp_ctx* ctx, int64_t n, double s, double a, const double* x_dev, double* y_dev) {
  if (!ctx || n < 0) return DCP_ERR_ARG;
  if (n == 0) return DCP_OK;
  sadd_kernel<<<vgrid(ctx, n), 256, 0, ctx->stream>>>(n, s, a, x_dev, y_dev);
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

int dcp_vec_scale(dcp_ctx* ctx, int64_t n, double a, double* y_dev) {
  if (!ctx || n < 0) return DCP_ERR_ARG;
  if (n == 0) return DCP_OK;
  scale_kernel<<<vgrid(ctx, n), 256, 0, ctx->stream>>>(n, a, y_dev);
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

int dcp_vec_fill(dcp_ctx* ctx, int64_t n, double value, double* y_dev) {
  if (!ctx || n < 0 || (n > 0 && !y_dev)) return DCP_ERR_ARG;
  DCP_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return DCP_OK;
  return dcp_launch_fill(ctx, y_dev, n, value);
}

int dcp_vec_shift(dcp_ctx* ctx, int64_t n, double a, double* y_dev) {
  if (!ctx || n < 0 || (n > 0 && !y_dev)) return DCP_ERR_ARG;
  DCP_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return DCP_OK;
  shift_kernel<<<vgrid(ctx, n), 256, 0, ctx->stream>>>(n, a, y_dev);
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

int dcp_vec_copy(dcp_ctx* ctx, int64_t n, const double* x_dev, double* y_dev) {
  if (!ctx || n < 0) return DCP_ERR_ARG;
  if (n == 0) return DCP_OK;
  DCP_CUDA(cudaMemcpyAsync(y_dev, x_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
  return DCP_OK;
}

}  // extern "C"
