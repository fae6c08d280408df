// Device-resident conjugate gradients: deal.II's SolverCG<Vector>::solve(A, x, b, P) with
// SolverControl(max_steps, tol) as it runs behind LinearAlgebra::InverseMatrix::vmult
// (/root/reference/include/linear_algebra/inverse_matrix.hpp:90-121), ApproximateInverseMatrix
// (approximate_inverse.hpp:97-128) and the temperature solve (include/core/boussinesq_model.tpp:1426-1440).
//
// The reference evaluates two inner products and a norm per iteration through MPI_Allreduce, i.e. the host sees every
// scalar.  Here alpha, beta and the residual stay in device memory: every kernel of an iteration reads them there, a
// flag records convergence (iterations that are enqueued behind it are no-ops, so result and step count are those of
// the unbatched algorithm), and the host looks at the flag once per `check_every` iterations.  Inner products use the
// same fixed two-stage tree as dcp_vec_dot (bit-reproducible, and bit-identical to a loop that calls dcp_vec_dot).
#include <algorithm>
#include <cmath>

#include "dcp_internal.cuh"

namespace {

constexpr int DOT_BLOCKS = 592, DOT_THREADS = 256;   // as in vector_ops.cu

struct CgState {
  double* sc;      // [0] r.z of the current direction, [1] p.Ap, [2] r.r, [3] r.z after the update
  int* flags;      // [0] 0: running, 1: converged, 2: step limit reached; [1] steps done
  double* partial; // [DOT_BLOCKS]
};

__device__ __forceinline__ void block_partial(double acc, double* __restrict__ partial) {
  __shared__ double s[DOT_THREADS / 32];
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
__device__ __forceinline__ double block_total(int nb, const double* __restrict__ partial) {
  __shared__ double s[DOT_THREADS];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = DOT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  return s[0];
}

__global__ void __launch_bounds__(DOT_THREADS) cg_dot1(long long n, const int* __restrict__ flags, const double* __restrict__ x,
                                                       const double* __restrict__ y, double* __restrict__ partial) {
  if (flags && flags[0]) return;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) acc += x[i] * y[i];
  block_partial(acc, partial);
}
// *out = sum; the r.z of the previous iteration becomes the current one first (mode 1: start of an iteration)
__global__ void __launch_bounds__(DOT_THREADS) cg_dot2(CgState st, int out, int roll) {
  if (st.flags[0]) return;
  const double t = block_total(DOT_BLOCKS, st.partial);
  if (threadIdx.x == 0) {
    if (roll) st.sc[0] = st.sc[3];
    st.sc[out] = t;
  }
}
// x += alpha p, r -= alpha Ap, partial sums of r.r
__global__ void __launch_bounds__(DOT_THREADS) cg_update(long long n, CgState st, const double* __restrict__ p, const double* __restrict__ Ap,
                                                         double* __restrict__ x, double* __restrict__ r) {
  if (st.flags[0]) return;
  const double alpha = st.sc[0] / st.sc[1];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    const double ri = r[i] + (-alpha) * Ap[i];
    r[i] = ri;
    acc += ri * ri;
  }
  block_partial(acc, st.partial);
}
// r.r, the step counter and the SolverControl check
__global__ void __launch_bounds__(DOT_THREADS) cg_check(CgState st, double tol, long long max_steps) {
  if (st.flags[0]) return;
  const double t = block_total(DOT_BLOCKS, st.partial);
  if (threadIdx.x == 0) {
    st.sc[2] = t;
    const int steps = ++st.flags[1];
    if (sqrt(t) <= tol) st.flags[0] = 1;
    else if (steps >= max_steps) st.flags[0] = 2;
  }
}
// z = P r for the diagonal preconditioners (dinv == nullptr: identity), partial sums of r.z; apply == 0: z was made by
// another preconditioner, only the inner product
__global__ void __launch_bounds__(DOT_THREADS) cg_precondition(long long n, CgState st, const double* __restrict__ dinv, int apply,
                                                               const double* __restrict__ r, double* __restrict__ z) {
  if (st.flags[0]) return;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double ri = r[i];
    double zi;
    if (apply) {
      zi = dinv ? dinv[i] * ri : ri;
      z[i] = zi;
    } else
      zi = z[i];
    acc += ri * zi;
  }
  block_partial(acc, st.partial);
}
// p = z + beta p
__global__ void cg_direction(long long n, CgState st, const double* __restrict__ z, double* __restrict__ p) {
  if (st.flags[0]) return;
  const double beta = st.sc[3] / st.sc[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = beta * p[i] + z[i];
}
// r = b - r
__global__ void cg_residual0(long long n, const double* __restrict__ b, double* __restrict__ r) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) r[i] = -1.0 * r[i] + 1.0 * b[i];
}

inline unsigned vgrid(dcp_ctx* ctx, long long n) {
  long long b = (n + 255) / 256, cap = (long long)ctx->sm_count * 8;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int dcp_cg_solve(dcp_model* m, int which, int bi, int bj, int precond, int which_p, int bp, dcp_ilu* ilu, double* x_dev,
                            const double* b_dev, double tol, int64_t max_steps, int check_every, int64_t* last_step,
                            double* last_residual) {
  if (!m || !x_dev || !b_dev || !last_step || !last_residual || precond < DCP_PRECOND_IDENTITY || precond > DCP_PRECOND_ILU || max_steps < 0)
    return DCP_ERR_ARG;
  dcp_ctx* ctx = m->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  BlockMat* M = dcp_select_matrix(m, which);
  if (!M || bi < 0 || bj < 0 || bi >= M->nb || bj >= M->nb) {
    dcp_set_error("dcp_cg_solve: invalid matrix / block selector");
    return DCP_ERR_ARG;
  }
  const DevCsr& A = M->blk[bi][bj];
  if (A.n_rows != A.n_cols) {
    dcp_set_error("dcp_cg_solve: the block is not square");
    return DCP_ERR_ARG;
  }
  if (M->owned[bi] >= 0) {
    dcp_set_error("dcp_cg_solve: row-distributed model (the product needs the ghost exchange: use the halo operators)");
    return DCP_ERR_STATE;
  }
  const double* dinv = nullptr;
  if (precond == DCP_PRECOND_JACOBI) {
    BlockMat* Mp = dcp_select_matrix(m, which_p);
    if (!Mp || bp < 0 || bp >= Mp->nb || !Mp->diag_inv[bp] || Mp->blk[bp][bp].n_rows != A.n_rows) {
      dcp_set_error("dcp_cg_solve: no Jacobi diagonal of that size (assemble the matrix / call the Jacobi set-up first)");
      return DCP_ERR_STATE;
    }
    dinv = Mp->diag_inv[bp];
  }
  if (precond == DCP_PRECOND_ILU && !ilu) return DCP_ERR_ARG;
  const long long n = A.n_rows;
  *last_step = 0;
  *last_residual = 0.0;
  if (n == 0) return DCP_OK;
  if (check_every < 1) check_every = 8;

  // scratch: r, z, p, Ap, scalars, flags, partial sums
  double* buf = nullptr;
  int* flags = nullptr;
  struct HostView { int flags[2]; double rr; };
  HostView* h = nullptr;
  DCP_CUDA(cudaMalloc((void**)&buf, sizeof(double) * (4 * (size_t)n + 8 + DOT_BLOCKS)));
  if (cudaMalloc((void**)&flags, 2 * sizeof(int)) != cudaSuccess || cudaMallocHost((void**)&h, sizeof(HostView)) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(buf);
    cudaFree(flags);
    dcp_set_error("dcp_cg_solve: allocation failed");
    return DCP_ERR_CUDA;
  }
  double *r = buf, *z = buf + n, *p = buf + 2 * n, *Ap = buf + 3 * n;
  CgState st;
  st.sc = buf + 4 * n;
  st.flags = flags;
  st.partial = st.sc + 8;
  cudaStream_t s = ctx->stream;
  int rc = DCP_OK;
  auto finish = [&](int code) {
    cudaStreamSynchronize(s);
    cudaFree(buf);
    cudaFree(flags);
    cudaFreeHost(h);
    return code;
  };
  auto apply_P = [&]() -> int {   // z = P r and the partial sums of r.z
    if (precond == DCP_PRECOND_ILU) {
      const int e = dcp_ilu_vmult(ilu, z, r, DCP_DEVICE);
      if (e != DCP_OK) return e;
      cg_precondition<<<DOT_BLOCKS, DOT_THREADS, 0, s>>>(n, st, nullptr, 0, r, z);
    } else
      cg_precondition<<<DOT_BLOCKS, DOT_THREADS, 0, s>>>(n, st, dinv, 1, r, z);
    ctx->launches++;
    return DCP_OK;
  };
  cudaMemsetAsync(flags, 0, 2 * sizeof(int), s);
  cudaMemsetAsync(st.sc, 0, 8 * sizeof(double), s);
  // r = b - A x, first check (one host synchronisation)
  rc = dcp_launch_spmv(ctx, A, x_dev, r, false);
  if (rc != DCP_OK) return finish(rc);
  cg_residual0<<<vgrid(ctx, n), 256, 0, s>>>(n, b_dev, r);
  cg_dot1<<<DOT_BLOCKS, DOT_THREADS, 0, s>>>(n, nullptr, r, r, st.partial);
  cg_dot2<<<1, DOT_THREADS, 0, s>>>(st, 2, 0);
  ctx->launches += 3;
  if (cudaMemcpyAsync(&h->rr, st.sc + 2, sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
    return finish(DCP_ERR_CUDA);
  *last_residual = std::sqrt(h->rr);
  if (*last_residual <= tol) return finish(DCP_OK);
  // z = P r, p = z, r.z
  rc = apply_P();
  if (rc != DCP_OK) return finish(rc);
  cg_dot2<<<1, DOT_THREADS, 0, s>>>(st, 0, 0);
  cudaMemcpyAsync(p, z, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s);
  cudaMemcpyAsync(st.sc + 3, st.sc, sizeof(double), cudaMemcpyDeviceToDevice, s);
  ctx->launches++;
  int64_t enq = 0;
  for (;;) {
    const int64_t batch = std::min<int64_t>(check_every, max_steps - enq);
    if (batch <= 0) break;
    for (int64_t k = 0; k < batch; ++k) {
      rc = dcp_launch_spmv(ctx, A, p, Ap, false);
      if (rc != DCP_OK) return finish(rc);
      cg_dot1<<<DOT_BLOCKS, DOT_THREADS, 0, s>>>(n, flags, p, Ap, st.partial);
      cg_dot2<<<1, DOT_THREADS, 0, s>>>(st, 1, 1);
      cg_update<<<DOT_BLOCKS, DOT_THREADS, 0, s>>>(n, st, p, Ap, x_dev, r);
      cg_check<<<1, DOT_THREADS, 0, s>>>(st, tol, (long long)max_steps);
      rc = apply_P();
      if (rc != DCP_OK) return finish(rc);
      cg_dot2<<<1, DOT_THREADS, 0, s>>>(st, 3, 0);
      cg_direction<<<vgrid(ctx, n), 256, 0, s>>>(n, st, z, p);
      ctx->launches += 6;
    }
    enq += batch;
    if (cudaMemcpyAsync(h->flags, flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(&h->rr, st.sc + 2, sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
      return finish(DCP_ERR_CUDA);
    *last_step = h->flags[1];
    *last_residual = std::sqrt(h->rr);
    if (h->flags[0] == 1) return finish(DCP_OK);
    if (h->flags[0] == 2) break;
  }
  dcp_set_error("dcp_cg_solve: no convergence within the step limit");
  return finish(DCP_ERR_NO_CONVERGENCE);
}
