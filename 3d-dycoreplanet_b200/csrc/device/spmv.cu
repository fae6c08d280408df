// FP64 CSR SpMV and the small streaming kernels around it (Jacobi apply, CSR value axpy).
//
// Replaces Epetra_CrsMatrix::Multiply behind LA::SparseMatrix::vmult / vmult_add (call sites listed in
// include/dcp.h).  HBM-bound: algorithmic bytes per product = nnz*(8+4) + n_rows*(8+8) + n_cols*8
// (values + int32 columns streamed once, y written once, int64 row pointers, x read once; the x gather
// is served from the 126 MB L2).  Values and columns are read with ld.global.nc.L1::no_allocate (they
// have no reuse); x goes through the read-only path so that gathers hit L1/L2.
#include "dcp_internal.cuh"

namespace {

__device__ __forceinline__ double ld_stream_f64(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream_s32(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// LANES lanes cooperate on one row; 4-way unrolled so each lane keeps 8 independent loads in flight.
// Row selection for the overlap of the ghost exchange with the product (multi-GPU): `skip` != nullptr leaves out the
// rows flagged there (rows that read ghost columns); `list` != nullptr processes exactly the listed rows.
template <int LANES, bool ADD>
__global__ void __launch_bounds__(256) spmv_csr_kernel(long long n_rows, const long long* __restrict__ rowptr,
                                                       const int* __restrict__ col, const double* __restrict__ val,
                                                       const double* __restrict__ x, double* __restrict__ y,
                                                       const unsigned char* __restrict__ skip, const int* __restrict__ list) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long row = gtid / LANES;
  const int lane = threadIdx.x % LANES;
  double sum = 0.0;
  bool active = row < n_rows;
  if (active && list) row = list[row];
  if (active && skip && skip[row]) active = false;
  if (active) {
    const long long p0 = rowptr[row], p1 = rowptr[row + 1];
    // 4 x LANES entries per trip, every load predicated: the whole row (mean 200 nnz) is in flight after <= 2 trips
    for (long long p = p0 + lane; p < p1; p += 4 * LANES) {
      const bool k1 = p + LANES < p1, k2 = p + 2 * LANES < p1, k3 = p + 3 * LANES < p1;
      const int c0 = ld_stream_s32(col + p);
      const int c1 = k1 ? ld_stream_s32(col + p + LANES) : 0;
      const int c2 = k2 ? ld_stream_s32(col + p + 2 * LANES) : 0;
      const int c3 = k3 ? ld_stream_s32(col + p + 3 * LANES) : 0;
      const double v0 = ld_stream_f64(val + p);
      const double v1 = k1 ? ld_stream_f64(val + p + LANES) : 0.0;
      const double v2 = k2 ? ld_stream_f64(val + p + 2 * LANES) : 0.0;
      const double v3 = k3 ? ld_stream_f64(val + p + 3 * LANES) : 0.0;
      sum += v0 * __ldg(x + c0);
      sum += v1 * __ldg(x + c1);
      sum += v2 * __ldg(x + c2);
      sum += v3 * __ldg(x + c3);
    }
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (active && lane == 0) y[row] = ADD ? y[row] + sum : sum;
}

// rows of the owned range whose last (largest) column is a ghost column
__global__ void flag_ghost_rows_kernel(long long n_rows, const long long* __restrict__ rowptr, const int* __restrict__ col,
                                       int owned_cols, unsigned char* __restrict__ flag) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const long long a = rowptr[r], b = rowptr[r + 1];
  if (b > a && col[b - 1] >= owned_cols) flag[r] = 1;
}
__global__ void list_flagged_rows_kernel(long long n_rows, const unsigned char* __restrict__ flag, int* __restrict__ list,
                                         int* __restrict__ count) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rows && flag[r]) list[atomicAdd(count, 1)] = (int)r;
}

// with a cached offset of the diagonal inside each row (found once per pattern) the refresh is one gather
__global__ void diag_inv_from_offsets_kernel(long long n_rows, const long long* __restrict__ rowptr, const int* __restrict__ off,
                                             const double* __restrict__ val, double* __restrict__ dinv) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int o = off[r];
  const double d = o >= 0 ? val[rowptr[r] + o] : 0.0;
  dinv[r] = d != 0.0 ? 1.0 / d : 0.0;
}

__global__ void extract_diag_inv_kernel(long long n_rows, const long long* __restrict__ rowptr,
                                        const int* __restrict__ col, const double* __restrict__ val,
                                        double* __restrict__ dinv, int* __restrict__ off_out) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  long long lo = rowptr[r], hi = rowptr[r + 1];
  const long long end = hi;
  const int target = (int)r;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if (col[mid] < target) lo = mid + 1; else hi = mid;
  }
  const bool found = lo < end && col[lo] == target;
  double d = found ? val[lo] : 0.0;
  dinv[r] = d != 0.0 ? 1.0 / d : 0.0;
  if (off_out) off_out[r] = found ? (int)(lo - rowptr[r]) : -1;
}

__global__ void jacobi_kernel(long long n, const double* __restrict__ dinv, const double* __restrict__ x,
                              double* __restrict__ y) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] = dinv[i] * x[i];
}

__global__ void axpby_values_kernel(long long n, const double* __restrict__ a, const double* __restrict__ b, double fb,
                                    double* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = a[i] + fb * b[i];
}

__global__ void fill_kernel(double* __restrict__ p, long long n, double v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

template <int LANES>
int launch_spmv_t(dcp_ctx* ctx, const DevCsr& A, const double* x, double* y, bool add, long long n_rows,
                  const unsigned char* skip, const int* list) {
  const int threads = 256;
  const long long rows_per_block = threads / LANES;
  const long long blocks = (n_rows + rows_per_block - 1) / rows_per_block;
  if (blocks == 0) return DCP_OK;
  if (add)
    spmv_csr_kernel<LANES, true><<<(unsigned)blocks, threads, 0, ctx->stream>>>(
        n_rows, (const long long*)A.rowptr, A.col, A.val, x, y, skip, list);
  else
    spmv_csr_kernel<LANES, false><<<(unsigned)blocks, threads, 0, ctx->stream>>>(
        n_rows, (const long long*)A.rowptr, A.col, A.val, x, y, skip, list);
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

inline unsigned grid_for(dcp_ctx* ctx, long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  long long cap = (long long)ctx->sm_count * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int dcp_launch_spmv(dcp_ctx* ctx, const DevCsr& A, const double* x, double* y, bool add, int64_t row_limit,
                    const unsigned char* skip, const int* list, int64_t n_list) {
  long long n_rows = row_limit >= 0 && row_limit < A.n_rows ? row_limit : A.n_rows;
  if (list) n_rows = n_list;
  if (n_rows == 0) return DCP_OK;
  if (A.rowptr == nullptr || A.nnz == 0) {
    if (!add && !skip && !list) return dcp_launch_fill(ctx, y, n_rows, 0.0);
    return DCP_OK;
  }
  switch (A.lanes) {
    case 4: return launch_spmv_t<4>(ctx, A, x, y, add, n_rows, skip, list);
    case 8: return launch_spmv_t<8>(ctx, A, x, y, add, n_rows, skip, list);
    case 16: return launch_spmv_t<16>(ctx, A, x, y, add, n_rows, skip, list);
    default: return launch_spmv_t<32>(ctx, A, x, y, add, n_rows, skip, list);
  }
}

// flags + list of the rows of block row r (all its blocks) that read a ghost column
int dcp_build_ghost_rows(dcp_ctx* ctx, BlockMat& M, int r, const int64_t* owned_cols) {
  const int64_t n = M.owned[r] >= 0 ? M.owned[r] : M.start[r + 1] - M.start[r];
  cudaFree(M.ghost_flag[r]);
  cudaFree(M.ghost_list[r]);
  M.ghost_flag[r] = nullptr;
  M.ghost_list[r] = nullptr;
  M.n_ghost_rows[r] = 0;
  if (n == 0) return DCP_OK;
  DCP_CUDA(cudaMalloc((void**)&M.ghost_flag[r], (size_t)n));
  DCP_CUDA(cudaMemsetAsync(M.ghost_flag[r], 0, (size_t)n, ctx->stream));
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  for (int c = 0; c < M.nb; ++c) {
    const DevCsr& A = M.blk[r][c];
    if (A.nnz == 0) continue;
    flag_ghost_rows_kernel<<<blocks, threads, 0, ctx->stream>>>(n, (const long long*)A.rowptr, A.col, (int)owned_cols[c], M.ghost_flag[r]);
    ctx->launches++;
  }
  int* d_count = nullptr;
  DCP_CUDA(cudaMalloc((void**)&d_count, sizeof(int)));
  DCP_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), ctx->stream));
  DCP_CUDA(cudaMalloc((void**)&M.ghost_list[r], sizeof(int) * (size_t)n));
  list_flagged_rows_kernel<<<blocks, threads, 0, ctx->stream>>>(n, M.ghost_flag[r], M.ghost_list[r], d_count);
  ctx->launches++;
  int h_count = 0;
  DCP_CUDA(cudaMemcpyAsync(&h_count, d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_count);
  M.n_ghost_rows[r] = h_count;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

namespace {
template <bool SCATTER>
__global__ void gather_scatter_kernel(long long n, const int* __restrict__ idx, const double* __restrict__ src,
                                      double* __restrict__ dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    if (SCATTER) dst[idx[i]] = src[i]; else dst[i] = src[idx[i]];
  }
}
}  // namespace

int dcp_launch_gather(dcp_ctx* ctx, int64_t n, const int32_t* idx, const double* src, double* dst, bool scatter) {
  if (n == 0) return DCP_OK;
  if (scatter)
    gather_scatter_kernel<true><<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(n, idx, src, dst);
  else
    gather_scatter_kernel<false><<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(n, idx, src, dst);
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

int dcp_launch_extract_diag_inv(dcp_ctx* ctx, const DevCsr& A, double* diag_inv, int32_t** diag_off) {
  if (A.n_rows == 0) return DCP_OK;
  const int threads = 256;
  unsigned blocks = (unsigned)((A.n_rows + threads - 1) / threads);
  if (diag_off && *diag_off) {
    diag_inv_from_offsets_kernel<<<blocks, threads, 0, ctx->stream>>>(A.n_rows, (const long long*)A.rowptr, *diag_off, A.val, diag_inv);
  } else {
    if (diag_off) DCP_CUDA(cudaMalloc((void**)diag_off, sizeof(int32_t) * (size_t)A.n_rows));
    extract_diag_inv_kernel<<<blocks, threads, 0, ctx->stream>>>(A.n_rows, (const long long*)A.rowptr, A.col, A.val, diag_inv,
                                                                 diag_off ? *diag_off : nullptr);
  }
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

int dcp_launch_jacobi(dcp_ctx* ctx, int64_t n, const double* diag_inv, const double* x, double* y) {
  if (n == 0) return DCP_OK;
  jacobi_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(n, diag_inv, x, y);
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

int dcp_launch_axpby_values(dcp_ctx* ctx, int64_t n, const double* a, const double* b, double fb, double* out) {
  if (n == 0) return DCP_OK;
  axpby_values_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(n, a, b, fb, out);
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

int dcp_launch_fill(dcp_ctx* ctx, double* p, int64_t n, double v) {
  if (n == 0) return DCP_OK;
  fill_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(p, n, v);
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}
