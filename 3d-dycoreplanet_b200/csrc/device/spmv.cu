// FP64 CSR SpMV and the small streaming kernels around it (Jacobi apply, CSR value axpy).
//
// Replaces Epetra_CrsMatrix::Multiply behind LA::SparseMatrix::vmult / vmult_add (call sites listed in
// include/dcp.h).  HBM-bound: algorithmic bytes per product = nnz*(8+4) + n_rows*(8+8) + n_cols*8
// (values + int32 columns streamed once, y written once, int64 row pointers, x read once; the x gather
// is served from the 126 MB L2).  Values and columns are read with ld.global.nc.L1::no_allocate (they
// have no reuse); x goes through the read-only path so that gathers hit L1/L2.
#include <cstdlib>

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

// Node-blocked matrices (the velocity rows of nse_matrix and of its preconditioner: the three component rows of a Q2
// node are consecutive and couple to the same columns, boussinesq_model.tpp:194-204 + the all-to-all velocity coupling
// table :219-232): one group of lanes takes the three rows together, reads the column indices and the x entries once
// and three value streams -- 9.3 instead of 12 bytes per nonzero.  Same per-row summation order as spmv_csr_kernel.
// Used when check_row_triples_kernel found every group of three rows with one pattern.
template <int LANES, bool ADD>
__global__ void __launch_bounds__(256) spmv_csr3_kernel(long long n_groups, const long long* __restrict__ rowptr,
                                                        const int* __restrict__ col, const double* __restrict__ val,
                                                        const double* __restrict__ x, double* __restrict__ y,
                                                        const unsigned char* __restrict__ skip, const unsigned char* __restrict__ same) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long grp = gtid / LANES;
  const int lane = threadIdx.x % LANES;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  bool active = grp < n_groups;
  // the lanes of this group: groups of one warp take different paths
  const unsigned gmask = LANES == 32 ? 0xffffffffu : ((1u << (LANES & 31)) - 1u) << (LANES * ((threadIdx.x & 31) / LANES));
  const bool shared_pattern = active && same[grp];
  if (active && !shared_pattern) {
    // rows with their own patterns (nodes next to constrained dofs): row by row, skip flags per row
    bool any = false;
#pragma unroll 1
    for (int r = 0; r < 3; ++r) {
      const long long row = 3 * grp + r;
      double sum = 0.0;
      const bool on = !(skip && skip[row]);
      if (on) {
        const long long p0 = rowptr[row], p1 = rowptr[row + 1];
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
      for (int o = LANES / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(gmask, sum, o, LANES);
      if (on && lane == 0) y[row] = ADD ? y[row] + sum : sum;
      any |= on;
    }
    (void)any;
    return;
  }
  if (active && skip && skip[3 * grp]) active = false;
  if (active) {
    const long long p0 = rowptr[3 * grp], len = rowptr[3 * grp + 1] - p0;
    const int* c = col + p0;
    const double *va = val + p0, *vb = va + len, *vc = vb + len;
    for (long long p = lane; p < len; p += 4 * LANES) {
      const bool k1 = p + LANES < len, k2 = p + 2 * LANES < len, k3 = p + 3 * LANES < len;
      const int c0 = ld_stream_s32(c + p);
      const int c1 = k1 ? ld_stream_s32(c + p + LANES) : 0;
      const int c2 = k2 ? ld_stream_s32(c + p + 2 * LANES) : 0;
      const int c3 = k3 ? ld_stream_s32(c + p + 3 * LANES) : 0;
      const double a0 = ld_stream_f64(va + p), b0 = ld_stream_f64(vb + p), d0 = ld_stream_f64(vc + p);
      const double a1 = k1 ? ld_stream_f64(va + p + LANES) : 0.0, b1 = k1 ? ld_stream_f64(vb + p + LANES) : 0.0,
                   d1 = k1 ? ld_stream_f64(vc + p + LANES) : 0.0;
      const double a2 = k2 ? ld_stream_f64(va + p + 2 * LANES) : 0.0, b2 = k2 ? ld_stream_f64(vb + p + 2 * LANES) : 0.0,
                   d2 = k2 ? ld_stream_f64(vc + p + 2 * LANES) : 0.0;
      const double a3 = k3 ? ld_stream_f64(va + p + 3 * LANES) : 0.0, b3 = k3 ? ld_stream_f64(vb + p + 3 * LANES) : 0.0,
                   d3 = k3 ? ld_stream_f64(vc + p + 3 * LANES) : 0.0;
      const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
      s0 += a0 * x0; s0 += a1 * x1; s0 += a2 * x2; s0 += a3 * x3;
      s1 += b0 * x0; s1 += b1 * x1; s1 += b2 * x2; s1 += b3 * x3;
      s2 += d0 * x0; s2 += d1 * x1; s2 += d2 * x2; s2 += d3 * x3;
    }
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(gmask, s0, o, LANES);
    s1 += __shfl_xor_sync(gmask, s1, o, LANES);
    s2 += __shfl_xor_sync(gmask, s2, o, LANES);
  }
  if (active && lane < 3) {
    const double s = lane == 0 ? s0 : (lane == 1 ? s1 : s2);
    double* out = y + 3 * grp + lane;
    *out = ADD ? *out + s : s;
  }
}

// same[g] = 1: the three rows 3g .. 3g+2 share one pattern; *n_same counts them
__global__ void check_row_triples_kernel(long long n_groups, const long long* __restrict__ rowptr, const int* __restrict__ col,
                                         unsigned char* __restrict__ same, unsigned long long* __restrict__ n_same) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long grp = gtid / 8;
  const int lane = threadIdx.x % 8;
  if (grp >= n_groups) return;
  const long long p0 = rowptr[3 * grp], p1 = rowptr[3 * grp + 1], p2 = rowptr[3 * grp + 2], p3 = rowptr[3 * grp + 3];
  const long long len = p1 - p0;
  bool ok = (p2 - p1 == len) && (p3 - p2 == len);
  if (ok)
    for (long long p = lane; p < len && ok; p += 8) ok = col[p0 + p] == col[p1 + p] && col[p0 + p] == col[p2 + p];
  ok = __all_sync(0xffu << (8 * ((threadIdx.x & 31) / 8)), ok);
  if (lane == 0) {
    same[grp] = ok ? 1 : 0;
    if (ok) atomicAdd(n_same, 1ull);
  }
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

template <int LANES>
int launch_spmv3_t(dcp_ctx* ctx, const DevCsr& A, const double* x, double* y, bool add, long long n_groups, const unsigned char* skip) {
  const int threads = 256;
  const long long per_block = threads / LANES;
  const long long blocks = (n_groups + per_block - 1) / per_block;
  if (blocks == 0) return DCP_OK;
  if (add)
    spmv_csr3_kernel<LANES, true><<<(unsigned)blocks, threads, 0, ctx->stream>>>(n_groups, (const long long*)A.rowptr, A.col, A.val, x, y, skip,
                                                                                 A.triple_same);
  else
    spmv_csr3_kernel<LANES, false><<<(unsigned)blocks, threads, 0, ctx->stream>>>(n_groups, (const long long*)A.rowptr, A.col, A.val, x, y, skip,
                                                                                  A.triple_same);
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

// one-time test of a pattern: which groups of three consecutive rows share their columns?  The grouped kernel is used
// when most do (all velocity nodes away from constrained dofs).
int row_triples(dcp_ctx* ctx, const DevCsr& A) {
  if (A.triple >= 0) return A.triple;
  A.triple = 0;
  if (A.n_rows % 3 != 0 || A.n_rows == 0 || std::getenv("DCP_NO_SPMV3")) return 0;
  const long long n_groups = A.n_rows / 3;
  unsigned long long* d_n = nullptr;
  unsigned char* d_same = nullptr;
  if (cudaMalloc((void**)&d_n, sizeof(unsigned long long)) != cudaSuccess || cudaMalloc((void**)&d_same, (size_t)n_groups) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(d_n);
    return 0;
  }
  cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), ctx->stream);
  check_row_triples_kernel<<<(unsigned)((n_groups * 8 + 255) / 256), 256, 0, ctx->stream>>>(n_groups, (const long long*)A.rowptr, A.col, d_same, d_n);
  unsigned long long h_n = 0;
  if (cudaMemcpyAsync(&h_n, d_n, sizeof(h_n), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    h_n = 0;
  cudaFree(d_n);
  if (2 * h_n > (unsigned long long)n_groups) {
    A.triple = 1;
    A.triple_same = d_same;   // lives as long as the pattern (freed with the model's matrices)
  } else
    cudaFree(d_same);
  return A.triple;
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
  // long rows only: with a dozen entries per row (block(0,1)) the grouped kernel was measured slower (0.246 against 0.215 ms)
  if (!list && A.lanes >= 16 && n_rows % 3 == 0 && row_triples(ctx, A)) {
    static const int lanes3 = std::getenv("DCP_SPMV3_LANES") ? std::atoi(std::getenv("DCP_SPMV3_LANES")) : 0;
    switch (lanes3 > 0 ? lanes3 : A.lanes) {
      case 8: return launch_spmv3_t<8>(ctx, A, x, y, add, n_rows / 3, skip);
      case 16: return launch_spmv3_t<16>(ctx, A, x, y, add, n_rows / 3, skip);
      default: return launch_spmv3_t<32>(ctx, A, x, y, add, n_rows / 3, skip);
    }
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
