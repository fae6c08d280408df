// ILU(0) of a diagonal block and its triangular solves.
//
// Replaces LA::PreconditionILU = TrilinosWrappers::PreconditionILU (Ifpack ILU, deal.II defaults ilu_fill = 0,
// ilu_atol = 0, ilu_rtol = 1, overlap = 0) where the classic Schur-complement path uses it: the inner
// preconditioner of InverseMatrix<block(0,0)> (boussinesq_model.tpp:1265-1275, preconditioner.h:36-42) and the
// block(0,0) approximation inside ApproximateSchurComplement (approximate_schur_complement.hpp:118-141).
// With several ranks Ifpack factorises the rank-local square block (overlap 0): rows and columns >= `owned` are
// ignored here in the same way.
//
// The factorisation has the sparsity of A: IKJ variant, row i needs the finished rows k < i of its own pattern.
// Rows are grouped into dependency levels once per pattern (host, from the downloaded CSR pattern); one launch per
// level, one warp per row.  The same levels order the forward substitution; the backward substitution uses the
// levels of the upper triangle.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "dcp_internal.cuh"

struct dcp_ilu {
  dcp_model* model = nullptr;
  const DevCsr* A = nullptr;
  int64_t n = 0;                  // rows/columns of the factorised (rank-local) square block
  double* lu = nullptr;           // factors on the pattern of A: strict lower part = L (unit diagonal), rest = U
  int64_t* diag = nullptr;        // position of the diagonal entry of each row
  int32_t *rows_l = nullptr, *rows_u = nullptr;  // rows sorted by level (lower / upper dependencies)
  std::vector<int64_t> lvl_l, lvl_u;             // level start offsets into rows_l / rows_u
  double* tmp = nullptr;          // intermediate vector of the two substitutions
  // the level launches of one application, captured once per (source, destination) pair: a Krylov loop applies the
  // preconditioner to the same two vectors every iteration, so hundreds of launches become one graph launch
  cudaGraphExec_t solve_graph = nullptr;
  const double* graph_src = nullptr;
  double* graph_dst = nullptr;
  const double* last_src = nullptr;   // operands of the previous application: a pair is captured when it comes twice in a row
  double* last_dst = nullptr;
  int64_t graph_nodes = 0;
};

namespace {

__device__ __forceinline__ long long find_col(const int* __restrict__ col, long long lo, long long hi, int target) {
  const long long end = hi;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (col[mid] < target) lo = mid + 1; else hi = mid;
  }
  return (lo < end && col[lo] == target) ? lo : -1;
}

// one warp per row of the level
__global__ void __launch_bounds__(128) ilu_factor_level(const int* __restrict__ rows, int n_rows, long long n, const long long* __restrict__ rp,
                                                        const int* __restrict__ col, const long long* __restrict__ diag,
                                                        double* __restrict__ lu) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_rows) return;
  const int i = rows[w];
  const long long r0 = rp[i], r1 = rp[i + 1];
  for (long long p = r0; p < r1; ++p) {
    const int k = col[p];
    if (k >= i) break;
    const long long dk = diag[k];
    const double lik = lu[p] / lu[dk];
    __syncwarp();
    if (lane == 0) lu[p] = lik;
    // a_ij -= l_ik u_kj for the entries j > k of row k that exist in row i
    for (long long q = dk + 1 + lane; q < rp[k + 1]; q += 32) {
      const int j = col[q];
      if (j >= n) break;
      const long long t = find_col(col, p + 1, r1, j);
      if (t >= 0) lu[t] -= lik * lu[q];
    }
    __syncwarp();
  }
}

// forward: y_i = x_i - sum_{k<i} l_ik y_k ; backward: z_i = (y_i - sum_{j>i} u_ij z_j) / u_ii
template <bool UPPER>
__global__ void __launch_bounds__(128) ilu_solve_level(const int* __restrict__ rows, int n_rows, long long n, const long long* __restrict__ rp,
                                                       const int* __restrict__ col, const long long* __restrict__ diag,
                                                       const double* __restrict__ lu, const double* __restrict__ rhs, double* __restrict__ x) {
  constexpr int LANES = 8;
  const int lane = threadIdx.x % LANES;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const bool active = w < n_rows;
  const int i = active ? rows[w] : 0;
  double s = 0.0;
  if (active) {
    const long long d = diag[i];
    const long long a = UPPER ? d + 1 : rp[i], b = UPPER ? rp[i + 1] : d;
    for (long long p = a + lane; p < b; p += LANES) {
      const int c = col[p];
      if (c < n) s += lu[p] * x[c];
    }
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (active && lane == 0) x[i] = UPPER ? (rhs[i] - s) / lu[diag[i]] : rhs[i] - s;
}

}  // namespace

int dcp_ilu_destroy(dcp_ilu* p) {
  if (!p) return DCP_OK;
  cudaSetDevice(p->model->ctx->device);
  cudaStreamSynchronize(p->model->ctx->stream);
  auto& live = p->model->ilus;
  live.erase(std::remove(live.begin(), live.end(), p), live.end());
  cudaFree(p->lu);
  cudaFree(p->diag);
  cudaFree(p->rows_l);
  cudaFree(p->rows_u);
  cudaFree(p->tmp);
  if (p->solve_graph) cudaGraphExecDestroy(p->solve_graph);
  delete p;
  return DCP_OK;
}

int dcp_ilu_refactor(dcp_ilu* p) {
  if (!p) return DCP_ERR_ARG;
  dcp_ctx* ctx = p->model->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  const DevCsr& A = *p->A;
  DCP_CUDA(cudaMemcpyAsync(p->lu, A.val, sizeof(double) * (size_t)A.nnz, cudaMemcpyDeviceToDevice, ctx->stream));
  for (size_t l = 0; l + 1 < p->lvl_l.size(); ++l) {
    const int nr = (int)(p->lvl_l[l + 1] - p->lvl_l[l]);
    if (l == 0 || nr == 0) continue;  // rows of the first level have no lower entries
    ilu_factor_level<<<(nr * 32 + 127) / 128, 128, 0, ctx->stream>>>(p->rows_l + p->lvl_l[l], nr, p->n, (const long long*)A.rowptr, A.col,
                                                                     (const long long*)p->diag, p->lu);
    ++ctx->launches;
  }
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

int dcp_ilu_create(dcp_model* m, int which, int bi, dcp_ilu** out) {
  if (!m || !out) return DCP_ERR_ARG;
  *out = nullptr;
  BlockMat* M = dcp_select_matrix(m, which);
  if (!M || bi < 0 || bi >= M->nb) return DCP_ERR_ARG;
  const DevCsr& A = M->blk[bi][bi];
  if (A.nnz == 0 || A.n_rows != A.n_cols) {
    dcp_set_error("dcp_ilu_create: the block is empty or not square");
    return DCP_ERR_ARG;
  }
  dcp_ctx* ctx = m->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = M->owned[bi] >= 0 ? M->owned[bi] : A.n_rows;
  // ---- symbolic analysis on the host, from the pattern
  std::vector<int64_t> rp((size_t)A.n_rows + 1);
  std::vector<int32_t> col((size_t)A.nnz);
  DCP_CUDA(cudaMemcpyAsync(rp.data(), A.rowptr, sizeof(int64_t) * rp.size(), cudaMemcpyDeviceToHost, ctx->stream));
  DCP_CUDA(cudaMemcpyAsync(col.data(), A.col, sizeof(int32_t) * col.size(), cudaMemcpyDeviceToHost, ctx->stream));
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<int64_t> diag((size_t)n);
  std::vector<int32_t> level_l((size_t)n, 0), level_u((size_t)n, 0);
  int32_t max_l = 0, max_u = 0;
  for (int64_t i = 0; i < n; ++i) {
    int32_t lv = 0;
    int64_t d = -1;
    for (int64_t p = rp[i]; p < rp[i + 1]; ++p) {
      const int32_t c = col[p];
      if (c < i) lv = std::max(lv, level_l[c] + 1);
      if (c == i) d = p;
    }
    if (d < 0) {
      dcp_set_error("dcp_ilu_create: a row has no diagonal entry in the pattern");
      return DCP_ERR_PATTERN;
    }
    diag[i] = d;
    level_l[i] = lv;
    max_l = std::max(max_l, lv);
  }
  for (int64_t i = n - 1; i >= 0; --i) {
    int32_t lv = 0;
    for (int64_t p = diag[i] + 1; p < rp[i + 1]; ++p)
      if (col[p] < n) lv = std::max(lv, level_u[col[p]] + 1);
    level_u[i] = lv;
    max_u = std::max(max_u, lv);
  }
  auto bucket = [&](const std::vector<int32_t>& level, int32_t max_level, std::vector<int64_t>& start, std::vector<int32_t>& rows) {
    start.assign((size_t)max_level + 2, 0);
    for (int64_t i = 0; i < n; ++i) ++start[(size_t)level[i] + 1];
    for (size_t l = 0; l + 1 < start.size(); ++l) start[l + 1] += start[l];
    rows.resize((size_t)n);
    std::vector<int64_t> cur(start.begin(), start.end() - 1);
    for (int64_t i = 0; i < n; ++i) rows[(size_t)cur[level[i]]++] = (int32_t)i;
  };
  dcp_ilu* p = new dcp_ilu;
  p->model = m;
  p->A = &A;
  p->n = n;
  std::vector<int32_t> rows_l, rows_u;
  bucket(level_l, max_l, p->lvl_l, rows_l);
  bucket(level_u, max_u, p->lvl_u, rows_u);
  int rc = dcp_upload(ctx, &p->diag, diag.data(), n);
  if (rc == DCP_OK) rc = dcp_upload(ctx, &p->rows_l, rows_l.data(), n);
  if (rc == DCP_OK) rc = dcp_upload(ctx, &p->rows_u, rows_u.data(), n);
  if (rc == DCP_OK && cudaMalloc((void**)&p->lu, sizeof(double) * (size_t)A.nnz) != cudaSuccess) rc = DCP_ERR_CUDA;
  if (rc == DCP_OK && cudaMalloc((void**)&p->tmp, sizeof(double) * (size_t)n) != cudaSuccess) rc = DCP_ERR_CUDA;
  cudaStreamSynchronize(ctx->stream);
  if (rc == DCP_OK) rc = dcp_ilu_refactor(p);
  if (rc != DCP_OK) {
    dcp_ilu_destroy(p);
    return rc;
  }
  m->ilus.push_back(p);
  *out = p;
  return DCP_OK;
}

int dcp_ilu_levels(const dcp_ilu* p, int64_t* n_lower, int64_t* n_upper) {
  if (!p) return DCP_ERR_ARG;
  if (n_lower) *n_lower = (int64_t)p->lvl_l.size() - 1;
  if (n_upper) *n_upper = (int64_t)p->lvl_u.size() - 1;
  return DCP_OK;
}

int dcp_ilu_vmult(dcp_ilu* p, double* dst, const double* src, int mem) {
  if (!p || !dst || !src) return DCP_ERR_ARG;
  dcp_ctx* ctx = p->model->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  const DevCsr& A = *p->A;
  const double* dx;
  double* dy;
  DCP_TRY(dcp_stage_in(ctx, 0, src, p->n, mem, &dx));
  DCP_TRY(dcp_stage_out_alloc(ctx, 1, dst, p->n, mem, &dy));
  constexpr int LANES = 8;
  auto enqueue = [&](cudaStream_t st) {
    int64_t nodes = 0;
    for (size_t l = 0; l + 1 < p->lvl_l.size(); ++l) {
      const int nr = (int)(p->lvl_l[l + 1] - p->lvl_l[l]);
      if (nr == 0) continue;
      ilu_solve_level<false><<<(nr * LANES + 127) / 128, 128, 0, st>>>(p->rows_l + p->lvl_l[l], nr, p->n, (const long long*)A.rowptr, A.col,
                                                                      (const long long*)p->diag, p->lu, dx, p->tmp);
      ++nodes;
    }
    for (size_t l = 0; l + 1 < p->lvl_u.size(); ++l) {
      const int nr = (int)(p->lvl_u[l + 1] - p->lvl_u[l]);
      if (nr == 0) continue;
      ilu_solve_level<true><<<(nr * LANES + 127) / 128, 128, 0, st>>>(p->rows_u + p->lvl_u[l], nr, p->n, (const long long*)A.rowptr, A.col,
                                                                     (const long long*)p->diag, p->lu, p->tmp, dy);
      ++nodes;
    }
    return nodes;
  };
  static const bool use_graph = std::getenv("DCP_NO_ILU_GRAPH") == nullptr;
  if (use_graph && mem == DCP_DEVICE) {
    const bool have = p->solve_graph && p->graph_src == dx && p->graph_dst == dy;
    const bool repeated = p->last_src == dx && p->last_dst == dy;
    p->last_src = dx;
    p->last_dst = dy;
    if (!have && repeated) {
      if (p->solve_graph) cudaGraphExecDestroy(p->solve_graph);
      p->solve_graph = nullptr;
      cudaGraph_t g = nullptr;
      if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        p->graph_nodes = enqueue(ctx->stream);
        if (cudaStreamEndCapture(ctx->stream, &g) == cudaSuccess && g &&
            cudaGraphInstantiate(&p->solve_graph, g, nullptr, nullptr, 0) == cudaSuccess) {
          p->graph_src = dx;
          p->graph_dst = dy;
        } else
          p->solve_graph = nullptr;
        if (g) cudaGraphDestroy(g);
      }
      cudaGetLastError();
    }
    if (p->solve_graph && p->graph_src == dx && p->graph_dst == dy) {
      DCP_CUDA(cudaGraphLaunch(p->solve_graph, ctx->stream));
      ctx->launches += p->graph_nodes;
      return DCP_OK;
    }
  }
  ctx->launches += enqueue(ctx->stream);
  DCP_CUDA(cudaGetLastError());
  return dcp_stage_out_finish(ctx, 1, dst, p->n, mem);
}
