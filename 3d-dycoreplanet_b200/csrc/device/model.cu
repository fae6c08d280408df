// Context, model object and the C ABI entry points declared in include/dcp.h.
#include <algorithm>
#include <cstring>

#include "dcp_internal.cuh"

static thread_local std::string g_dcp_err;
void dcp_set_error(const std::string& s) { g_dcp_err = s; }

template <class T>
int dcp_upload(dcp_ctx* ctx, T** dst, const T* src, int64_t n) {
  *dst = nullptr;
  if (n <= 0) return DCP_OK;
  if (src == nullptr) {
    dcp_set_error("dcp_upload: null source for a non-empty array");
    return DCP_ERR_ARG;
  }
  DCP_CUDA(cudaMalloc((void**)dst, sizeof(T) * (size_t)n));
  DCP_CUDA(cudaMemcpyAsync(*dst, src, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  return DCP_OK;
}
template int dcp_upload<double>(dcp_ctx*, double**, const double*, int64_t);
template int dcp_upload<int32_t>(dcp_ctx*, int32_t**, const int32_t*, int64_t);
template int dcp_upload<int64_t>(dcp_ctx*, int64_t**, const int64_t*, int64_t);
template int dcp_upload<uint16_t>(dcp_ctx*, uint16_t**, const uint16_t*, int64_t);
template int dcp_upload<uint8_t>(dcp_ctx*, uint8_t**, const uint8_t*, int64_t);

int dcp_check_device_errors(dcp_ctx* ctx, const char* what) {
  DCP_CUDA(cudaMemcpyAsync(ctx->h_err, ctx->d_err, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->h_err[0] != 0) {
    int n = ctx->h_err[0];
    cudaMemsetAsync(ctx->d_err, 0, 4 * sizeof(int), ctx->stream);
    dcp_set_error(std::string(what) + ": " + std::to_string(n) + " scatter targets are not in the sparsity pattern");
    return DCP_ERR_PATTERN;
  }
  if (ctx->h_err[1] != 0) {
    cudaMemsetAsync(ctx->d_err, 0, 4 * sizeof(int), ctx->stream);
    dcp_set_error(std::string(what) + ": a device-side wait was abandoned; the matrices of that pass are invalid");
    return DCP_ERR_STATE;
  }
  return DCP_OK;
}

static int ensure_stage(dcp_ctx* ctx, int slot, int64_t n) {
  if (ctx->stage_cap[slot] >= n) return DCP_OK;
  if (ctx->stage[slot]) DCP_CUDA(cudaFree(ctx->stage[slot]));
  ctx->stage[slot] = nullptr;
  ctx->stage_cap[slot] = 0;
  DCP_CUDA(cudaMalloc((void**)&ctx->stage[slot], sizeof(double) * (size_t)n));
  ctx->stage_cap[slot] = n;
  return DCP_OK;
}

int dcp_stage_in(dcp_ctx* ctx, int slot, const double* src, int64_t n, int mem, const double** dev) {
  if (mem == DCP_DEVICE) {
    *dev = src;
    return DCP_OK;
  }
  DCP_TRY(ensure_stage(ctx, slot, n));
  DCP_CUDA(cudaMemcpyAsync(ctx->stage[slot], src, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  *dev = ctx->stage[slot];
  return DCP_OK;
}
int dcp_stage_out_alloc(dcp_ctx* ctx, int slot, double* dst, int64_t n, int mem, double** dev) {
  if (mem == DCP_DEVICE) {
    *dev = dst;
    return DCP_OK;
  }
  DCP_TRY(ensure_stage(ctx, slot, n));
  *dev = ctx->stage[slot];
  return DCP_OK;
}
int dcp_stage_out_finish(dcp_ctx* ctx, int slot, double* dst, int64_t n, int mem) {
  if (mem == DCP_DEVICE) return DCP_OK;
  DCP_CUDA(cudaMemcpyAsync(dst, ctx->stage[slot], sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  return DCP_OK;
}

// -------------------------------------------------------------------------------------------------
static int upload_cs(dcp_ctx* ctx, const dcp_constraints_desc& d, DevCs& out) {
  out.n_dofs = d.n_dofs;
  out.n_lines = d.n_lines;
  std::vector<int32_t> lod((size_t)d.n_dofs, -1);
  for (int64_t l = 0; l < d.n_lines; ++l) {
    if (d.line_dof[l] < 0 || d.line_dof[l] >= d.n_dofs) {
      dcp_set_error("constraints: line_dof out of range");
      return DCP_ERR_ARG;
    }
    lod[d.line_dof[l]] = (int32_t)l;
  }
  DCP_TRY(dcp_upload(ctx, &out.line_of_dof, lod.data(), d.n_dofs));
  // the async copy above reads from `lod`: finish it before the vector dies
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  static const int32_t zero_ptr[1] = {0};
  DCP_TRY(dcp_upload(ctx, &out.line_ptr, d.n_lines ? d.line_ptr : zero_ptr, d.n_lines + 1));
  int64_t ne = d.n_lines ? d.line_ptr[d.n_lines] : 0;
  // keep one valid element so that kernels never see a null pointer
  static const int32_t zi[1] = {0};
  static const double zd[1] = {0.0};
  DCP_TRY(dcp_upload(ctx, &out.entry_dof, ne ? d.entry_dof : zi, ne ? ne : 1));
  DCP_TRY(dcp_upload(ctx, &out.entry_w, ne ? d.entry_w : zd, ne ? ne : 1));
  DCP_TRY(dcp_upload(ctx, &out.inhom, d.n_lines ? d.inhom : zd, d.n_lines ? d.n_lines : 1));
  return DCP_OK;
}

static void free_cs(DevCs& c) {
  cudaFree(c.line_of_dof);
  cudaFree(c.line_ptr);
  cudaFree(c.entry_dof);
  cudaFree(c.entry_w);
  cudaFree(c.inhom);
  c = DevCs();
}

static int pick_lanes(int64_t n_rows, int64_t nnz) {
  if (n_rows == 0) return 32;
  double mean = (double)nnz / (double)n_rows;
  // measured on block(0,1) of the shell (14.5 entries per row): 16 lanes 2.2 TB/s, 8 lanes 3.5 TB/s, 4 lanes 4.4 TB/s
  if (mean <= 16) return 4;
  if (mean <= 32) return 8;
  if (mean <= 256) return 16;
  return 32;
}

static int upload_csr_pattern(dcp_ctx* ctx, const dcp_csr_desc& d, DevCsr& A, bool alloc_values) {
  A = DevCsr();
  A.n_rows = d.n_rows;
  A.n_cols = d.n_cols;
  if (d.rowptr == nullptr || d.n_rows == 0) return DCP_OK;
  A.nnz = d.rowptr[d.n_rows];
  if (A.nnz == 0) return DCP_OK;
  DCP_TRY(dcp_upload(ctx, &A.rowptr, d.rowptr, d.n_rows + 1));
  DCP_TRY(dcp_upload(ctx, &A.col, d.col, A.nnz));
  if (alloc_values) {
    DCP_CUDA(cudaMalloc((void**)&A.val, sizeof(double) * (size_t)A.nnz));
    DCP_CUDA(cudaMemsetAsync(A.val, 0, sizeof(double) * (size_t)A.nnz, ctx->stream));
  }
  A.lanes = pick_lanes(A.n_rows, A.nnz);
  return DCP_OK;
}

static int share_csr_pattern(dcp_ctx* ctx, const DevCsr& src, DevCsr& A) {
  A = src;
  A.owns_pattern = false;
  A.val = nullptr;
  if (A.nnz) {
    DCP_CUDA(cudaMalloc((void**)&A.val, sizeof(double) * (size_t)A.nnz));
    DCP_CUDA(cudaMemsetAsync(A.val, 0, sizeof(double) * (size_t)A.nnz, ctx->stream));
  }
  return DCP_OK;
}

static void free_csr(DevCsr& A) {
  cudaFree(A.triple_same);
  A.triple_same = nullptr;
  if (A.owns_pattern) {
    cudaFree(A.rowptr);
    cudaFree(A.col);
  }
  cudaFree(A.val);
  A = DevCsr();
}

static void free_blockmat(BlockMat& M) {
  for (int i = 0; i < DCP_MAXB; ++i) {
    for (int j = 0; j < DCP_MAXB; ++j) free_csr(M.blk[i][j]);
    cudaFree(M.diag_inv[i]);
    M.diag_inv[i] = nullptr;
    cudaFree(M.diag_off[i]);
    M.diag_off[i] = nullptr;
    if (M.owns_ghost_rows) {
      cudaFree(M.ghost_flag[i]);
      cudaFree(M.ghost_list[i]);
    }
    M.ghost_flag[i] = nullptr;
    M.ghost_list[i] = nullptr;
  }
}

static int zero_blockmat(dcp_ctx* ctx, BlockMat& M) {
  for (int i = 0; i < M.nb; ++i)
    for (int j = 0; j < M.nb; ++j)
      if (M.blk[i][j].val) DCP_CUDA(cudaMemsetAsync(M.blk[i][j].val, 0, sizeof(double) * (size_t)M.blk[i][j].nnz, ctx->stream));
  return DCP_OK;
}

static int refresh_jacobi(dcp_ctx* ctx, BlockMat& M) {
  for (int b = 0; b < M.nb; ++b) {
    DevCsr& A = M.blk[b][b];
    if (A.nnz == 0) continue;
    if (!M.diag_inv[b]) DCP_CUDA(cudaMalloc((void**)&M.diag_inv[b], sizeof(double) * (size_t)A.n_rows));
    DCP_TRY(dcp_launch_extract_diag_inv(ctx, A, M.diag_inv[b], &M.diag_off[b]));
  }
  return DCP_OK;
}

static BlockMat* select_matrix(dcp_model* m, int which) {
  switch (which) {
    case DCP_MAT_NSE: return &m->nse;
    case DCP_MAT_NSE_PRECOND: return &m->pre;
    case DCP_MAT_TEMP_MASS: return &m->tmass;
    case DCP_MAT_TEMP_STIFF: return &m->tstiff;
    case DCP_MAT_TEMP: return &m->tmat;
    default: return nullptr;
  }
}

BlockMat* dcp_select_matrix(dcp_model* m, int which) { return select_matrix(m, which); }

static int get_block(dcp_model* m, int which, int bi, int bj, DevCsr** out, BlockMat** bm = nullptr) {
  BlockMat* M = select_matrix(m, which);
  if (!M || bi < 0 || bj < 0 || bi >= M->nb || bj >= M->nb) {
    dcp_set_error("invalid matrix / block selector");
    return DCP_ERR_ARG;
  }
  *out = &M->blk[bi][bj];
  if (bm) *bm = M;
  return DCP_OK;
}

extern "C" {

const char* dcp_last_error(void) { return g_dcp_err.c_str(); }

int dcp_ctx_create(int device, dcp_ctx** out) {
  if (!out) return DCP_ERR_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    dcp_set_error(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                  "); this library has no CPU fallback");
    return DCP_ERR_CUDA;
  }
  if (device < 0 || device >= n) {
    dcp_set_error("device index out of range");
    return DCP_ERR_ARG;
  }
  DCP_CUDA(cudaSetDevice(device));
  dcp_ctx* ctx = new dcp_ctx;
  ctx->device = device;
  cudaDeviceProp prop;
  // a failure after this point releases what has been created so far (dcp_ctx_destroy copes with the unset members)
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess ||
      (e = cudaMalloc((void**)&ctx->d_err, 4 * sizeof(int))) != cudaSuccess ||
      (e = cudaMemset(ctx->d_err, 0, 4 * sizeof(int))) != cudaSuccess ||
      (e = cudaMallocHost((void**)&ctx->h_err, 4 * sizeof(int))) != cudaSuccess) {
    dcp_set_error(std::string("dcp_ctx_create: ") + cudaGetErrorString(e));
    cudaGetLastError();
    if (!ctx->stream) ctx->own_stream = false;
    dcp_ctx_destroy(ctx);
    return DCP_ERR_CUDA;
  }
  ctx->sm_count = prop.multiProcessorCount;
  *out = ctx;
  return DCP_OK;
}

int dcp_ctx_destroy(dcp_ctx* ctx) {
  if (!ctx) return DCP_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int s = 0; s < 3; ++s) cudaFree(ctx->stage[s]);
  cudaFree(ctx->d_err);
  cudaFreeHost(ctx->h_err);
  cudaFree(ctx->dot_scratch);
  cudaFreeHost(ctx->dot_host);
  cudaFree(ctx->mgs_scalars);
  cudaFreeHost(ctx->mgs_host);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaEventDestroy(ctx->copy_event);
  }
  delete ctx;
  return DCP_OK;
}

int dcp_ctx_set_stream(dcp_ctx* ctx, void* cuda_stream) {
  if (!ctx) return DCP_ERR_ARG;
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  if (cuda_stream == nullptr) {
    if (!ctx->own_stream) {
      DCP_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
      ctx->own_stream = true;
    }
    return DCP_OK;
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return DCP_OK;
}

int dcp_ctx_synchronize(dcp_ctx* ctx) {
  if (!ctx) return DCP_ERR_ARG;
  return dcp_check_device_errors(ctx, "dcp_ctx_synchronize");
}

int64_t dcp_ctx_launch_count(const dcp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int dcp_malloc(dcp_ctx* ctx, int64_t bytes, void** out) {
  if (!ctx || !out || bytes < 0) return DCP_ERR_ARG;
  DCP_CUDA(cudaSetDevice(ctx->device));
  DCP_CUDA(cudaMalloc(out, (size_t)(bytes > 0 ? bytes : 1)));
  return DCP_OK;
}
int dcp_free(dcp_ctx* ctx, void* p) {
  if (!ctx) return DCP_ERR_ARG;
  DCP_CUDA(cudaFree(p));
  return DCP_OK;
}
int dcp_memcpy_h2d(dcp_ctx* ctx, void* dst_dev, const void* src_host, int64_t bytes) {
  if (!ctx) return DCP_ERR_ARG;
  DCP_CUDA(cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  return DCP_OK;
}
int dcp_memcpy_d2h(dcp_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes) {
  if (!ctx) return DCP_ERR_ARG;
  DCP_CUDA(cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  return DCP_OK;
}

// ---- asynchronous host <-> device copies on the context's copy stream ---------------------------------------------
static int copy_stream(dcp_ctx* ctx) {
  if (ctx->copy_stream) return DCP_OK;
  DCP_CUDA(cudaSetDevice(ctx->device));
  DCP_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  DCP_CUDA(cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming));
  return DCP_OK;
}
int dcp_memcpy_h2d_async(dcp_ctx* ctx, void* dst_dev, const void* src_host, int64_t bytes) {
  if (!ctx || bytes < 0) return DCP_ERR_ARG;
  DCP_TRY(copy_stream(ctx));
  DCP_CUDA(cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
  return DCP_OK;
}
int dcp_memcpy_d2h_async(dcp_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes) {
  if (!ctx || bytes < 0) return DCP_ERR_ARG;
  DCP_TRY(copy_stream(ctx));
  DCP_CUDA(cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
  return DCP_OK;
}
int dcp_copy_fence(dcp_ctx* ctx, int direction) {
  if (!ctx || (direction != DCP_COMPUTE_WAITS_FOR_COPIES && direction != DCP_COPIES_WAIT_FOR_COMPUTE)) return DCP_ERR_ARG;
  DCP_TRY(copy_stream(ctx));
  cudaStream_t from = direction == DCP_COMPUTE_WAITS_FOR_COPIES ? ctx->copy_stream : ctx->stream;
  cudaStream_t to = direction == DCP_COMPUTE_WAITS_FOR_COPIES ? ctx->stream : ctx->copy_stream;
  DCP_CUDA(cudaEventRecord(ctx->copy_event, from));
  DCP_CUDA(cudaStreamWaitEvent(to, ctx->copy_event, 0));
  return DCP_OK;
}
int dcp_copy_synchronize(dcp_ctx* ctx) {
  if (!ctx) return DCP_ERR_ARG;
  if (ctx->copy_stream) DCP_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  return DCP_OK;
}

int dcp_model_destroy(dcp_model* m) {
  if (!m) return DCP_OK;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  while (!m->ilus.empty()) dcp_ilu_destroy(m->ilus.back());  // handles that outlived their model's owner
  cudaFree(m->nse_l2g);
  cudaFree(m->temp_l2g);
  cudaFree(m->temp_pos);
  cudaFree(m->temp_q2_tab);
  cudaFree(m->temp_fast_cells);
  cudaFree(m->temp_general_cells);
  cudaFree(m->temp_bc_flag);
  cudaFree(m->temp_bc_cells);
  cudaFree(m->vel_dof);
  cudaFree(m->cell_vertices);
  cudaFree(m->nse_local_field);
  cudaFree(m->nse_local_base);
  cudaFree(m->nse_constrained_cells);
  free_cs(m->nse_cs);
  free_cs(m->temp_cs);
  cudaFree(m->phi_u_qn);
  cudaFree(m->dphi_u_qn);
  cudaFree(m->phi_p_qn);
  cudaFree(m->phi_t_qn);
  cudaFree(m->phi_u_qt);
  cudaFree(m->phi_t_qt);
  cudaFree(m->dphi_t_qt);
  cudaFree(m->phi_u_qt_T);
  cudaFree(m->phi_t_qt_T);
  cudaFree(m->dphi_t_qt_T);
  cudaFree(m->geom_qn);
  if (!m->geom_shared) cudaFree(m->geom_qt);
  cudaFree(m->geom_qp);
  cudaFree(m->nse_sign);
  cudaFree(m->feec_w_qn);
  cudaFree(m->feec_c_qn);
  cudaFree(m->feec_u_qn);
  cudaFree(m->feec_w_qn_t);
  cudaFree(m->feec_c_qn_t);
  cudaFree(m->feec_u_qn_t);
  cudaFree(m->feec_w_qp);
  cudaFree(m->feec_c_qp);
  cudaFree(m->feec_u_qp);
  cudaFree(m->feec_u_qt);
  cudaFree(m->feec_div);
  cudaFree(m->feec_pos_nse);
  cudaFree(m->feec_general_cells);
  cudaFree(m->feec_fast_cells);
  cudaFree(m->feec_pos_pre);
  free_blockmat(m->nse);
  free_blockmat(m->pre);
  free_blockmat(m->tmass);
  free_blockmat(m->tstiff);
  free_blockmat(m->tmat);
  cudaFree(m->nse_rhs);
  cudaFree(m->temp_rhs);
  dcp_owner_plan_free(m->owner_nse);
  dcp_owner_plan_free(m->owner_pre);
  dcp_fast_plan_free(m->fast_nse);
  dcp_fast_plan_free(m->fast_pre);
  dcp_masked_plan_free(m->masked_nse);
  dcp_masked_plan_free(m->masked_pre);
  delete m;
  return DCP_OK;
}

int dcp_model_create(dcp_ctx* ctx, const dcp_model_desc* d, dcp_model** out) {
  if (!ctx || !d || !out) return DCP_ERR_ARG;
  *out = nullptr;
  if (d->dim != 2 && d->dim != 3) {
    dcp_set_error("dim must be 2 or 3");
    return DCP_ERR_ARG;
  }
  if (d->family != DCP_FAMILY_CLASSIC && d->family != DCP_FAMILY_FEEC) {
    dcp_set_error("unknown element family");
    return DCP_ERR_ARG;
  }
  const int dim = d->dim;
  const bool feec = d->family == DCP_FAMILY_FEEC;
  const int nu = dim == 3 ? 27 : 9, np = dim == 3 ? 8 : 4;
  if (!feec && (d->ndu != nu || d->ndp != np || d->nq_nse != nu || d->nse_n_local != dim * nu + np || d->nse_n_blocks != 2)) {
    dcp_set_error("classic family expects Q2^dim x Q1 with QGauss(3): ndu=3^dim, ndp=2^dim, nq_nse=3^dim");
    return DCP_ERR_ARG;
  }
  if (feec && (dim != 3 || d->nse_n_local != 19 || d->nse_n_blocks != 3 || d->nq_nse > 27 || d->nq_pre > 27 ||
               !d->nse_sign || !d->feec_phi_w_qn || !d->feec_curl_w_qn || !d->feec_phi_u_qn || !d->feec_phi_w_qp ||
               !d->feec_curl_w_qp || !d->feec_phi_u_qp || !d->feec_phi_u_qt || !d->feec_div_u || !d->geom_qp)) {
    dcp_set_error("FEEC family expects dim=3, 19 dofs per cell, 3 blocks, <=27 quadrature points and all feec_* tables");
    return DCP_ERR_ARG;
  }
  if (!d->geom_qn || !d->geom_qt) {
    dcp_set_error("geom_qn / geom_qt missing (host records, or device records from dcp_geometry_create)");
    return DCP_ERR_ARG;
  }
  DCP_CUDA(cudaSetDevice(ctx->device));
  dcp_model* m = new dcp_model;
  m->ctx = ctx;
  m->dim = dim;
  m->family = d->family;
  m->n_cells = d->n_cells;
  m->nse_n_local = d->nse_n_local;
  m->nse_nb = d->nse_n_blocks;
  m->temp_n_local = d->temp_n_local;
  m->nse_n_dofs = 0;
  for (int b = 0; b < m->nse_nb; ++b) m->nse_n_dofs += d->nse_block_size[b];
  m->temp_n_dofs = d->temp_cs.n_dofs;
  m->nq_nse = d->nq_nse;
  m->nq_temp = d->nq_temp;
  m->ndu = d->ndu;
  m->ndp = d->ndp;
  m->ndt = d->ndt;
  int rc = DCP_OK;
  auto fail = [&](int code) {
    if (d->geom_on_device) {  // device mapping buffers stay the caller's unless the create succeeds
      m->geom_qn = m->geom_qt = m->geom_qp = nullptr;
      m->geom_shared = false;
    }
    dcp_model_destroy(m);
    return code;
  };
#define M_TRY(x)                       \
  do {                                 \
    rc = (x);                          \
    if (rc != DCP_OK) return fail(rc); \
  } while (0)
  const int64_t nc = d->n_cells;
  const int rec = feec ? (1 + dim * dim + dim + dim * dim + 1) : (1 + dim * dim + dim);
  const int gs_n = d->nq_nse * rec, gs_t = d->nq_temp * rec;
  m->gs_n = gs_n;
  m->gs_t = gs_t;
  m->gs_p = d->nq_pre * rec;
  m->nq_pre = d->nq_pre;
  M_TRY(dcp_upload(ctx, &m->nse_l2g, d->nse_l2g, nc * d->nse_n_local));
  M_TRY(dcp_upload(ctx, &m->temp_l2g, d->temp_l2g, nc * d->temp_n_local));
  M_TRY(dcp_upload(ctx, &m->nse_local_field, d->nse_local_field, d->nse_n_local));
  M_TRY(dcp_upload(ctx, &m->nse_local_base, d->nse_local_base, d->nse_n_local));
  m->h_local_field.assign(d->nse_local_field, d->nse_local_field + d->nse_n_local);
  m->h_local_base.assign(d->nse_local_base, d->nse_local_base + d->nse_n_local);
  M_TRY(upload_cs(ctx, d->nse_cs, m->nse_cs));
  M_TRY(upload_cs(ctx, d->temp_cs, m->temp_cs));
  if (d->nse_cs.n_dofs != m->nse_n_dofs) {
    dcp_set_error("nse_cs.n_dofs != sum of block sizes");
    return fail(DCP_ERR_ARG);
  }
  if (!feec) {
    M_TRY(dcp_upload(ctx, &m->phi_u_qn, d->phi_u_qn, (int64_t)d->nq_nse * d->ndu));
    M_TRY(dcp_upload(ctx, &m->dphi_u_qn, d->dphi_u_qn, (int64_t)d->nq_nse * d->ndu * dim));
    M_TRY(dcp_upload(ctx, &m->phi_p_qn, d->phi_p_qn, (int64_t)d->nq_nse * d->ndp));
    M_TRY(dcp_upload(ctx, &m->phi_u_qt, d->phi_u_qt, (int64_t)d->nq_temp * d->ndu));
  } else {
    M_TRY(dcp_upload(ctx, &m->nse_sign, d->nse_sign, nc * d->nse_n_local));
    M_TRY(dcp_upload(ctx, &m->feec_w_qn, d->feec_phi_w_qn, (int64_t)d->nq_nse * 36));
    M_TRY(dcp_upload(ctx, &m->feec_c_qn, d->feec_curl_w_qn, (int64_t)d->nq_nse * 36));
    M_TRY(dcp_upload(ctx, &m->feec_u_qn, d->feec_phi_u_qn, (int64_t)d->nq_nse * 18));
    {
      // point-fastest copies [function][component][q] for the CTA-per-cell system kernel, whose threads of one warp
      // are consecutive quadrature points: a warp-wide load then touches 2 lines instead of 27
      auto upload_t = [&](double** dst, const double* src, int nq, int nf) -> int {
        std::vector<double> t((size_t)nq * nf * 3);
        for (int q = 0; q < nq; ++q)
          for (int k = 0; k < nf * 3; ++k) t[(size_t)k * nq + q] = src[(size_t)q * nf * 3 + k];
        int r = dcp_upload(ctx, dst, t.data(), (int64_t)t.size());
        cudaStreamSynchronize(ctx->stream);  // `t` dies here
        return r;
      };
      M_TRY(upload_t(&m->feec_w_qn_t, d->feec_phi_w_qn, d->nq_nse, 12));
      M_TRY(upload_t(&m->feec_c_qn_t, d->feec_curl_w_qn, d->nq_nse, 12));
      M_TRY(upload_t(&m->feec_u_qn_t, d->feec_phi_u_qn, d->nq_nse, 6));
    }
    M_TRY(dcp_upload(ctx, &m->feec_w_qp, d->feec_phi_w_qp, (int64_t)d->nq_pre * 36));
    M_TRY(dcp_upload(ctx, &m->feec_c_qp, d->feec_curl_w_qp, (int64_t)d->nq_pre * 36));
    M_TRY(dcp_upload(ctx, &m->feec_u_qp, d->feec_phi_u_qp, (int64_t)d->nq_pre * 18));
    M_TRY(dcp_upload(ctx, &m->feec_u_qt, d->feec_phi_u_qt, (int64_t)d->nq_temp * 18));
    M_TRY(dcp_upload(ctx, &m->feec_div, d->feec_div_u, 6));
    if (d->geom_on_device)
      m->geom_qp = const_cast<double*>(d->geom_qp);
    else
      M_TRY(dcp_upload(ctx, &m->geom_qp, d->geom_qp, nc * (int64_t)m->gs_p));
  }
  M_TRY(dcp_upload(ctx, &m->phi_t_qn, d->phi_t_qn, (int64_t)d->nq_nse * d->ndt));
  M_TRY(dcp_upload(ctx, &m->phi_t_qt, d->phi_t_qt, (int64_t)d->nq_temp * d->ndt));
  M_TRY(dcp_upload(ctx, &m->dphi_t_qt, d->dphi_t_qt, (int64_t)d->nq_temp * d->ndt * dim));
  {
    // point-fastest copies of the temperature-rule tables: the temperature kernels read them with lane = point
    auto upload_T = [&](double** dst, const double* src, int nq, int inner) -> int {
      std::vector<double> t((size_t)nq * inner);
      for (int q = 0; q < nq; ++q)
        for (int k = 0; k < inner; ++k) t[(size_t)k * nq + q] = src[(size_t)q * inner + k];
      int r = dcp_upload(ctx, dst, t.data(), (int64_t)t.size());
      cudaStreamSynchronize(ctx->stream);  // `t` dies here
      return r;
    };
    M_TRY(upload_T(&m->phi_t_qt_T, d->phi_t_qt, d->nq_temp, d->ndt));
    M_TRY(upload_T(&m->dphi_t_qt_T, d->dphi_t_qt, d->nq_temp, d->ndt * dim));
    if (!feec) M_TRY(upload_T(&m->phi_u_qt_T, d->phi_u_qt, d->nq_temp, d->ndu));
  }
  if (d->geom_on_device)
    m->geom_qn = const_cast<double*>(d->geom_qn);
  else
    M_TRY(dcp_upload(ctx, &m->geom_qn, d->geom_qn, nc * gs_n));
  if (d->geom_qt == d->geom_qn && gs_n == gs_t) {
    m->geom_qt = m->geom_qn;
    m->geom_shared = true;
  } else if (d->geom_on_device)
    m->geom_qt = const_cast<double*>(d->geom_qt);
  else
    M_TRY(dcp_upload(ctx, &m->geom_qt, d->geom_qt, nc * gs_t));

  // matrices
  auto setup_blocks = [&](BlockMat& M, const dcp_csr_desc pat[DCP_MAXB][DCP_MAXB]) -> int {
    M.nb = m->nse_nb;
    M.start[0] = 0;
    for (int b = 0; b < M.nb; ++b) M.start[b + 1] = M.start[b] + d->nse_block_size[b];
    for (int b = M.nb + 1; b <= DCP_MAXB; ++b) M.start[b] = M.start[M.nb];
    for (int i = 0; i < M.nb; ++i)
      for (int j = 0; j < M.nb; ++j) {
        if (pat[i][j].rowptr &&
            (pat[i][j].n_rows != d->nse_block_size[i] || pat[i][j].n_cols != d->nse_block_size[j])) {
          dcp_set_error("block pattern shape does not match the block sizes");
          return DCP_ERR_ARG;
        }
        DCP_TRY(upload_csr_pattern(ctx, pat[i][j], M.blk[i][j], true));
        M.blk[i][j].n_rows = d->nse_block_size[i];
        M.blk[i][j].n_cols = d->nse_block_size[j];
      }
    return DCP_OK;
  };
  M_TRY(setup_blocks(m->nse, d->nse_pattern));
  M_TRY(setup_blocks(m->pre, d->pre_pattern));
  m->tmass.nb = m->tstiff.nb = m->tmat.nb = 1;
  for (BlockMat* T : {&m->tmass, &m->tstiff, &m->tmat}) {
    T->start[0] = 0;
    for (int b = 1; b <= DCP_MAXB; ++b) T->start[b] = m->temp_n_dofs;
  }
  M_TRY(upload_csr_pattern(ctx, d->temp_pattern, m->tmass.blk[0][0], true));
  M_TRY(share_csr_pattern(ctx, m->tmass.blk[0][0], m->tstiff.blk[0][0]));
  M_TRY(share_csr_pattern(ctx, m->tmass.blk[0][0], m->tmat.blk[0][0]));
  rc = cudaMalloc((void**)&m->nse_rhs, sizeof(double) * (size_t)std::max<int64_t>(m->nse_n_dofs, 1)) == cudaSuccess ? DCP_OK : DCP_ERR_CUDA;
  if (rc != DCP_OK) {
    dcp_set_error("cudaMalloc nse_rhs failed");
    return fail(rc);
  }
  rc = cudaMalloc((void**)&m->temp_rhs, sizeof(double) * (size_t)std::max<int64_t>(m->temp_n_dofs, 1)) == cudaSuccess ? DCP_OK : DCP_ERR_CUDA;
  if (rc != DCP_OK) {
    dcp_set_error("cudaMalloc temp_rhs failed");
    return fail(rc);
  }
  m->n_owned_cells = d->n_owned_cells;
  if (d->cell_vertices) M_TRY(dcp_upload(ctx, &m->cell_vertices, d->cell_vertices, nc * (int64_t)(dim << dim)));
  if (!feec) {
    std::vector<int32_t> vd((size_t)dim * d->ndu, 0);
    for (int k = 0; k < d->nse_n_local; ++k)
      if (d->nse_local_field[k] < dim) vd[(size_t)d->nse_local_field[k] * d->ndu + d->nse_local_base[k]] = k;
    M_TRY(dcp_upload(ctx, &m->vel_dof, vd.data(), (int64_t)vd.size()));
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(DCP_ERR_CUDA);
  }
  // temperature matrices: scatter positions for the cells without constrained temperature dofs
  {
    const int nd = d->temp_n_local;
    std::vector<int32_t> lod((size_t)m->temp_n_dofs, -1);
    for (int64_t l = 0; l < d->temp_cs.n_lines; ++l) lod[d->temp_cs.line_dof[l]] = (int32_t)l;
    std::vector<uint16_t> tp((size_t)nc * nd * nd, 0xffff);
    const int64_t* rp = d->temp_pattern.rowptr;
    const int32_t* col = d->temp_pattern.col;
    if (rp && col) {
#pragma omp parallel for schedule(static)
      for (int64_t c = 0; c < nc; ++c) {
        const int32_t* idx = d->temp_l2g + c * nd;
        bool ok = true;
        for (int i = 0; i < nd && ok; ++i) ok = lod[idx[i]] < 0;
        uint16_t* out = &tp[(size_t)c * nd * nd];
        for (int i = 0; i < nd && ok; ++i)
          for (int j = 0; j < nd && ok; ++j) {
            const int32_t* b = col + rp[idx[i]];
            const int32_t* e = col + rp[idx[i] + 1];
            const int32_t* p = std::lower_bound(b, e, idx[j]);
            ok = p != e && *p == idx[j] && (p - b) < 0xffff;
            if (ok) out[i * nd + j] = (uint16_t)(p - b);
          }
        if (!ok) out[0] = 0xffff;
      }
    }
    M_TRY(dcp_upload(ctx, &m->temp_pos, tp.data(), (int64_t)tp.size()));
    // the two classes as lists (Q2 temperature in 3-D: tensor-core kernel on the first, general kernel on the second)
    {
      std::vector<int32_t> fast, general;
      for (int64_t c = 0; c < nc; ++c) (tp[(size_t)c * nd * nd] != 0xffff ? fast : general).push_back((int32_t)c);
      m->n_temp_fast = (int64_t)fast.size();
      m->n_temp_general = (int64_t)general.size();
      if (!fast.empty()) M_TRY(dcp_upload(ctx, &m->temp_fast_cells, fast.data(), (int64_t)fast.size()));
      if (!general.empty()) M_TRY(dcp_upload(ctx, &m->temp_general_cells, general.data(), (int64_t)general.size()));
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(DCP_ERR_CUDA);
    }
    // cells whose right-hand side needs matrix_for_bc: an inhomogeneously constrained temperature dof
    std::vector<uint8_t> flag((size_t)std::max<int64_t>(nc, 1), 0);
    std::vector<int32_t> bc_cells;
    for (int64_t c = 0; c < nc; ++c) {
      const int32_t* idx = d->temp_l2g + c * nd;
      for (int i = 0; i < nd; ++i) {
        const int32_t li = lod[idx[i]];
        if (li >= 0 && d->temp_cs.inhom[li] != 0.0) flag[(size_t)c] = 1;
      }
      if (flag[(size_t)c]) bc_cells.push_back((int32_t)c);
    }
    m->n_temp_bc_cells = (int64_t)bc_cells.size();
    M_TRY(dcp_upload(ctx, &m->temp_bc_flag, flag.data(), (int64_t)flag.size()));
    if (!bc_cells.empty()) M_TRY(dcp_upload(ctx, &m->temp_bc_cells, bc_cells.data(), (int64_t)bc_cells.size()));
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(DCP_ERR_CUDA);
  }
  // cells holding a constrained NSE dof
  {
    std::vector<int32_t> lod((size_t)m->nse_n_dofs, -1);
    for (int64_t l = 0; l < d->nse_cs.n_lines; ++l) lod[d->nse_cs.line_dof[l]] = (int32_t)l;
    std::vector<int32_t> cells;
    for (int64_t c = 0; c < nc; ++c) {
      const int32_t* idx = d->nse_l2g + c * d->nse_n_local;
      bool any = false;
      for (int i = 0; i < d->nse_n_local && !any; ++i) any = lod[idx[i]] >= 0;
      if (any) cells.push_back((int32_t)c);
    }
    m->n_nse_constrained_cells = (int64_t)cells.size();
    M_TRY(dcp_upload(ctx, &m->nse_constrained_cells, cells.data(), (int64_t)cells.size()));
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(DCP_ERR_CUDA);
  }
  if (feec) {
    M_TRY(dcp_feec_positions_build(m, d, true, &m->feec_pos_nse));
    M_TRY(dcp_feec_positions_build(m, d, false, &m->feec_pos_pre));
  }
  // position tables for the unconstrained cells (DCP_STRATEGY_POSITIONS; classic family)
  if (!feec) {
    if (dim == 3) {
      M_TRY(dcp_masked_plan_build(m, d, true, &m->masked_nse));
      M_TRY(dcp_masked_plan_build(m, d, false, &m->masked_pre));
      if (m->masked_nse->gather) {
        m->strategy = DCP_STRATEGY_STAGED;
        M_TRY(dcp_gather_plan_attach_pre(m, d, m->masked_nse->gather, m->masked_nse, m->masked_pre));
      }
    } else {
      M_TRY(dcp_fast_plan_build(m, d, true, &m->fast_nse));
      M_TRY(dcp_fast_plan_build(m, d, false, &m->fast_pre));
    }
    // row-owner tiles (DCP_STRATEGY_OWNER): optional -- a numbering that is not node-blocked keeps POSITIONS
    if (dim == 3 && d->build_owner_plan) {
      rc = dcp_owner_plan_build(m, true, d);
      if (rc == DCP_OK) rc = dcp_owner_plan_build(m, false, d);
      if (rc == DCP_ERR_CUDA) return fail(rc);
      if (rc != DCP_OK) {
        dcp_owner_plan_free(m->owner_nse);
        dcp_owner_plan_free(m->owner_pre);
        m->owner_nse = m->owner_pre = nullptr;
      }
    }
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(DCP_ERR_CUDA);
#undef M_TRY
  *out = m;
  return DCP_OK;
}

int dcp_model_set_strategy(dcp_model* m, int strategy) {
  if (!m || strategy < DCP_STRATEGY_SEARCH || strategy > DCP_STRATEGY_STAGED) return DCP_ERR_ARG;
  if (strategy == DCP_STRATEGY_STAGED && !(m->masked_nse && m->masked_nse->gather)) {
    dcp_set_error("DCP_STRATEGY_STAGED is not available for this model (classic 3-D family, every cell in the position plan, rows within the gather capacity)");
    return DCP_ERR_STATE;
  }
  if (strategy == DCP_STRATEGY_OWNER && (!m->owner_nse || !m->owner_pre)) {
    dcp_set_error("DCP_STRATEGY_OWNER is not available for this model (needs dcp_model_desc.build_owner_plan, classic 3-D family, node-blocked numbering)");
    return DCP_ERR_STATE;
  }
  m->strategy = strategy;
  m->pre_fused_valid = false;
  return DCP_OK;
}

int dcp_model_get_strategy(const dcp_model* m) { return m ? m->strategy : -1; }

int dcp_model_set_owned(dcp_model* m, const int64_t* nse_owned_per_block, int64_t temp_owned) {
  if (!m || !nse_owned_per_block) return DCP_ERR_ARG;
  for (int b = 0; b < m->nse_nb; ++b) {
    if (nse_owned_per_block[b] < 0 || nse_owned_per_block[b] > m->nse.start[b + 1] - m->nse.start[b]) {
      dcp_set_error("dcp_model_set_owned: owned count exceeds the block size");
      return DCP_ERR_ARG;
    }
    m->nse.owned[b] = m->pre.owned[b] = nse_owned_per_block[b];
  }
  if (temp_owned < 0 || temp_owned > m->temp_n_dofs) return DCP_ERR_ARG;
  m->tmass.owned[0] = m->tstiff.owned[0] = m->tmat.owned[0] = temp_owned;
  // rows that read ghost columns, per block row: dcp_vmult_rows / dcp_block_vmult_rows can then run the other rows
  // while the ghost exchange is still in flight
  dcp_ctx* ctx = m->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  for (int b = 0; b < m->nse_nb; ++b) {
    DCP_TRY(dcp_build_ghost_rows(ctx, m->nse, b, nse_owned_per_block));
    DCP_TRY(dcp_build_ghost_rows(ctx, m->pre, b, nse_owned_per_block));
  }
  DCP_TRY(dcp_build_ghost_rows(ctx, m->tmat, 0, &temp_owned));
  for (BlockMat* M : {&m->tmass, &m->tstiff}) {  // same pattern as tmat
    M->owns_ghost_rows = false;
    M->ghost_flag[0] = m->tmat.ghost_flag[0];
    M->ghost_list[0] = m->tmat.ghost_list[0];
    M->n_ghost_rows[0] = m->tmat.n_ghost_rows[0];
  }
  return DCP_OK;
}

int dcp_gather_f64(dcp_ctx* ctx, int64_t n, const int32_t* idx_dev, const double* src_dev, double* dst_dev) {
  if (!ctx || n < 0) return DCP_ERR_ARG;
  return dcp_launch_gather(ctx, n, idx_dev, src_dev, dst_dev, false);
}
int dcp_scatter_f64(dcp_ctx* ctx, int64_t n, const int32_t* idx_dev, const double* src_dev, double* dst_dev) {
  if (!ctx || n < 0) return DCP_ERR_ARG;
  return dcp_launch_gather(ctx, n, idx_dev, src_dev, dst_dev, true);
}

int dcp_assemble_nse_system(dcp_model* m, const dcp_params* p, const double* old_nse, const double* old_temp, int mem) {
  if (!m || !p || !old_nse || !old_temp) return DCP_ERR_ARG;
  dcp_ctx* ctx = m->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  const double *d_nse, *d_temp;
  DCP_TRY(dcp_stage_in(ctx, 0, old_nse, m->nse_n_dofs, mem, &d_nse));
  DCP_TRY(dcp_stage_in(ctx, 1, old_temp, m->temp_n_dofs, mem, &d_temp));
  DCP_CUDA(cudaMemsetAsync(m->nse_rhs, 0, sizeof(double) * (size_t)m->nse_n_dofs, ctx->stream));
  if (m->family == DCP_FAMILY_FEEC) {
    DCP_TRY(zero_blockmat(ctx, m->nse));
    DCP_TRY(dcp_launch_feec(m, *p, true, d_nse, d_temp));
    // Jacobi of Mw = block(0,0) and Mu = block(1,1) used by solve_NSE_block_preconditioned (:1283-1304)
    DCP_TRY(refresh_jacobi(ctx, m->nse));
  } else if (m->strategy == DCP_STRATEGY_OWNER) {
    // every CSR value of an unconstrained row is written once by its owning tile; then rhs (all cells) and the
    // constrained-dof fix-up, which accumulates: those entries must start from zero on every pass
    DCP_TRY(zero_blockmat(ctx, m->nse));
    DCP_TRY(dcp_launch_th_owner(m, *p, true));
    DCP_TRY(dcp_launch_th_rhs(m, *p, d_nse, d_temp));
    DCP_TRY(dcp_launch_th_cells(m, *p, true, d_nse, d_temp, m->nse_constrained_cells, m->n_nse_constrained_cells, true));
  } else if (m->strategy == DCP_STRATEGY_STAGED) {
    // write-once: no zero-fill, no reductions into the matrix (assemble_th_stage.cu)
    DCP_TRY(dcp_launch_th_staged(m, *p, m->masked_nse, d_nse, d_temp));
  } else if (m->strategy == DCP_STRATEGY_POSITIONS) {
    DCP_TRY(zero_blockmat(ctx, m->nse));
    if (m->dim == 3) {
      // every cell of the plan on the tensor cores, Dirichlet / no-normal-flux lines resolved in the epilogue; cells
      // with other constraint kinds (periodic, inhomogeneous, ...) or an unverified layout take the general kernel
      DCP_TRY(dcp_launch_th_mma(m, *p, true, m->masked_nse, d_nse, d_temp));
      if (m->masked_nse->n_other > 0)
        DCP_TRY(dcp_launch_th_cells(m, *p, true, d_nse, d_temp, m->masked_nse->other_cells, m->masked_nse->n_other, false));
    } else {
      DCP_TRY(dcp_launch_th_fast(m, *p, true, m->fast_nse, d_nse, d_temp));
      int64_t ng = 0;
      dcp_fast_plan_counts(m->fast_nse, &ng);
      if (ng > 0) DCP_TRY(dcp_launch_th_cells(m, *p, true, d_nse, d_temp, dcp_fast_plan_general_cells(m->fast_nse), ng, false));
    }
  } else {
    DCP_TRY(zero_blockmat(ctx, m->nse));
    DCP_TRY(dcp_launch_th_cells(m, *p, true, d_nse, d_temp, nullptr, -1, false));
  }
  if (mem == DCP_HOST) DCP_TRY(dcp_check_device_errors(ctx, "dcp_assemble_nse_system"));
  return DCP_OK;
}

int dcp_assemble_nse_preconditioner(dcp_model* m, const dcp_params* p) {
  if (!m || !p) return DCP_ERR_ARG;
  dcp_ctx* ctx = m->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  if (m->family == DCP_FAMILY_FEEC) {
    DCP_TRY(zero_blockmat(ctx, m->pre));
    DCP_TRY(dcp_launch_feec(m, *p, false, nullptr, nullptr));
  } else if (m->strategy == DCP_STRATEGY_OWNER) {
    DCP_TRY(zero_blockmat(ctx, m->pre));
    DCP_TRY(dcp_launch_th_owner(m, *p, false));
    DCP_TRY(dcp_launch_th_cells(m, *p, false, nullptr, nullptr, m->nse_constrained_cells, m->n_nse_constrained_cells, true));
  } else if (m->strategy == DCP_STRATEGY_STAGED && m->pre_fused_valid && m->pre_fused_dt == p->dt && m->pre_fused_inv_re == p->inv_re) {
    // nse_preconditioner_matrix was written by the staged system pass for the same parameters: the reference calls
    // the two assemblers back to back (boussinesq_model.tpp:1869-1880), the velocity block m + nu k is a by-product
    // of the system's contraction.  Only the Jacobi set-up is left.
  } else if (m->strategy == DCP_STRATEGY_POSITIONS || m->strategy == DCP_STRATEGY_STAGED) {
    DCP_TRY(zero_blockmat(ctx, m->pre));
    if (m->dim == 3) {
      DCP_TRY(dcp_launch_th_mma(m, *p, false, m->masked_pre, nullptr, nullptr));
      if (m->masked_pre->n_other > 0)
        DCP_TRY(dcp_launch_th_cells(m, *p, false, nullptr, nullptr, m->masked_pre->other_cells, m->masked_pre->n_other, false));
    } else {
      DCP_TRY(dcp_launch_th_fast(m, *p, false, m->fast_pre, nullptr, nullptr));
      int64_t ng = 0;
      dcp_fast_plan_counts(m->fast_pre, &ng);
      if (ng > 0) DCP_TRY(dcp_launch_th_cells(m, *p, false, nullptr, nullptr, dcp_fast_plan_general_cells(m->fast_pre), ng, false));
    }
  } else {
    DCP_TRY(zero_blockmat(ctx, m->pre));
    DCP_TRY(dcp_launch_th_cells(m, *p, false, nullptr, nullptr, nullptr, -1, false));
  }
  // build_nse_preconditioner: Jacobi of block(0,0) and block(1,1)  (boussinesq_model.tpp:531-539)
  DCP_TRY(refresh_jacobi(ctx, m->pre));
  return DCP_OK;
}

int dcp_assemble_temperature_matrix(dcp_model* m, const dcp_params* p) {
  if (!m || !p) return DCP_ERR_ARG;
  dcp_ctx* ctx = m->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  DCP_TRY(zero_blockmat(ctx, m->tmass));
  DCP_TRY(zero_blockmat(ctx, m->tstiff));
  DCP_TRY(dcp_launch_temperature_matrix(m, *p));
  m->temp_matrices_ready = true;
  return DCP_OK;
}

int dcp_assemble_temperature_rhs(dcp_model* m, const dcp_params* p, const double* old_temp, const double* nse_solution, int mem) {
  if (!m || !p || !old_temp || !nse_solution) return DCP_ERR_ARG;
  if (!m->temp_matrices_ready) {
    dcp_set_error("dcp_assemble_temperature_rhs: call dcp_assemble_temperature_matrix first");
    return DCP_ERR_STATE;
  }
  dcp_ctx* ctx = m->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  const double *d_temp, *d_nse;
  DCP_TRY(dcp_stage_in(ctx, 0, nse_solution, m->nse_n_dofs, mem, &d_nse));
  DCP_TRY(dcp_stage_in(ctx, 1, old_temp, m->temp_n_dofs, mem, &d_temp));
  // temperature_matrix.copy_from(mass); temperature_matrix.add(dt/n, stiffness)  (:975-978)
  DevCsr& T = m->tmat.blk[0][0];
  DCP_TRY(dcp_launch_axpby_values(ctx, T.nnz, m->tmass.blk[0][0].val, m->tstiff.blk[0][0].val,
                                  p->dt / p->nse_interval, T.val));
  // T_preconditioner = Jacobi(temperature_matrix)  (:980-986)
  DCP_TRY(refresh_jacobi(ctx, m->tmat));
  DCP_CUDA(cudaMemsetAsync(m->temp_rhs, 0, sizeof(double) * (size_t)m->temp_n_dofs, ctx->stream));
  DCP_TRY(dcp_launch_temperature_rhs(m, *p, d_temp, d_nse));
  if (mem == DCP_HOST) DCP_TRY(dcp_check_device_errors(ctx, "dcp_assemble_temperature_rhs"));
  return DCP_OK;
}

int dcp_velocity_extrema(dcp_model* m, const double* nse_solution, int mem, double* result_host) {
  if (!m || !nse_solution || !result_host) return DCP_ERR_ARG;
  dcp_ctx* ctx = m->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  if (m->family == DCP_FAMILY_FEEC && !m->cell_vertices) {
    dcp_set_error("dcp_velocity_extrema: the FEEC family needs desc.cell_vertices");
    return DCP_ERR_STATE;
  }
  const double* d_nse = nullptr;
  DCP_TRY(dcp_stage_in(ctx, 0, nse_solution, m->nse_n_dofs, mem, &d_nse));
  double* d_out = nullptr;
  DCP_TRY(dcp_stage_out_alloc(ctx, 2, nullptr, 2, DCP_HOST, &d_out));
  DCP_TRY(dcp_launch_velocity_extrema(m, d_nse, d_out));
  DCP_CUDA(cudaMemcpyAsync(result_host, d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!m->cell_vertices) result_host[1] = -1.0;  // no diameters: CFL not available
  return DCP_OK;
}

int dcp_constraints_distribute(dcp_model* m, int space, double* x, int mem) {
  if (!m || !x || (space != 0 && space != 1)) return DCP_ERR_ARG;
  dcp_ctx* ctx = m->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  const DevCs& cs = space == 0 ? m->nse_cs : m->temp_cs;
  const int64_t n = space == 0 ? m->nse_n_dofs : m->temp_n_dofs;
  const double* in = nullptr;
  DCP_TRY(dcp_stage_in(ctx, 0, x, n, mem, &in));
  double* dx = const_cast<double*>(in);
  DCP_TRY(dcp_launch_distribute(ctx, cs, dx));
  if (mem == DCP_HOST) {
    DCP_CUDA(cudaMemcpyAsync(x, dx, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return DCP_OK;
}

int dcp_matrix_info(const dcp_model* m, int which, int bi, int bj, int64_t* n_rows, int64_t* n_cols, int64_t* nnz) {
  DevCsr* A;
  DCP_TRY(get_block(const_cast<dcp_model*>(m), which, bi, bj, &A));
  if (n_rows) *n_rows = A->n_rows;
  if (n_cols) *n_cols = A->n_cols;
  if (nnz) *nnz = A->nnz;
  return DCP_OK;
}

int dcp_matrix_values_device(dcp_model* m, int which, int bi, int bj, double** out) {
  DevCsr* A;
  DCP_TRY(get_block(m, which, bi, bj, &A));
  *out = A->val;
  return DCP_OK;
}

int dcp_matrix_download(dcp_model* m, int which, int bi, int bj, double* host_values) {
  DevCsr* A;
  DCP_TRY(get_block(m, which, bi, bj, &A));
  DCP_TRY(dcp_check_device_errors(m->ctx, "dcp_matrix_download"));
  if (A->nnz == 0) return DCP_OK;
  return dcp_memcpy_d2h(m->ctx, host_values, A->val, (int64_t)sizeof(double) * A->nnz);
}

int dcp_matrix_upload(dcp_model* m, int which, int bi, int bj, const double* host_values) {
  DevCsr* A;
  BlockMat* M;
  DCP_TRY(get_block(m, which, bi, bj, &A, &M));
  if (A->nnz == 0) return DCP_OK;
  DCP_TRY(dcp_memcpy_h2d(m->ctx, A->val, host_values, (int64_t)sizeof(double) * A->nnz));
  if (bi == bj) DCP_TRY(refresh_jacobi(m->ctx, *M));
  if (which == DCP_MAT_TEMP_MASS || which == DCP_MAT_TEMP_STIFF) m->temp_matrices_ready = true;
  if (which == DCP_MAT_NSE_PRECOND) m->pre_fused_valid = false;   // the caller's values replace the fused pass's
  return DCP_OK;
}

int dcp_vector_device(dcp_model* m, int which, double** out, int64_t* n) {
  if (!m || !out) return DCP_ERR_ARG;
  if (which == DCP_VEC_NSE_RHS) {
    *out = m->nse_rhs;
    if (n) *n = m->nse_n_dofs;
  } else if (which == DCP_VEC_TEMP_RHS) {
    *out = m->temp_rhs;
    if (n) *n = m->temp_n_dofs;
  } else
    return DCP_ERR_ARG;
  return DCP_OK;
}

int dcp_vector_download(dcp_model* m, int which, double* host) {
  double* d;
  int64_t n;
  DCP_TRY(dcp_vector_device(m, which, &d, &n));
  DCP_TRY(dcp_check_device_errors(m->ctx, "dcp_vector_download"));
  return dcp_memcpy_d2h(m->ctx, host, d, (int64_t)sizeof(double) * n);
}

// After dcp_model_set_owned only the owned rows are computed; a host destination would receive the staging buffer's
// stale tail for the other rows.  Row-distributed products take device vectors (the ghost exchange lives there, too).
static int partitioned_host_error() {
  dcp_set_error("row-distributed model (dcp_model_set_owned): operator products take DCP_DEVICE vectors only");
  return DCP_ERR_STATE;
}

static int vmult_impl(dcp_model* m, int which, int bi, int bj, double* dst, const double* src, int mem, bool add) {
  if (!m || !dst || !src) return DCP_ERR_ARG;
  DevCsr* A;
  BlockMat* BM;
  DCP_TRY(get_block(m, which, bi, bj, &A, &BM));
  dcp_ctx* ctx = m->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  if (mem == DCP_HOST && BM->owned[bi] >= 0) return partitioned_host_error();
  const double* dx;
  double* dy;
  DCP_TRY(dcp_stage_in(ctx, 0, src, A->n_cols, mem, &dx));
  if (add && mem == DCP_HOST) {
    const double* tmp;
    DCP_TRY(dcp_stage_in(ctx, 1, dst, A->n_rows, mem, &tmp));
    dy = const_cast<double*>(tmp);
  } else
    DCP_TRY(dcp_stage_out_alloc(ctx, 1, dst, A->n_rows, mem, &dy));
  DCP_TRY(dcp_launch_spmv(ctx, *A, dx, dy, add, BM->owned[bi]));
  return dcp_stage_out_finish(ctx, 1, dst, A->n_rows, mem);
}

// one launch of block (bi,bj) restricted to a row class of block row bi
static int spmv_rows(dcp_ctx* ctx, BlockMat& M, int bi, int bj, const double* x, double* y, bool add, int rows) {
  const DevCsr& A = M.blk[bi][bj];
  if (rows == DCP_ROWS_ALL) return dcp_launch_spmv(ctx, A, x, y, add, M.owned[bi]);
  if (!M.ghost_flag[bi]) {
    dcp_set_error("row classes need dcp_model_set_owned first");
    return DCP_ERR_STATE;
  }
  if (rows == DCP_ROWS_INTERIOR) return dcp_launch_spmv(ctx, A, x, y, add, M.owned[bi], M.ghost_flag[bi]);
  return dcp_launch_spmv(ctx, A, x, y, add, M.owned[bi], nullptr, M.ghost_list[bi], M.n_ghost_rows[bi]);
}

int dcp_vmult_rows(dcp_model* m, int which, int bi, int bj, double* dst, const double* src, int rows) {
  if (!m || !dst || !src || rows < DCP_ROWS_ALL || rows > DCP_ROWS_GHOSTED) return DCP_ERR_ARG;
  DevCsr* A;
  BlockMat* BM;
  DCP_TRY(get_block(m, which, bi, bj, &A, &BM));
  DCP_CUDA(cudaSetDevice(m->ctx->device));
  return spmv_rows(m->ctx, *BM, bi, bj, src, dst, false, rows);
}

int dcp_block_vmult_rows(dcp_model* m, int which, double* dst, const double* src, int rows) {
  if (!m || !dst || !src || rows < DCP_ROWS_ALL || rows > DCP_ROWS_GHOSTED) return DCP_ERR_ARG;
  BlockMat* M = select_matrix(m, which);
  if (!M) return DCP_ERR_ARG;
  dcp_ctx* ctx = m->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  for (int r = 0; r < M->nb; ++r) {
    bool first = true;
    for (int c = 0; c < M->nb; ++c) {
      if (M->blk[r][c].nnz == 0) continue;
      DCP_TRY(spmv_rows(ctx, *M, r, c, src + M->start[c], dst + M->start[r], !first, rows));
      first = false;
    }
    if (first && rows != DCP_ROWS_GHOSTED)
      DCP_TRY(dcp_launch_fill(ctx, dst + M->start[r], M->owned[r] >= 0 ? M->owned[r] : M->start[r + 1] - M->start[r], 0.0));
  }
  return DCP_OK;
}

int dcp_vmult(dcp_model* m, int which, int bi, int bj, double* dst, const double* src, int mem) {
  return vmult_impl(m, which, bi, bj, dst, src, mem, false);
}
int dcp_vmult_add(dcp_model* m, int which, int bi, int bj, double* dst, const double* src, int mem) {
  return vmult_impl(m, which, bi, bj, dst, src, mem, true);
}

int dcp_block_vmult(dcp_model* m, int which, double* dst, const double* src, int mem) {
  if (!m || !dst || !src) return DCP_ERR_ARG;
  BlockMat* M = select_matrix(m, which);
  if (!M) return DCP_ERR_ARG;
  dcp_ctx* ctx = m->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = M->start[M->nb];
  if (mem == DCP_HOST && M->owned[0] >= 0) return partitioned_host_error();
  const double* dx;
  double* dy;
  DCP_TRY(dcp_stage_in(ctx, 0, src, n, mem, &dx));
  DCP_TRY(dcp_stage_out_alloc(ctx, 1, dst, n, mem, &dy));
  // BlockMatrixBase::vmult: dst.block(r) = sum_c block(r,c) * src.block(c)
  for (int r = 0; r < M->nb; ++r) {
    bool first = true;
    for (int c = 0; c < M->nb; ++c) {
      DevCsr& A = M->blk[r][c];
      if (A.nnz == 0) continue;
      DCP_TRY(dcp_launch_spmv(ctx, A, dx + M->start[c], dy + M->start[r], !first, M->owned[r]));
      first = false;
    }
    if (first)
      DCP_TRY(dcp_launch_fill(ctx, dy + M->start[r], M->owned[r] >= 0 ? M->owned[r] : M->start[r + 1] - M->start[r], 0.0));
  }
  return dcp_stage_out_finish(ctx, 1, dst, n, mem);
}

int dcp_jacobi_vmult(dcp_model* m, int which, int bi, double* dst, const double* src, int mem) {
  if (!m || !dst || !src) return DCP_ERR_ARG;
  BlockMat* M = select_matrix(m, which);
  if (!M || bi < 0 || bi >= M->nb || !M->diag_inv[bi]) {
    dcp_set_error("dcp_jacobi_vmult: diagonal not available (assemble first)");
    return DCP_ERR_STATE;
  }
  dcp_ctx* ctx = m->ctx;
  const int64_t n = M->start[bi + 1] - M->start[bi];
  if (mem == DCP_HOST && M->owned[bi] >= 0) return partitioned_host_error();
  const double* dx;
  double* dy;
  DCP_TRY(dcp_stage_in(ctx, 0, src, n, mem, &dx));
  DCP_TRY(dcp_stage_out_alloc(ctx, 1, dst, n, mem, &dy));
  DCP_TRY(dcp_launch_jacobi(ctx, M->owned[bi] >= 0 ? M->owned[bi] : n, M->diag_inv[bi], dx, dy));
  return dcp_stage_out_finish(ctx, 1, dst, n, mem);
}

}  // extern "C"
