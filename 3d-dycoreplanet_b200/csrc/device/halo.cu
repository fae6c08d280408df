// Multi-GPU data plane behind the C ABI: NCCL communicator, ghost-dof halo exchange for the row-distributed SpMV and
// all-reduced scalars for the Krylov solvers.
//
// What it replaces in the reference (one MPI rank per subdomain, everything hidden inside Trilinos / deal.II):
//   * the Epetra_Import of off-rank source entries inside every LA::SparseMatrix::vmult
//     (include/linear_algebra/schur_complement.hpp:143-150, block_schur_preconditioner.hpp:55, ...),
//   * the ghost refresh `nse_solution = distributed_nse_solution` (include/core/boussinesq_model.tpp:1241, 1444),
//   * the MPI_Allreduce behind l2_norm / operator* (:1165, 1427) and Utilities::MPI::max (:1050, 1094, 1467).
// One process per GPU.  The exchange is: pack kernel (owned boundary entries -> per-peer contiguous send buffer),
// ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd over NVLink, unpack kernel (receive buffer -> ghost slots of the
// same vector).  Everything is enqueued on CUDA streams; no host synchronisation inside a product.  With overlap the
// exchange runs on the communicator's own stream while the rows that read owned columns only are computed, and the
// rows with ghost columns follow (row classes of dcp_model_set_owned).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already loaded NCCL (torch) shares that copy, a
// single-GPU process never needs the library.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>

#include "dcp_internal.cuh"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.error = std::string("NCCL not found: ") + dlerror();
      return;
    }
    auto sym = [&](const char* n) {
      void* p = dlsym(api.handle, n);
      if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + n;
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  return api;
}

int nccl_ready() {
  NcclApi& n = nccl();
  if (!n.error.empty()) {
    dcp_set_error(n.error);
    return DCP_ERR_STATE;
  }
  return DCP_OK;
}

#define DCP_NCCL(call)                                                                                          \
  do {                                                                                                          \
    ncclResult_t r__ = (call);                                                                                  \
    if (r__ != ncclSuccess) {                                                                                   \
      dcp_set_error(std::string(#call) + ": " + nccl().GetErrorString(r__) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
      return DCP_ERR_CUDA;                                                                                      \
    }                                                                                                           \
  } while (0)

__global__ void pack_kernel(long long n, const int* __restrict__ idx, const double* __restrict__ x, double* __restrict__ buf) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) buf[i] = x[idx[i]];
}
__global__ void unpack_kernel(long long n, const int* __restrict__ idx, const double* __restrict__ buf, double* __restrict__ x) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[idx[i]] = buf[i];
}

// sum over several index ranges of x[i] * y[i]: one block per 4096 entries, fixed tree inside the block, partials summed
// by the last stage in block order (bit-reproducible for a given partition)
constexpr int RD_THREADS = 256, RD_BLOCKS = 592;
__global__ void __launch_bounds__(RD_THREADS) range_dot_stage1(int n_ranges, const long long* __restrict__ rb, const long long* __restrict__ re,
                                                               const double* __restrict__ x, const double* __restrict__ y,
                                                               double* __restrict__ partial) {
  __shared__ double sh[RD_THREADS];
  double s = 0.0;
  for (int r = 0; r < n_ranges; ++r)
    for (long long i = rb[r] + blockIdx.x * (long long)RD_THREADS + threadIdx.x; i < re[r]; i += (long long)gridDim.x * RD_THREADS) s += x[i] * y[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int k = RD_THREADS / 2; k > 0; k >>= 1) {
    if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(RD_THREADS) range_dot_stage2(int nb, const double* __restrict__ partial, double* __restrict__ out) {
  __shared__ double sh[RD_THREADS];
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += RD_THREADS) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int k = RD_THREADS / 2; k > 0; k >>= 1) {
    if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

}  // namespace

struct dcp_comm {
  dcp_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  bool owns_comm = true;
  int rank = 0, n_ranks = 1;
  cudaStream_t stream = nullptr;       // communication stream (overlapped products)
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  double* d_scalar = nullptr;          // [8] device scalars for the reductions
  double* d_partial = nullptr;         // [RD_BLOCKS]
  long long* d_ranges = nullptr;       // [2][16]
  double* h_scalar = nullptr;          // pinned [8]
};

struct dcp_halo {
  dcp_comm* comm = nullptr;
  int64_t n_local = 0, n_send = 0, n_recv = 0;
  std::vector<int64_t> send_counts, recv_counts;
  int32_t *send_idx = nullptr, *recv_idx = nullptr;
  double *send_buf = nullptr, *recv_buf = nullptr;
};

extern "C" {

int dcp_comm_unique_id(void* id_out) {
  if (!id_out) return DCP_ERR_ARG;
  DCP_TRY(nccl_ready());
  static_assert(sizeof(ncclUniqueId) == DCP_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  DCP_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return DCP_OK;
}

static int comm_finish(dcp_comm* c) {
  DCP_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  DCP_CUDA(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
  DCP_CUDA(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
  DCP_CUDA(cudaMalloc((void**)&c->d_scalar, sizeof(double) * 8));
  DCP_CUDA(cudaMalloc((void**)&c->d_partial, sizeof(double) * RD_BLOCKS));
  DCP_CUDA(cudaMalloc((void**)&c->d_ranges, sizeof(long long) * 32));
  DCP_CUDA(cudaMallocHost((void**)&c->h_scalar, sizeof(double) * 8));
  return DCP_OK;
}

int dcp_comm_destroy(dcp_comm* c) {
  if (!c) return DCP_OK;
  cudaSetDevice(c->ctx->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->comm && c->owns_comm && nccl().CommDestroy) nccl().CommDestroy(c->comm);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  cudaFree(c->d_scalar);
  cudaFree(c->d_partial);
  cudaFree(c->d_ranges);
  if (c->h_scalar) cudaFreeHost(c->h_scalar);
  delete c;
  return DCP_OK;
}

int dcp_comm_create(dcp_ctx* ctx, const void* id, int rank, int n_ranks, dcp_comm** out) {
  if (!ctx || !id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return DCP_ERR_ARG;
  *out = nullptr;
  DCP_TRY(nccl_ready());
  DCP_CUDA(cudaSetDevice(ctx->device));
  dcp_comm* c = new dcp_comm;
  c->ctx = ctx;
  c->rank = rank;
  c->n_ranks = n_ranks;
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  ncclResult_t r = nccl().CommInitRank(&c->comm, n_ranks, uid, rank);
  if (r != ncclSuccess) {
    dcp_set_error(std::string("ncclCommInitRank: ") + nccl().GetErrorString(r));
    c->comm = nullptr;
    dcp_comm_destroy(c);
    return DCP_ERR_CUDA;
  }
  const int rc = comm_finish(c);
  if (rc != DCP_OK) {
    dcp_comm_destroy(c);
    return rc;
  }
  *out = c;
  return DCP_OK;
}

int dcp_comm_adopt(dcp_ctx* ctx, void* nccl_comm, int rank, int n_ranks, dcp_comm** out) {
  if (!ctx || !nccl_comm || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return DCP_ERR_ARG;
  *out = nullptr;
  DCP_TRY(nccl_ready());
  DCP_CUDA(cudaSetDevice(ctx->device));
  dcp_comm* c = new dcp_comm;
  c->ctx = ctx;
  c->comm = (ncclComm_t)nccl_comm;
  c->owns_comm = false;
  c->rank = rank;
  c->n_ranks = n_ranks;
  const int rc = comm_finish(c);
  if (rc != DCP_OK) {
    dcp_comm_destroy(c);
    return rc;
  }
  *out = c;
  return DCP_OK;
}

int dcp_halo_destroy(dcp_halo* h) {
  if (!h) return DCP_OK;
  cudaSetDevice(h->comm->ctx->device);
  cudaStreamSynchronize(h->comm->stream);
  cudaStreamSynchronize(h->comm->ctx->stream);
  cudaFree(h->send_idx);
  cudaFree(h->recv_idx);
  cudaFree(h->send_buf);
  cudaFree(h->recv_buf);
  delete h;
  return DCP_OK;
}

int dcp_halo_create(dcp_comm* c, int64_t n_local, const int32_t* send_idx, const int64_t* send_counts, const int32_t* recv_idx,
                    const int64_t* recv_counts, dcp_halo** out) {
  if (!c || !out || !send_counts || !recv_counts || n_local < 0) return DCP_ERR_ARG;
  *out = nullptr;
  dcp_ctx* ctx = c->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  dcp_halo* h = new dcp_halo;
  h->comm = c;
  h->n_local = n_local;
  h->send_counts.assign(send_counts, send_counts + c->n_ranks);
  h->recv_counts.assign(recv_counts, recv_counts + c->n_ranks);
  for (int p = 0; p < c->n_ranks; ++p) {
    if (send_counts[p] < 0 || recv_counts[p] < 0 || (p == c->rank && (send_counts[p] || recv_counts[p]))) {
      delete h;
      dcp_set_error("dcp_halo_create: negative count or an exchange with the own rank");
      return DCP_ERR_ARG;
    }
    h->n_send += send_counts[p];
    h->n_recv += recv_counts[p];
  }
  if ((h->n_send && !send_idx) || (h->n_recv && !recv_idx)) {
    delete h;
    return DCP_ERR_ARG;
  }
  for (int64_t i = 0; i < h->n_send; ++i)
    if (send_idx[i] < 0 || send_idx[i] >= n_local) {
      delete h;
      dcp_set_error("dcp_halo_create: send index outside the local vector");
      return DCP_ERR_ARG;
    }
  for (int64_t i = 0; i < h->n_recv; ++i)
    if (recv_idx[i] < 0 || recv_idx[i] >= n_local) {
      delete h;
      dcp_set_error("dcp_halo_create: receive index outside the local vector");
      return DCP_ERR_ARG;
    }
  int rc = DCP_OK;
  if (h->n_send) rc = dcp_upload(ctx, &h->send_idx, send_idx, h->n_send);
  if (rc == DCP_OK && h->n_recv) rc = dcp_upload(ctx, &h->recv_idx, recv_idx, h->n_recv);
  if (rc == DCP_OK && cudaMalloc((void**)&h->send_buf, sizeof(double) * (size_t)std::max<int64_t>(h->n_send, 1)) != cudaSuccess) rc = DCP_ERR_CUDA;
  if (rc == DCP_OK && cudaMalloc((void**)&h->recv_buf, sizeof(double) * (size_t)std::max<int64_t>(h->n_recv, 1)) != cudaSuccess) rc = DCP_ERR_CUDA;
  if (rc == DCP_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = DCP_ERR_CUDA;
  if (rc != DCP_OK) {
    cudaGetLastError();
    dcp_set_error("dcp_halo_create: device allocation / copy failed");
    dcp_halo_destroy(h);
    return rc;
  }
  *out = h;
  return DCP_OK;
}

// pack -> grouped send/recv -> unpack, all enqueued on `s`
static int exchange_on(dcp_halo* h, double* x, cudaStream_t s) {
  dcp_comm* c = h->comm;
  dcp_ctx* ctx = c->ctx;
  if (c->n_ranks == 1) return DCP_OK;
  auto grid = [&](int64_t n) { return (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8); };
  if (h->n_send) {
    pack_kernel<<<grid(h->n_send), 256, 0, s>>>(h->n_send, h->send_idx, x, h->send_buf);
    ctx->launches++;
  }
  DCP_NCCL(nccl().GroupStart());
  int64_t so = 0, ro = 0;
  for (int p = 0; p < c->n_ranks; ++p) {
    if (h->send_counts[p]) DCP_NCCL(nccl().Send(h->send_buf + so, (size_t)h->send_counts[p], ncclDouble, p, c->comm, s));
    so += h->send_counts[p];
  }
  for (int p = 0; p < c->n_ranks; ++p) {
    if (h->recv_counts[p]) DCP_NCCL(nccl().Recv(h->recv_buf + ro, (size_t)h->recv_counts[p], ncclDouble, p, c->comm, s));
    ro += h->recv_counts[p];
  }
  DCP_NCCL(nccl().GroupEnd());
  if (h->n_recv) {
    unpack_kernel<<<grid(h->n_recv), 256, 0, s>>>(h->n_recv, h->recv_idx, h->recv_buf, x);
    ctx->launches++;
  }
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

int dcp_halo_exchange(dcp_halo* h, double* x_dev) {
  if (!h || !x_dev) return DCP_ERR_ARG;
  DCP_CUDA(cudaSetDevice(h->comm->ctx->device));
  return exchange_on(h, x_dev, h->comm->ctx->stream);
}

int dcp_halo_block_vmult(dcp_model* m, int which, dcp_halo* h, double* dst_dev, double* src_dev, int overlap) {
  if (!m || !h || !dst_dev || !src_dev) return DCP_ERR_ARG;
  dcp_ctx* ctx = m->ctx;
  dcp_comm* c = h->comm;
  if (c->ctx != ctx) {
    dcp_set_error("dcp_halo_block_vmult: the halo belongs to another context");
    return DCP_ERR_ARG;
  }
  DCP_CUDA(cudaSetDevice(ctx->device));
  if (!overlap || c->n_ranks == 1) {
    DCP_TRY(exchange_on(h, src_dev, ctx->stream));
    return dcp_block_vmult(m, which, dst_dev, src_dev, DCP_DEVICE);
  }
  // src is final on the main stream from here on; the exchange only writes ghost slots, which the interior rows
  // never read
  DCP_CUDA(cudaEventRecord(c->ev_ready, ctx->stream));
  DCP_CUDA(cudaStreamWaitEvent(c->stream, c->ev_ready, 0));
  DCP_TRY(exchange_on(h, src_dev, c->stream));
  DCP_CUDA(cudaEventRecord(c->ev_done, c->stream));
  DCP_TRY(dcp_block_vmult_rows(m, which, dst_dev, src_dev, DCP_ROWS_INTERIOR));
  DCP_CUDA(cudaStreamWaitEvent(ctx->stream, c->ev_done, 0));
  return dcp_block_vmult_rows(m, which, dst_dev, src_dev, DCP_ROWS_GHOSTED);
}

int dcp_vec_dot_allreduce(dcp_comm* c, int n_ranges, const int64_t* range_begin, const int64_t* range_end, const double* x_dev,
                          const double* y_dev, double* result_host) {
  if (!c || n_ranges < 1 || n_ranges > 16 || !range_begin || !range_end || !x_dev || !y_dev || !result_host) return DCP_ERR_ARG;
  dcp_ctx* ctx = c->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  long long r[32];
  for (int i = 0; i < n_ranges; ++i) {
    if (range_begin[i] < 0 || range_end[i] < range_begin[i]) return DCP_ERR_ARG;
    r[i] = range_begin[i];
    r[16 + i] = range_end[i];
  }
  DCP_CUDA(cudaMemcpyAsync(c->d_ranges, r, sizeof(r), cudaMemcpyHostToDevice, ctx->stream));
  range_dot_stage1<<<RD_BLOCKS, RD_THREADS, 0, ctx->stream>>>(n_ranges, c->d_ranges, c->d_ranges + 16, x_dev, y_dev, c->d_partial);
  range_dot_stage2<<<1, RD_THREADS, 0, ctx->stream>>>(RD_BLOCKS, c->d_partial, c->d_scalar);
  ctx->launches += 2;
  if (c->n_ranks > 1) DCP_NCCL(nccl().AllReduce(c->d_scalar, c->d_scalar, 1, ncclDouble, ncclSum, c->comm, ctx->stream));
  DCP_CUDA(cudaMemcpyAsync(c->h_scalar, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  *result_host = c->h_scalar[0];
  return DCP_OK;
}

int dcp_allreduce_max(dcp_comm* c, int n, double* values_host) {
  if (!c || n < 1 || n > 8 || !values_host) return DCP_ERR_ARG;
  dcp_ctx* ctx = c->ctx;
  DCP_CUDA(cudaSetDevice(ctx->device));
  if (c->n_ranks == 1) return DCP_OK;
  std::memcpy(c->h_scalar, values_host, sizeof(double) * n);
  DCP_CUDA(cudaMemcpyAsync(c->d_scalar, c->h_scalar, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  DCP_NCCL(nccl().AllReduce(c->d_scalar, c->d_scalar, (size_t)n, ncclDouble, ncclMax, c->comm, ctx->stream));
  DCP_CUDA(cudaMemcpyAsync(c->h_scalar, c->d_scalar, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  std::memcpy(values_host, c->h_scalar, sizeof(double) * n);
  return DCP_OK;
}

int dcp_comm_info(const dcp_comm* c, int* rank, int* n_ranks) {
  if (!c) return DCP_ERR_ARG;
  if (rank) *rank = c->rank;
  if (n_ranks) *n_ranks = c->n_ranks;
  return DCP_OK;
}

}  // extern "C"
