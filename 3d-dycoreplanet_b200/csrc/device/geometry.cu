// Mapping data producer: evaluates what FEValues::reinit computes for the assemblers --
// Jacobian, inverse Jacobian, JxW and the quadrature point location -- from the support points of the cell mapping,
// on the device, instead of shipping 13 (classic) or 23 (FEEC) doubles per quadrature point from the host
// (4.4 GB at refine 6).  MappingQ(p) as the reference uses it (boussinesq_model.tpp:20; boussineq_model_assembly
// .tpp:25,69-73): cells with boundary lines carry (p+1)^dim support points, interior cells their 2^dim vertices.
//
// One warp per cell, lane = quadrature point (two rounds for the 64-point rule).  The record layout is the
// one dcp_model_desc documents: [JxW | Kinv[e][d] | xq[d] (| J[i][j] | detJ)], nq entries each, so every store is
// contiguous over the lanes.
#include "dcp_internal.cuh"

namespace {

struct MapArgs {
  long long n_cells;
  int nq, extended;
  int n_low, n_high;
  const long long* ptr;
  const double* X;
  const double *N_low, *dN_low, *N_high, *dN_high, *w;
  double* out;
};

template <int DIM>
__global__ void __launch_bounds__(128) mapping_kernel(MapArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nrec = 1 + DIM * DIM + DIM + (a.extended ? DIM * DIM + 1 : 0);
  for (long long c = warp; c < a.n_cells; c += n_warps) {
    const long long p0 = a.ptr[c];
    const int ns = (int)(a.ptr[c + 1] - p0);
    const bool high = ns != a.n_low;
    const double* N = high ? a.N_high : a.N_low;
    const double* dN = high ? a.dN_high : a.dN_low;
    const double* X = a.X + p0 * DIM;
    double* g = a.out + c * (long long)(nrec * a.nq);
    for (int q = lane; q < a.nq; q += 32) {
      double J[DIM * DIM], x[DIM];
#pragma unroll
      for (int e = 0; e < DIM * DIM; ++e) J[e] = 0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) x[d] = 0;
      for (int s = 0; s < ns; ++s) {
        const double n = __ldg(N + (size_t)q * ns + s);
        double dn[DIM], xs[DIM];
#pragma unroll
        for (int j = 0; j < DIM; ++j) dn[j] = __ldg(dN + ((size_t)q * ns + s) * DIM + j);
#pragma unroll
        for (int i = 0; i < DIM; ++i) xs[i] = __ldg(X + s * DIM + i);
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
          x[i] += n * xs[i];
#pragma unroll
          for (int j = 0; j < DIM; ++j) J[i * DIM + j] += xs[i] * dn[j];
        }
      }
      double K[DIM * DIM], det;
      if (DIM == 2) {
        det = J[0] * J[3] - J[1] * J[2];
        const double id = 1.0 / det;
        K[0] = J[3] * id;
        K[1] = -J[1] * id;
        K[2] = -J[2] * id;
        K[3] = J[0] * id;
      } else {
        const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
        det = J[0] * c00 + J[1] * c01 + J[2] * c02;
        const double id = 1.0 / det;
        K[0] = c00 * id;
        K[1] = (J[2] * J[7] - J[1] * J[8]) * id;
        K[2] = (J[1] * J[5] - J[2] * J[4]) * id;
        K[3] = c01 * id;
        K[4] = (J[0] * J[8] - J[2] * J[6]) * id;
        K[5] = (J[2] * J[3] - J[0] * J[5]) * id;
        K[6] = c02 * id;
        K[7] = (J[1] * J[6] - J[0] * J[7]) * id;
        K[8] = (J[0] * J[4] - J[1] * J[3]) * id;
      }
      g[q] = det * __ldg(a.w + q);
#pragma unroll
      for (int e = 0; e < DIM * DIM; ++e) g[(1 + e) * a.nq + q] = K[e];
#pragma unroll
      for (int d = 0; d < DIM; ++d) g[(1 + DIM * DIM + d) * a.nq + q] = x[d];
      if (a.extended) {
        const int o = 1 + DIM * DIM + DIM;
#pragma unroll
        for (int e = 0; e < DIM * DIM; ++e) g[(o + e) * a.nq + q] = J[e];
        g[(o + DIM * DIM) * a.nq + q] = det;
      }
    }
  }
}

}  // namespace

int dcp_geometry_create(dcp_ctx* ctx, const dcp_mapping_desc* d, double** geom_dev) {
  if (!ctx || !d || !geom_dev) return DCP_ERR_ARG;
  *geom_dev = nullptr;
  if ((d->dim != 2 && d->dim != 3) || d->nq <= 0 || d->n_cells < 0 || !d->support_ptr || !d->support_points || !d->N_low ||
      !d->dN_low || !d->weights || d->n_low != (1 << d->dim)) {
    dcp_set_error("dcp_geometry_create: bad mapping description");
    return DCP_ERR_ARG;
  }
  DCP_CUDA(cudaSetDevice(ctx->device));
  const int dim = d->dim;
  const int64_t nc = d->n_cells;
  bool any_high = false;
  for (int64_t c = 0; c < nc; ++c) {
    const int64_t ns = d->support_ptr[c + 1] - d->support_ptr[c];
    if (ns != d->n_low && !(d->N_high && d->dN_high && ns == d->n_high)) {
      dcp_set_error("dcp_geometry_create: a cell has neither n_low nor n_high support points");
      return DCP_ERR_ARG;
    }
    any_high |= ns != d->n_low;
  }
  const int nrec = 1 + dim * dim + dim + (d->extended ? dim * dim + 1 : 0);
  int64_t* ptr = nullptr;
  double *X = nullptr, *Nl = nullptr, *dNl = nullptr, *Nh = nullptr, *dNh = nullptr, *w = nullptr, *out = nullptr;
  int rc = DCP_OK;
  auto cleanup = [&]() {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ptr);
    cudaFree(X);
    cudaFree(Nl);
    cudaFree(dNl);
    cudaFree(Nh);
    cudaFree(dNh);
    cudaFree(w);
  };
#define G_TRY(x)           \
  do {                     \
    rc = (x);              \
    if (rc != DCP_OK) {    \
      cleanup();           \
      cudaFree(out);       \
      return rc;           \
    }                      \
  } while (0)
  G_TRY(dcp_upload(ctx, &ptr, d->support_ptr, nc + 1));
  G_TRY(dcp_upload(ctx, &X, d->support_points, d->support_ptr[nc] * dim));
  G_TRY(dcp_upload(ctx, &Nl, d->N_low, (int64_t)d->nq * d->n_low));
  G_TRY(dcp_upload(ctx, &dNl, d->dN_low, (int64_t)d->nq * d->n_low * dim));
  if (any_high) {
    G_TRY(dcp_upload(ctx, &Nh, d->N_high, (int64_t)d->nq * d->n_high));
    G_TRY(dcp_upload(ctx, &dNh, d->dN_high, (int64_t)d->nq * d->n_high * dim));
  }
  G_TRY(dcp_upload(ctx, &w, d->weights, d->nq));
  if (cudaMalloc((void**)&out, sizeof(double) * (size_t)std::max<int64_t>(1, nc * nrec * d->nq)) != cudaSuccess) {
    dcp_set_error("dcp_geometry_create: out of device memory");
    cleanup();
    return DCP_ERR_CUDA;
  }
  if (nc > 0) {
    MapArgs a{nc, d->nq, d->extended, d->n_low, d->n_high, (const long long*)ptr, X, Nl, dNl, Nh, dNh, w, out};
    const int threads = 128;
    const int blocks = (int)std::min<int64_t>((nc + 3) / 4, (int64_t)ctx->sm_count * 16);
    if (dim == 3)
      mapping_kernel<3><<<blocks, threads, 0, ctx->stream>>>(a);
    else
      mapping_kernel<2><<<blocks, threads, 0, ctx->stream>>>(a);
    ++ctx->launches;
    if (cudaGetLastError() != cudaSuccess) {
      dcp_set_error("dcp_geometry_create: kernel launch failed");
      cleanup();
      cudaFree(out);
      return DCP_ERR_CUDA;
    }
  }
  cleanup();
  *geom_dev = out;
  return DCP_OK;
#undef G_TRY
}
