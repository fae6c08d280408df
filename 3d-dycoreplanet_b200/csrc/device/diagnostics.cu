// Small data-parallel passes next to the solves: the velocity maximum and the CFL number the
// time-step control reads (boussinesq_model.tpp:1023-1101; FEEC boussineq_model_FEEC.tpp:1158-1240) and
// AffineConstraints::distribute applied to a solution vector after a solve (boussinesq_model.tpp:1233, 1442).
// Keeping them on the device avoids a download of the whole solution vector every time step.
#include "dcp_internal.cuh"

namespace {

// non-negative doubles order like their bit patterns
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

template <int DIM>
__device__ double cell_diameter(const double* X) {
  // deal.II TriaAccessor::diameter(): the longest of the space diagonals between opposite vertices
  constexpr int NV = 1 << DIM;
  double best = 0;
#pragma unroll
  for (int v = 0; v < NV / 2; ++v) {
    const int w = NV - 1 - v;
    double s = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const double t = X[w * DIM + d] - X[v * DIM + d];
      s += t * t;
    }
    best = fmax(best, s);
  }
  return sqrt(best);
}

// classic family: QIterated(QTrapez, velocity_degree) points are the Lagrange nodes of the velocity element, so the
// velocity values there are the nodal values themselves.  One thread per cell.
template <int DIM>
__global__ void __launch_bounds__(128) velocity_extrema_classic(long long n_cells, int nd, int ndu, const int32_t* __restrict__ l2g,
                                                                const int32_t* __restrict__ vel_dof,  // [DIM][ndu] cell dof of (component, node)
                                                                const double* __restrict__ x, const double* __restrict__ vertices,
                                                                double* __restrict__ out) {
  double vmax = 0, cmax = 0;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n_cells; c += (long long)gridDim.x * blockDim.x) {
    const int32_t* idx = l2g + c * nd;
    double cell_max = 0;
    for (int n = 0; n < ndu; ++n) {
      double s = 0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        const double u = x[idx[vel_dof[d * ndu + n]]];
        s += u * u;
      }
      cell_max = fmax(cell_max, s);
    }
    cell_max = sqrt(cell_max);
    vmax = fmax(vmax, cell_max);
    if (vertices) cmax = fmax(cmax, fmax(cell_max, 1e-10) / cell_diameter<DIM>(vertices + c * (DIM << DIM)));
  }
  for (int o = 16; o > 0; o >>= 1) {
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_max_nonneg(out, vmax);
    atomic_max_nonneg(out + 1, cmax);
  }
}

// FEEC family: QIterated(QTrapez, 1) = the cell vertices, FEValues with the default Q1 mapping, velocity = the
// Raviart-Thomas component (extractor `dim`), values u = J u_hat / det J without the face signs (quirk: the
// get_function_values path is unsigned).  At vertex v the reference value of the face function (d, side) is
// e_d if side == bit d of v, else 0.
__global__ void __launch_bounds__(128) velocity_extrema_feec(long long n_cells, const int32_t* __restrict__ l2g, const double* __restrict__ x,
                                                             const double* __restrict__ vertices, double* __restrict__ out) {
  double vmax = 0, cmax = 0;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n_cells; c += (long long)gridDim.x * blockDim.x) {
    const int32_t* idx = l2g + c * 19;
    const double* X = vertices + c * 24;
    double U[6];
#pragma unroll
    for (int f = 0; f < 6; ++f) U[f] = x[idx[12 + f]];
    double cell_max = 0;
    for (int v = 0; v < 8; ++v) {
      const int b[3] = {v & 1, (v >> 1) & 1, (v >> 2) & 1};
      double J[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) J[e] = 0;
      for (int s = 0; s < 8; ++s) {
        const int t[3] = {s & 1, (s >> 1) & 1, (s >> 2) & 1};
        // d/dxi_j of the trilinear vertex function s at vertex v
        double g[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double p = t[j] ? 1.0 : -1.0;
#pragma unroll
          for (int k = 0; k < 3; ++k)
            if (k != j) p *= (t[k] == b[k]) ? 1.0 : 0.0;
          g[j] = p;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) J[i * 3 + j] += X[s * 3 + i] * g[j];
      }
      const double det = J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
      const double uh[3] = {U[0 + b[0]], U[2 + b[1]], U[4 + b[2]]};
      double s2 = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double u = (J[i * 3] * uh[0] + J[i * 3 + 1] * uh[1] + J[i * 3 + 2] * uh[2]) / det;
        s2 += u * u;
      }
      cell_max = fmax(cell_max, s2);
    }
    cell_max = sqrt(cell_max);
    vmax = fmax(vmax, cell_max);
    cmax = fmax(cmax, fmax(cell_max, 1e-10) / cell_diameter<3>(X));
  }
  for (int o = 16; o > 0; o >>= 1) {
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_max_nonneg(out, vmax);
    atomic_max_nonneg(out + 1, cmax);
  }
}

// x[line] = sum_k w_k x[master_k] + inhomogeneity.  Masters are never constrained themselves (closed constraints),
// so reads and writes touch disjoint entries.
__global__ void __launch_bounds__(256) distribute_kernel(long long n_dofs, CsView cs, double* __restrict__ x) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_dofs) return;
  const int l = cs.line_of_dof[i];
  if (l < 0) return;
  double v = cs.inhom[l];
  for (int k = cs.line_ptr[l]; k < cs.line_ptr[l + 1]; ++k) v += cs.entry_w[k] * x[cs.entry_dof[k]];
  x[i] = v;
}

}  // namespace

int dcp_launch_velocity_extrema(dcp_model* m, const double* nse_solution, double* out2_dev) {
  dcp_ctx* ctx = m->ctx;
  DCP_CUDA(cudaMemsetAsync(out2_dev, 0, 2 * sizeof(double), ctx->stream));
  const long long nc = m->n_owned_cells > 0 ? m->n_owned_cells : m->n_cells;
  if (nc == 0) return DCP_OK;
  const int threads = 128;
  const int blocks = (int)std::min<long long>((nc + threads - 1) / threads, (long long)ctx->sm_count * 16);
  if (m->family == DCP_FAMILY_FEEC) {
    velocity_extrema_feec<<<blocks, threads, 0, ctx->stream>>>(nc, m->nse_l2g, nse_solution, m->cell_vertices, out2_dev);
  } else if (m->dim == 3) {
    velocity_extrema_classic<3><<<blocks, threads, 0, ctx->stream>>>(nc, m->nse_n_local, m->ndu, m->nse_l2g, m->vel_dof, nse_solution,
                                                                     m->cell_vertices, out2_dev);
  } else {
    velocity_extrema_classic<2><<<blocks, threads, 0, ctx->stream>>>(nc, m->nse_n_local, m->ndu, m->nse_l2g, m->vel_dof, nse_solution,
                                                                     m->cell_vertices, out2_dev);
  }
  ++ctx->launches;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

int dcp_launch_distribute(dcp_ctx* ctx, const DevCs& cs, double* x) {
  if (cs.n_lines == 0 || cs.n_dofs == 0) return DCP_OK;
  CsView v{cs.line_of_dof, cs.line_ptr, cs.entry_dof, cs.entry_w, cs.inhom};
  const int threads = 256;
  distribute_kernel<<<(unsigned)((cs.n_dofs + threads - 1) / threads), threads, 0, ctx->stream>>>(cs.n_dofs, v, x);
  ++ctx->launches;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}
