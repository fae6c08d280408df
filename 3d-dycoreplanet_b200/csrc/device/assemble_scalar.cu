// Temperature mass / stiffness matrices and right-hand side (scalar Q1 or Q2 Lagrange space).
//
// Replaces local_assemble_temperature_matrix + copier (/root/reference/include/core/boussinesq_model.tpp:
// 748-817) and local_assemble_temperature_rhs + copier (:873-964); identical integrands are used by the
// FEEC model (include/core/boussineq_model_FEEC.tpp:883-952, 1008-1099).
//   M[i,j] += phi_i phi_j JxW,   K[i,j] += (1/Pe) grad phi_i . grad phi_j JxW      (:786-797)
//   r[i]   += (phi_i T_old - tau phi_i (u_new . grad T_old) - tau*0*phi_i) JxW,  tau = dt/n   (:928-937)
//   matrix_for_bc(j,i) = (phi_i phi_j + tau/Pe grad phi_i . grad phi_j) JxW  for inhomogeneous i (:939-949)
// One warp per cell: 8..27 dofs per cell make these kernels HBM/latency bound, not FLOP bound.
#include "scatter.cuh"

namespace {

using namespace dcpdev;

constexpr int MAX_ND = 27;
constexpr int WARPS = 3;

struct ScalarArgs {
  long long n_cells;
  int nd, nq;
  int gstride;          // doubles per cell record
  int feec;             // rhs: velocity is the Raviart-Thomas field (Piola-mapped, unsigned)
  const double* geom;
  const int* l2g;
  const double* phi;   // [nq][nd]
  const double* dphi;  // [nq][nd][dim]
  dcp_params prm;
  // rhs only
  const int* l2g_nse;
  int nd_nse, ndu;
  const int* nse_field;
  const int* nse_base;
  const double* phi_u;  // [nq][ndu] velocity base element on the temperature rule
  const double* old_temp;
  const double* nse_solution;
  double* rhs;
};

struct WarpScratch {
  double A[MAX_ND * MAX_ND];
  double B[MAX_ND * MAX_ND];
  double g[MAX_ND * 3];
  double ph[MAX_ND];
  double l[MAX_ND];
  double T[MAX_ND];
  int idx[MAX_ND + 1];
  int lines[MAX_ND + 1];
};

template <int DIM>
__device__ __forceinline__ void point_shapes(const ScalarArgs& a, const double* g, int q, int lane, WarpScratch& s) {
  if (lane < a.nd) {
    const double* dr = a.dphi + ((size_t)q * a.nd + lane) * DIM;
    double r[DIM];
#pragma unroll
    for (int e = 0; e < DIM; ++e) r[e] = __ldg(dr + e);
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double v = 0.0;
#pragma unroll
      for (int e = 0; e < DIM; ++e) v += __ldg(g + a.nq * (1 + e * DIM + d) + q) * r[e];
      s.g[lane * 3 + d] = v;
    }
    s.ph[lane] = __ldg(a.phi + (size_t)q * a.nd + lane);
  }
}

template <int DIM>
__global__ void __launch_bounds__(32 * WARPS) temperature_matrix_kernel(ScalarArgs a, CsView cs, BlockView Mass,
                                                                        BlockView Stiff, int* err) {
  __shared__ WarpScratch scratch[WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  WarpScratch& s = scratch[wid];
  const int nd = a.nd, nn = nd * nd;
  const int gs = a.gstride;
  for (long long cell = (long long)blockIdx.x * WARPS + wid; cell < a.n_cells; cell += (long long)gridDim.x * WARPS) {
    const double* g = a.geom + cell * gs;
    for (int i = lane; i < nn; i += 32) {
      s.A[i] = 0.0;
      s.B[i] = 0.0;
    }
    if (lane < nd) s.idx[lane] = a.l2g[cell * nd + lane];
    __syncwarp();
    for (int q = 0; q < a.nq; ++q) {
      point_shapes<DIM>(a, g, q, lane, s);
      __syncwarp();
      const double w = __ldg(g + q);
      for (int e = lane; e < nn; e += 32) {
        const int i = e / nd, j = e - i * nd;
        double gg = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) gg += s.g[i * 3 + d] * s.g[j * 3 + d];
        s.A[e] += s.ph[i] * s.ph[j] * w;
        s.B[e] += a.prm.inv_pe * gg * w;
      }
      __syncwarp();
    }
    distribute_local_matrix<true>(cs, nd, nd, s.A, nullptr, s.idx, s.lines, Mass, nullptr, lane, 32, false, err);
    __syncwarp();
    distribute_local_matrix<true>(cs, nd, nd, s.B, nullptr, s.idx, s.lines, Stiff, nullptr, lane, 32, false, err);
    __syncwarp();
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int DIM>
__global__ void __launch_bounds__(32 * WARPS) temperature_rhs_kernel(ScalarArgs a, CsView cs) {
  __shared__ WarpScratch scratch[WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  WarpScratch& s = scratch[wid];
  const int nd = a.nd, nn = nd * nd;
  const int gs = a.gstride;
  const double tau = a.prm.dt / a.prm.nse_interval;
  for (long long cell = (long long)blockIdx.x * WARPS + wid; cell < a.n_cells; cell += (long long)gridDim.x * WARPS) {
    const double* g = a.geom + cell * gs;
    bool inhom = false;
    if (lane < nd) {
      const int gi = a.l2g[cell * nd + lane];
      s.idx[lane] = gi;
      s.T[lane] = a.old_temp[gi];
      s.l[lane] = 0.0;
      const int li = cs.line_of_dof[gi];
      s.lines[lane] = li;
      inhom = li >= 0 && cs.inhom[li] != 0.0;
    }
    const bool need_bc = __any_sync(0xffffffffu, inhom);
    if (need_bc)
      for (int i = lane; i < nn; i += 32) s.A[i] = 0.0;
    // velocity dofs of this cell held in registers, strided over lanes
    __syncwarp();
    for (int q = 0; q < a.nq; ++q) {
      point_shapes<DIM>(a, g, q, lane, s);
      __syncwarp();
      double pT = 0.0, pg[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) pg[d] = 0.0;
      if (lane < nd) {
        pT = s.T[lane] * s.ph[lane];
#pragma unroll
        for (int d = 0; d < DIM; ++d) pg[d] = s.T[lane] * s.g[lane * 3 + d];
      }
      double pu[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) pu[d] = 0.0;
      if (a.feec) {
        // u(q) = sum_k U_k J phi_hat_k / det J over the six face dofs (cell dofs 12..17), no face sign
        // (get_function_values, boussineq_model_FEEC.tpp:1039-1040)
        if (DIM == 3 && lane < 6) {
          const double U = a.nse_solution[a.l2g_nse[cell * a.nd_nse + 12 + lane]];
          const double det = __ldg(g + a.nq * 22 + q);
          const double* ph = a.phi_u + ((size_t)q * 6 + lane) * 3;
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            double v = 0.0;
#pragma unroll
            for (int e = 0; e < DIM; ++e) v += __ldg(g + a.nq * (13 + d * 3 + e) + q) * __ldg(ph + e);
            pu[d] = U * v / det;
          }
        }
      } else
      for (int k = lane; k < a.nd_nse; k += 32) {
        const int f = __ldg(a.nse_field + k);
        if (f < DIM) {
          const double v = a.nse_solution[a.l2g_nse[cell * a.nd_nse + k]] * __ldg(a.phi_u + (size_t)q * a.ndu + __ldg(a.nse_base + k));
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            if (d == f) pu[d] += v;
        }
      }
      const double oldT = warp_sum(pT);
      double ugT = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) ugT += warp_sum(pu[d]) * warp_sum(pg[d]);
      const double w = __ldg(g + q);
      const double gamma = 0.0;  // heat source multiplied by literal 0 in the reference (:922-926)
      if (lane < nd) s.l[lane] += (s.ph[lane] * oldT - tau * s.ph[lane] * ugT - tau * gamma * s.ph[lane]) * w;
      if (need_bc)
        for (int e = lane; e < nn; e += 32) {
          const int j = e / nd, i = e - j * nd;
          double gg = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) gg += s.g[i * 3 + d] * s.g[j * 3 + d];
          s.A[e] += (s.ph[i] * s.ph[j] + tau * a.prm.inv_pe * gg) * w;
        }
      __syncwarp();
    }
    distribute_local_vector_bc<true>(cs, nd, nd, s.l, s.A, s.idx, s.lines, a.rhs, lane, 32);
    __syncwarp();
  }
}

unsigned scalar_grid(dcp_ctx* ctx, long long n_cells) {
  long long b = (n_cells + WARPS - 1) / WARPS;
  long long cap = (long long)ctx->sm_count * 8;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int dcp_launch_temperature_matrix(dcp_model* m, const dcp_params& p) {
  dcp_ctx* ctx = m->ctx;
  if (m->temp_n_local > MAX_ND) {
    dcp_set_error("temperature space: more than 27 dofs per cell is not supported");
    return DCP_ERR_ARG;
  }
  ScalarArgs a{};
  a.n_cells = m->n_cells;
  a.nd = m->temp_n_local;
  a.nq = m->nq_temp;
  a.geom = m->geom_qt;
  a.gstride = m->gs_t;
  a.l2g = m->temp_l2g;
  a.phi = m->phi_t_qt;
  a.dphi = m->dphi_t_qt;
  a.prm = p;
  if (m->n_cells == 0) return DCP_OK;
  if (m->dim == 3)
    temperature_matrix_kernel<3><<<scalar_grid(ctx, m->n_cells), 32 * WARPS, 0, ctx->stream>>>(
        a, make_view(m->temp_cs), make_view(m->tmass), make_view(m->tstiff), ctx->d_err);
  else
    temperature_matrix_kernel<2><<<scalar_grid(ctx, m->n_cells), 32 * WARPS, 0, ctx->stream>>>(
        a, make_view(m->temp_cs), make_view(m->tmass), make_view(m->tstiff), ctx->d_err);
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

int dcp_launch_temperature_rhs(dcp_model* m, const dcp_params& p, const double* old_temp, const double* nse_solution) {
  dcp_ctx* ctx = m->ctx;
  ScalarArgs a{};
  a.n_cells = m->n_cells;
  a.nd = m->temp_n_local;
  a.nq = m->nq_temp;
  a.geom = m->geom_qt;
  a.gstride = m->gs_t;
  a.feec = m->family == DCP_FAMILY_FEEC ? 1 : 0;
  a.l2g = m->temp_l2g;
  a.phi = m->phi_t_qt;
  a.dphi = m->dphi_t_qt;
  a.prm = p;
  a.l2g_nse = m->nse_l2g;
  a.nd_nse = m->nse_n_local;
  a.ndu = m->ndu;
  a.nse_field = m->nse_local_field;
  a.nse_base = m->nse_local_base;
  a.phi_u = a.feec ? m->feec_u_qt : m->phi_u_qt;
  a.old_temp = old_temp;
  a.nse_solution = nse_solution;
  a.rhs = m->temp_rhs;
  if (m->n_cells == 0) return DCP_OK;
  if (m->dim == 3)
    temperature_rhs_kernel<3><<<scalar_grid(ctx, m->n_cells), 32 * WARPS, 0, ctx->stream>>>(a, make_view(m->temp_cs));
  else
    temperature_rhs_kernel<2><<<scalar_grid(ctx, m->n_cells), 32 * WARPS, 0, ctx->stream>>>(a, make_view(m->temp_cs));
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}
