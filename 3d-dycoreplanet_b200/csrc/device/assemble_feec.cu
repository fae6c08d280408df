// FEEC Navier-Stokes system / preconditioner assembly: FESystem(FE_Nedelec(0), FE_RaviartThomas(0), FE_DGQ(0)).
//
// Replaces ExteriorCalculus::BoussinesqModel<3>::local_assemble_nse_system + copier
// (/root/reference/include/core/boussineq_model_FEEC.tpp:669-822) and local_assemble_nse_preconditioner +
// copier (:509-584).  19 dofs per cell: 12 line dofs (vorticity w, covariant map J^{-T} phi_hat, curl = J c_hat/det),
// 6 face dofs (velocity u, contravariant Piola J phi_hat/det, div = d_hat/det, times the per-cell face sign of
// source/base/utilities.cc:20-46) and one cell dof (pressure, 1).
//
// One warp per cell.  Lane q maps all shape functions at quadrature point q (coalesced reads of the mapping
// record, reference tables through the read-only path) and parks the 96 mapped values in a shared-memory row
// S[q][.] (odd stride: conflict-free both for the writer "lane = q" and the readers "lane = matrix entry").
// Then lanes own the 177 distinct entries of the block-structured local matrix
//   L[w_i,w_j] = sum w phi_w_i.phi_w_j            L[w_i,u_j] = -sum w curl_w_i.u_j
//   L[u_i,u_j] = sum w u_i.u_j                    L[u_j,w_i] = +dt/Re sum w u_j.curl_w_i   (:753-769)
//   L[u_i,p] = L[p,u_i] = -sum w div u_i
// and the scatter is the general AffineConstraints path of scatter.cuh.
// Preconditioner (quirk Q5, :558-569): JxW multiplies only p*p; dt/Re curl.curl and the sign(u.w) terms are
// summed unweighted over the 8 points.
#include "scatter.cuh"

namespace {

using namespace dcpdev;

constexpr int NW = 12, NU = 6, ND = 19;
constexpr int SV = 97;    // values per quadrature point (96) padded to an odd stride
constexpr int NQMAX = 27;
constexpr int FWARPS = 2;
// offsets inside S[q][.]
constexpr int O_W = 0, O_C = 36, O_U = 72, O_D = 90;

struct FeecArgs {
  long long n_cells;
  int nq, ndt, gstride;
  const double* geom;
  const double* sign;
  const int* l2g;
  const int* l2g_t;
  const double *tw, *tc, *tu, *td;  // reference tables on this rule
  const double* phi_t;
  const double* old_nse;
  const double* old_temp;
  double* rhs;
  dcp_params prm;
};

struct FeecScratch {
  double S[NQMAX * SV];
  double L[ND * ND];
  double l[ND];
  double F[NQMAX * 4];  // rhs integrand per point: A[3] (dotted with u_i) and B (times div u_i), JxW folded in
  double wq[NQMAX];
  double U[ND];
  double sg[ND];
  int idx[ND + 1];
  int lines[ND + 1];
};

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

template <bool SYSTEM>
__global__ void __launch_bounds__(32 * FWARPS) feec_kernel(FeecArgs a, CsView cs, BlockView A, int* err) {
  extern __shared__ __align__(16) unsigned char raw_smem[];
  FeecScratch* all = reinterpret_cast<FeecScratch*>(raw_smem);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  FeecScratch& s = all[wid];
  const int nq = a.nq;
  const double nu = a.prm.dt * a.prm.inv_re;
  for (long long cell = (long long)blockIdx.x * FWARPS + wid; cell < a.n_cells; cell += (long long)gridDim.x * FWARPS) {
    const double* g = a.geom + cell * a.gstride;
    if (lane < ND) {
      const int gi = a.l2g[cell * ND + lane];
      s.idx[lane] = gi;
      s.sg[lane] = a.sign[cell * ND + lane];
      if (SYSTEM) s.U[lane] = a.old_nse[gi];
    }
    for (int i = lane; i < ND * ND; i += 32) s.L[i] = 0.0;
    __syncwarp();
    if (lane < nq) {
      const int q = lane;
      double J[3][3], K[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          K[i][j] = g[nq * (1 + i * 3 + j) + q];
          J[i][j] = g[nq * (13 + i * 3 + j) + q];
        }
      const double det = g[nq * 22 + q], idet = 1.0 / det, w = g[q];
      double* row = s.S + q * SV;
      double ow[3] = {0, 0, 0}, ou[3] = {0, 0, 0};
      for (int k = 0; k < NW; ++k) {
        const double* ph = a.tw + ((size_t)q * NW + k) * 3;
        const double* ch = a.tc + ((size_t)q * NW + k) * 3;
        const double p0 = __ldg(ph), p1 = __ldg(ph + 1), p2 = __ldg(ph + 2);
        const double c0 = __ldg(ch), c1 = __ldg(ch + 1), c2 = __ldg(ch + 2);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double vw = K[0][d] * p0 + K[1][d] * p1 + K[2][d] * p2;
          row[O_W + k * 3 + d] = vw;
          row[O_C + k * 3 + d] = (J[d][0] * c0 + J[d][1] * c1 + J[d][2] * c2) / det;
          if (SYSTEM) ow[d] += s.U[k] * vw;
        }
      }
      for (int k = 0; k < NU; ++k) {
        const double* ph = a.tu + ((size_t)q * NU + k) * 3;
        const double p0 = __ldg(ph), p1 = __ldg(ph + 1), p2 = __ldg(ph + 2);
        const double sgk = s.sg[NW + k];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double ru = (J[d][0] * p0 + J[d][1] * p1 + J[d][2] * p2) / det;
          row[O_U + k * 3 + d] = sgk * ru;
          if (SYSTEM) ou[d] += s.U[NW + k] * ru;  // get_function_values: no face sign (:705-708)
        }
        row[O_D + k] = sgk * __ldg(a.td + k) / det;
      }
      (void)idet;
      s.wq[q] = w;
      if (SYSTEM) {
        double T = 0.0;
        for (int k = 0; k < a.ndt; ++k) T += a.old_temp[a.l2g_t[cell * a.ndt + k]] * __ldg(a.phi_t + q * a.ndt + k);
        const double rho = 1.0 - a.prm.beta * (T - a.prm.T_ref);
        double x[3] = {g[nq * 10 + q], g[nq * 11 + q], g[nq * 12 + q]}, grav[3];
        if (a.prm.cuboid) {
          grav[0] = grav[1] = 0.0;
          grav[2] = -a.prm.g_const;
        } else {
          const double r = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
          const double sc = r > 1.0 ? r : sqrt(r);
          for (int d = 0; d < 3; ++d) grav[d] = -a.prm.g_const * x[d] / sc;
        }
        const double cz = a.prm.cuboid ? a.prm.cor_scale * a.prm.omega : 0.0;
        const double wxu[3] = {ow[1] * ou[2] - ow[2] * ou[1], ow[2] * ou[0] - ow[0] * ou[2], ow[0] * ou[1] - ow[1] * ou[0]};
        const double cxu[3] = {-cz * ou[1], cz * ou[0], 0.0};
        const double dt = a.prm.dt;
#pragma unroll
        for (int d = 0; d < 3; ++d)
          s.F[q * 4 + d] = (ou[d] + dt * rho * (a.prm.g_scale * grav[d]) - dt * wxu[d] - dt * 2.0 * cxu[d]) * w;
        s.F[q * 4 + 3] = -dt * 0.5 * dot3(ou, ou) * w;
      }
    }
    __syncwarp();
    if (SYSTEM) {
      // 78 (w,w) + 72 (w,u) + 21 (u,u) + 6 (u,p) distinct entries
      for (int e = lane; e < 177; e += 32) {
        double acc = 0.0;
        if (e < 78) {
          int i = 0, r = e;
          while (r >= NW - i) { r -= NW - i; ++i; }
          const int j = i + r;
          for (int q = 0; q < nq; ++q) acc += s.wq[q] * dot3(s.S + q * SV + O_W + i * 3, s.S + q * SV + O_W + j * 3);
          s.L[i * ND + j] = acc;
          s.L[j * ND + i] = acc;
        } else if (e < 150) {
          const int i = (e - 78) / NU, j = (e - 78) % NU;
          for (int q = 0; q < nq; ++q) acc += s.wq[q] * dot3(s.S + q * SV + O_C + i * 3, s.S + q * SV + O_U + j * 3);
          s.L[i * ND + NW + j] = -acc;
          s.L[(NW + j) * ND + i] = nu * acc;
        } else if (e < 171) {
          int i = 0, r = e - 150;
          while (r >= NU - i) { r -= NU - i; ++i; }
          const int j = i + r;
          for (int q = 0; q < nq; ++q) acc += s.wq[q] * dot3(s.S + q * SV + O_U + i * 3, s.S + q * SV + O_U + j * 3);
          s.L[(NW + i) * ND + NW + j] = acc;
          s.L[(NW + j) * ND + NW + i] = acc;
        } else {
          const int i = e - 171;
          for (int q = 0; q < nq; ++q) acc += s.wq[q] * s.S[q * SV + O_D + i];
          s.L[(NW + i) * ND + NW + NU] = -acc;
          s.L[(NW + NU) * ND + NW + i] = -acc;
        }
      }
      if (lane < ND) {
        double acc = 0.0;
        if (lane >= NW && lane < NW + NU) {
          const int i = lane - NW;
          for (int q = 0; q < nq; ++q)
            acc += dot3(s.S + q * SV + O_U + i * 3, s.F + q * 4) + s.S[q * SV + O_D + i] * s.F[q * 4 + 3];
        }
        s.l[lane] = acc;
      }
    } else {
      // preconditioner: 78 curl-curl, 72 sign terms, 1 pressure mass
      for (int e = lane; e < 151; e += 32) {
        double acc = 0.0;
        if (e < 78) {
          int i = 0, r = e;
          while (r >= NW - i) { r -= NW - i; ++i; }
          const int j = i + r;
          for (int q = 0; q < nq; ++q) acc += dot3(s.S + q * SV + O_C + i * 3, s.S + q * SV + O_C + j * 3);
          s.L[i * ND + j] = nu * acc;
          s.L[j * ND + i] = nu * acc;
        } else if (e < 150) {
          const int i = (e - 78) / NW, j = (e - 78) % NW;  // u_i, w_j
          for (int q = 0; q < nq; ++q) {
            const double x = dot3(s.S + q * SV + O_U + i * 3, s.S + q * SV + O_W + j * 3);
            acc += fabs(x) > 1.0e-9 ? (signbit(x) ? -1.0 : 1.0) : 0.0;
          }
          s.L[(NW + i) * ND + j] = acc;
          s.L[j * ND + NW + i] = acc;
        } else {
          for (int q = 0; q < nq; ++q) acc += s.wq[q];
          s.L[(NW + NU) * ND + NW + NU] = acc;
        }
      }
    }
    __syncwarp();
    distribute_local_matrix<true>(cs, ND, ND, s.L, SYSTEM ? s.l : nullptr, s.idx, s.lines, A, SYSTEM ? a.rhs : nullptr,
                                  lane, 32, false, err);
    __syncwarp();
  }
}

}  // namespace

int dcp_launch_feec(dcp_model* m, const dcp_params& p, bool system, const double* old_nse, const double* old_temp) {
  dcp_ctx* ctx = m->ctx;
  FeecArgs a{};
  a.n_cells = m->n_cells;
  a.nq = system ? m->nq_nse : m->nq_pre;
  a.ndt = m->ndt;
  a.gstride = system ? m->gs_n : m->gs_p;
  a.geom = system ? m->geom_qn : m->geom_qp;
  a.sign = m->nse_sign;
  a.l2g = m->nse_l2g;
  a.l2g_t = m->temp_l2g;
  a.tw = system ? m->feec_w_qn : m->feec_w_qp;
  a.tc = system ? m->feec_c_qn : m->feec_c_qp;
  a.tu = system ? m->feec_u_qn : m->feec_u_qp;
  a.td = m->feec_div;
  a.phi_t = m->phi_t_qn;
  a.old_nse = old_nse;
  a.old_temp = old_temp;
  a.rhs = system ? m->nse_rhs : nullptr;
  a.prm = p;
  if (a.nq > NQMAX) {
    dcp_set_error("FEEC: more than 27 quadrature points per cell is not supported");
    return DCP_ERR_ARG;
  }
  if (m->n_cells == 0) return DCP_OK;
  const size_t smem = sizeof(FeecScratch) * FWARPS;
  static bool attr_set = false;
  if (!attr_set) {
    DCP_CUDA(cudaFuncSetAttribute(feec_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DCP_CUDA(cudaFuncSetAttribute(feec_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  long long blocks = (m->n_cells + FWARPS - 1) / FWARPS;
  const long long cap = (long long)ctx->sm_count * 4;
  if (blocks > cap) blocks = cap;
  const BlockMat& mat = system ? m->nse : m->pre;
  if (system)
    feec_kernel<true><<<(unsigned)blocks, 32 * FWARPS, smem, ctx->stream>>>(a, make_view(m->nse_cs), make_view(mat), ctx->d_err);
  else
    feec_kernel<false><<<(unsigned)blocks, 32 * FWARPS, smem, ctx->stream>>>(a, make_view(m->nse_cs), make_view(mat), ctx->d_err);
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}
