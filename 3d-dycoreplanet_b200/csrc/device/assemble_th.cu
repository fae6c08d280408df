// Classic Taylor-Hood (Q2^dim x Q1) Navier-Stokes system / preconditioner assembly, cell-parallel
// general path ("ATOMIC" strategy, also the constrained-cell fix-up pass of the row-owner strategy).
//
// Replaces Standard::BoussinesqModel::local_assemble_nse_system + copy_local_to_global_nse_system
// (/root/reference/include/core/boussinesq_model.tpp:550-687) and local_assemble_nse_preconditioner +
// copier (:421-476).  One CTA integrates one cell at a time (persistent grid-stride loop): mapping data
// and reference tables are staged in shared memory, physical gradients are formed once per cell, and the
// 89x89 local matrix is built from its block structure instead of the reference's dense 89*89*27 loop:
//
//   L[(a,c),(b,d)] = delta_cd (m_ab + nu k_ab) + nu g_ab^{dc},   nu = dt/Re            (:626-632)
//     m_ab = sum_q w phi_a phi_b,  g_ab^{dc} = sum_q w d_d phi_a d_c phi_b,  k_ab = sum_d g_ab^{dd}
//     (from 2 eps(phi_a e_c):eps(phi_b e_d) = delta_cd grad phi_a . grad phi_b + d_d phi_a d_c phi_b)
//   L[(a,c),p_b] = L[p_b,(a,c)] = - sum_q w d_c phi_a psi_b                           (:633-635)
//   L[p,p] = 0
//   preconditioner: L[(a,c),(b,c)] = m_ab + nu k_ab,  L[p_a,p_b] = sum_q w psi_a psi_b (:455-462)
//   rhs l[(a,c)] = sum_q w phi_a ( u_c + dt rho g_c - dt (u.grad)u_c - dt C_c )        (:655-669)
//
// The scatter is scatter.cuh (general AffineConstraints semantics, atomics).
#include "scatter.cuh"

namespace {

using namespace dcpdev;

template <int DIM>
struct ThDims {
  static constexpr int NU = DIM == 3 ? 27 : 9;  // Q2 nodes
  static constexpr int NP = DIM == 3 ? 8 : 4;   // Q1 nodes
  static constexpr int NQ = NU;                 // QGauss(3)
  static constexpr int ND = DIM * NU + NP;
  static constexpr int GS = NQ * (1 + DIM * DIM + DIM);
};

struct ThArgs {
  long long n_cells;
  const int* cell_list;  // optional subset
  long long n_list;
  const double* geom;
  const int* l2g;
  const int* l2g_t;
  const int* local_field;
  const int* local_base;
  const double* phi_u;
  const double* dphi_u;
  const double* phi_p;
  const double* phi_t;
  int ndt;
  const double* old_nse;
  const double* old_temp;
  double* rhs;
  int system;            // 1: system matrix (+rhs if rhs != nullptr), 0: preconditioner
  int only_constrained;  // fix-up mode
  int skip_matrix;       // rhs only
  dcp_params prm;
};

template <int DIM>
__device__ __forceinline__ void coeff_gravity(const dcp_params& P, const double* x, double* g) {
  if (P.cuboid) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) g[d] = 0.0;
    g[DIM - 1] = -P.g_const;
    return;
  }
  double r = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) r += x[d] * x[d];
  r = sqrt(r);
  const double s = r > 1.0 ? r : sqrt(r);
#pragma unroll
  for (int d = 0; d < DIM; ++d) g[d] = -P.g_const * x[d] / s;
}

template <int DIM>
__global__ void __launch_bounds__(256) th_cell_kernel(ThArgs a, CsView cs, BlockView A, int* err) {
  using D = ThDims<DIM>;
  constexpr int NU = D::NU, NP = D::NP, NQ = D::NQ, ND = D::ND, GS = D::GS;
  extern __shared__ double smem[];
  double* L = smem;                      // ND*ND
  double* G = L + ND * ND;               // NQ*NU*DIM physical gradients
  double* sphi = G + NQ * NU * DIM;      // NQ*NU
  double* spsi = sphi + NQ * NU;         // NQ*NP
  double* sgeo = spsi + NQ * NP;         // GS
  double* sF = sgeo + GS;                // NQ*DIM   rhs integrand vector
  double* sl = sF + NQ * DIM;            // ND       local rhs
  double* sU = sl + ND;                  // ND       old nse values of the cell
  double* sT = sU + ND;                  // 32       old temperature at q (NQ <= 27) + dofs
  int* sidx = (int*)(sT + 64);           // ND
  int* slines = sidx + ND;               // ND
  int* sys_u = slines + ND;              // DIM*NU   system index of (component, node)
  int* sys_p = sys_u + DIM * NU;         // NP
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int i = tid; i < NQ * NU; i += nt) sphi[i] = a.phi_u[i];
  for (int i = tid; i < NQ * NP; i += nt) spsi[i] = a.phi_p[i];
  for (int i = tid; i < ND; i += nt) {
    const int f = a.local_field[i], bs = a.local_base[i];
    if (f < DIM) sys_u[f * NU + bs] = i; else sys_p[bs] = i;
  }
  __syncthreads();
  const double nu = a.prm.dt * a.prm.inv_re;
  const long long n_work = a.n_list >= 0 ? a.n_list : a.n_cells;  // n_list < 0: every cell

  for (long long w = blockIdx.x; w < n_work; w += gridDim.x) {
    const long long cell = a.n_list >= 0 ? a.cell_list[w] : w;
    const double* g = a.geom + cell * GS;
    for (int i = tid; i < GS; i += nt) sgeo[i] = g[i];
    for (int i = tid; i < ND; i += nt) {
      const int gi = a.l2g[cell * ND + i];
      sidx[i] = gi;
      if (a.system && a.rhs) sU[i] = a.old_nse[gi];
    }
    if (!a.skip_matrix)
      for (int i = tid; i < ND * ND; i += nt) L[i] = 0.0;
    __syncthreads();
    // physical gradients G[q][a][d] = sum_e Kinv[e][d](q) * dphi_ref[q][a][e]
    for (int i = tid; i < NQ * NU; i += nt) {
      const int q = i / NU;
      double r[DIM];
#pragma unroll
      for (int e = 0; e < DIM; ++e) r[e] = __ldg(a.dphi_u + i * DIM + e);
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        double s = 0.0;
#pragma unroll
        for (int e = 0; e < DIM; ++e) s += sgeo[NQ * (1 + e * DIM + d) + q] * r[e];
        G[i * DIM + d] = s;
      }
    }
    __syncthreads();
    if (!a.skip_matrix) {
      // velocity-velocity block
      for (int pr = tid; pr < NU * NU; pr += nt) {
        const int na = pr / NU, nb = pr - na * NU;
        double m = 0.0, gg[DIM][DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
          for (int c = 0; c < DIM; ++c) gg[d][c] = 0.0;
        for (int q = 0; q < NQ; ++q) {
          const double wq = sgeo[q];
          const double pa = sphi[q * NU + na] * wq;
          m += pa * sphi[q * NU + nb];
          double ga[DIM], gb[DIM];
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            ga[d] = G[(q * NU + na) * DIM + d] * wq;
            gb[d] = G[(q * NU + nb) * DIM + d];
          }
#pragma unroll
          for (int d = 0; d < DIM; ++d)
#pragma unroll
            for (int c = 0; c < DIM; ++c) gg[d][c] += ga[d] * gb[c];
        }
        double k = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) k += gg[d][d];
        const double diag = m + nu * k;
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
          const int i = sys_u[c * NU + na];
          if (a.system) {
#pragma unroll
            for (int d = 0; d < DIM; ++d) L[i * ND + sys_u[d * NU + nb]] = (c == d ? diag : 0.0) + nu * gg[d][c];
          } else
            L[i * ND + sys_u[c * NU + nb]] = diag;
        }
      }
      if (a.system) {
        // velocity-pressure coupling
        for (int it = tid; it < NU * NP * DIM; it += nt) {
          const int c = it % DIM, r2 = it / DIM, nb = r2 % NP, na = r2 / NP;
          double s = 0.0;
          for (int q = 0; q < NQ; ++q) s += sgeo[q] * G[(q * NU + na) * DIM + c] * spsi[q * NP + nb];
          const int i = sys_u[c * NU + na], j = sys_p[nb];
          L[i * ND + j] = -s;
          L[j * ND + i] = -s;
        }
      } else {
        for (int it = tid; it < NP * NP; it += nt) {
          const int na = it / NP, nb = it - na * NP;
          double s = 0.0;
          for (int q = 0; q < NQ; ++q) s += sgeo[q] * spsi[q * NP + na] * spsi[q * NP + nb];
          L[sys_p[na] * ND + sys_p[nb]] = s;
        }
      }
    }
    const bool do_rhs = a.system && a.rhs != nullptr && !a.only_constrained;
    if (do_rhs) {
      // old temperature at q
      for (int q = tid; q < NQ; q += nt) {
        double t = 0.0;
        for (int k = 0; k < a.ndt; ++k) t += a.old_temp[a.l2g_t[cell * a.ndt + k]] * __ldg(a.phi_t + q * a.ndt + k);
        sT[q] = t;
      }
      __syncthreads();
      for (int q = tid; q < NQ; q += nt) {
        double u[DIM], gu[DIM][DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
          u[c] = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) gu[c][d] = 0.0;
        }
        for (int n = 0; n < NU; ++n) {
          const double ph = sphi[q * NU + n];
#pragma unroll
          for (int c = 0; c < DIM; ++c) {
            const double U = sU[sys_u[c * NU + n]];
            u[c] += U * ph;
#pragma unroll
            for (int d = 0; d < DIM; ++d) gu[c][d] += U * G[(q * NU + n) * DIM + d];
          }
        }
        double x[DIM], grav[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) x[d] = sgeo[NQ * (1 + DIM * DIM + d) + q];
        coeff_gravity<DIM>(a.prm, x, grav);
        const double rho = 1.0 - a.prm.beta * (sT[q] - a.prm.T_ref);
        double ct[3] = {0.0, 0.0, 0.0};
        if (DIM == 2) {
          ct[0] = -2.0 * u[1];
          ct[1] = 2.0 * u[0];
        } else {
          double cor[3] = {0.0, 0.0, a.prm.cuboid ? a.prm.cor_scale * a.prm.omega : 0.0};
          const double u3[3] = {u[0], u[1], u[DIM - 1]};
          ct[0] = 2.0 * (cor[1] * u3[2] - cor[2] * u3[1]);
          ct[1] = 2.0 * (cor[2] * u3[0] - cor[0] * u3[2]);
          ct[2] = 2.0 * (cor[0] * u3[1] - cor[1] * u3[0]);
        }
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
          double adv = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) adv += u[d] * gu[c][d];
          sF[q * DIM + c] =
              (u[c] + a.prm.dt * rho * (a.prm.g_scale * grav[c]) - a.prm.dt * adv - a.prm.dt * ct[c]) * sgeo[q];
        }
      }
      __syncthreads();
      for (int i = tid; i < ND; i += nt) {
        const int f = a.local_field[i], bs = a.local_base[i];
        double s = 0.0;
        if (f < DIM)
          for (int q = 0; q < NQ; ++q) s += sphi[q * NU + bs] * sF[q * DIM + f];
        sl[i] = s;
      }
    }
    __syncthreads();
    if (a.skip_matrix) {
      // rhs-only scatter with constraint resolution (no inhomogeneity terms: they need L)
      for (int i = tid; i < ND; i += nt) {
        const int li = cs.line_of_dof[sidx[i]];
        if (li < 0)
          red_add_f64(a.rhs + sidx[i], sl[i]);
        else
          for (int k = cs.line_ptr[li]; k < cs.line_ptr[li + 1]; ++k)
            red_add_f64(a.rhs + cs.entry_dof[k], cs.entry_w[k] * sl[i]);
      }
    } else {
      distribute_local_matrix<false>(cs, ND, ND, L, do_rhs ? sl : nullptr, sidx, slines, A, a.system ? a.rhs : nullptr,
                                     tid, nt, a.only_constrained != 0, err);
    }
    __syncthreads();
  }
}

template <int DIM>
size_t th_smem_bytes() {
  using D = ThDims<DIM>;
  size_t dbl = (size_t)D::ND * D::ND + (size_t)D::NQ * D::NU * DIM + D::NQ * D::NU + D::NQ * D::NP + D::GS +
               D::NQ * DIM + 2 * D::ND + 64;
  size_t ints = 2 * D::ND + DIM * D::NU + D::NP + 8;
  return dbl * sizeof(double) + ints * sizeof(int);
}

template <int DIM>
int launch_th(dcp_model* m, const ThArgs& args, const BlockMat& mat) {
  dcp_ctx* ctx = m->ctx;
  const size_t smem = th_smem_bytes<DIM>();
  static bool attr_set[4] = {false, false, false, false};
  if (!attr_set[DIM]) {
    DCP_CUDA(cudaFuncSetAttribute(th_cell_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[DIM] = true;
  }
  const long long n_work = args.n_list >= 0 ? args.n_list : args.n_cells;
  if (n_work == 0) return DCP_OK;
  int per_sm = 1;
  DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, th_cell_kernel<DIM>, 256, smem));
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)ctx->sm_count * per_sm;
  if (grid > n_work) grid = n_work;
  th_cell_kernel<DIM><<<(unsigned)grid, 256, smem, ctx->stream>>>(args, make_view(m->nse_cs), make_view(mat), ctx->d_err);
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

}  // namespace

int dcp_launch_th_cells(dcp_model* m, const dcp_params& p, bool system, const double* old_nse, const double* old_temp,
                        const int32_t* cell_list, int64_t n_list, bool only_constrained_entries) {
  ThArgs a;
  a.n_cells = m->n_cells;
  a.cell_list = cell_list;
  a.n_list = n_list;
  a.geom = m->geom_qn;
  a.l2g = m->nse_l2g;
  a.l2g_t = m->temp_l2g;
  a.local_field = m->nse_local_field;
  a.local_base = m->nse_local_base;
  a.phi_u = m->phi_u_qn;
  a.dphi_u = m->dphi_u_qn;
  a.phi_p = m->phi_p_qn;
  a.phi_t = m->phi_t_qn;
  a.ndt = m->ndt;
  a.old_nse = old_nse;
  a.old_temp = old_temp;
  a.rhs = system ? m->nse_rhs : nullptr;
  a.system = system ? 1 : 0;
  a.only_constrained = only_constrained_entries ? 1 : 0;
  a.skip_matrix = 0;
  a.prm = p;
  const BlockMat& mat = system ? m->nse : m->pre;
  if (m->dim == 3) return launch_th<3>(m, a, mat);
  return launch_th<2>(m, a, mat);
}

// right-hand side only (all cells), used by the row-owner strategy where the matrix is built elsewhere
int dcp_launch_th_rhs(dcp_model* m, const dcp_params& p, const double* old_nse, const double* old_temp) {
  ThArgs a;
  a.n_cells = m->n_cells;
  a.cell_list = nullptr;
  a.n_list = -1;
  a.geom = m->geom_qn;
  a.l2g = m->nse_l2g;
  a.l2g_t = m->temp_l2g;
  a.local_field = m->nse_local_field;
  a.local_base = m->nse_local_base;
  a.phi_u = m->phi_u_qn;
  a.dphi_u = m->dphi_u_qn;
  a.phi_p = m->phi_p_qn;
  a.phi_t = m->phi_t_qn;
  a.ndt = m->ndt;
  a.old_nse = old_nse;
  a.old_temp = old_temp;
  a.rhs = m->nse_rhs;
  a.system = 1;
  a.only_constrained = 0;
  a.skip_matrix = 1;
  a.prm = p;
  if (m->dim == 3) return launch_th<3>(m, a, m->nse);
  return launch_th<2>(m, a, m->nse);
}
