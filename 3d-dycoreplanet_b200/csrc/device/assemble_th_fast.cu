// Classic Taylor-Hood NSE system / preconditioner: position-table path (DCP_STRATEGY_POSITIONS).
//
// Same integrals as assemble_th.cu (reference: include/core/boussinesq_model.tpp:421-464, 550-687) for the
// cells that hold no constrained dof.  For those cells the constraint machinery of
// AffineConstraints::distribute_local_to_global is the identity, so the 89x89 local matrix never has to
// exist: each thread owns one (row node a, column node b) pair, keeps its 3x3 block in registers and adds it
// straight into the CSR values with red.global.add.f64 at `rowptr[row] + position[a][b] + d`.  The position
// table (uint16 per node pair, built once per mesh on the host) replaces the per-entry binary search of
// the general path; it relies on the three velocity components of a node being adjacent dofs and columns
// (true after DoFRenumbering::component_wise, :204) and is verified entry by entry when it is built --
// cells that fail the check fall back to the general path.
#include <omp.h>

#include <algorithm>

#include "scatter.cuh"

namespace {

using namespace dcpdev;

template <int DIM>
struct FDims {
  static constexpr int NU = DIM == 3 ? 27 : 9;
  static constexpr int NP = DIM == 3 ? 8 : 4;
  static constexpr int NQ = NU;
  static constexpr int ND = DIM * NU + NP;
  static constexpr int NE = NU + NP;
  static constexpr int GS = NQ * (1 + DIM * DIM + DIM);
};

struct FastArgs {
  long long n_fast;
  const int* cells;
  const unsigned short* pos;
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
  int system;
  long long n_u;
  dcp_params prm;
};

template <int DIM>
__device__ __forceinline__ void fast_gravity(const dcp_params& P, const double* x, double* g) {
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
__global__ void __launch_bounds__(256, 3) th_fast_kernel(FastArgs a, BlockView A) {
  using D = FDims<DIM>;
  constexpr int NU = D::NU, NP = D::NP, NQ = D::NQ, ND = D::ND, NE = D::NE, GS = D::GS;
  extern __shared__ double smem[];
  double* G = smem;                      // NQ*NU*DIM
  double* sphi = G + NQ * NU * DIM;      // NQ*NU
  double* spsi = sphi + NQ * NU;         // NQ*NP
  double* sgeo = spsi + NQ * NP;         // GS
  double* sF = sgeo + GS;                // NQ*DIM
  double* sU = sF + NQ * DIM;            // ND
  double* sT = sU + ND;                  // NQ (+pad)
  long long* rb00 = (long long*)(sT + 32);  // DIM*NU   row starts in block(0,0) of the velocity rows
  long long* rb01 = rb00 + DIM * NU;        // DIM*NU   ... in block(0,1)
  long long* rb10 = rb01 + DIM * NU;        // NP       pressure rows in block(1,0) [system] / block(1,1) [precond]
  int* sidx = (int*)(rb10 + NP);            // ND
  int* sys_u = sidx + ND;                   // DIM*NU
  int* sys_p = sys_u + DIM * NU;            // NP
  unsigned short* spos = (unsigned short*)(sys_p + NP);  // NE*NE
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int i = tid; i < NQ * NU; i += nt) sphi[i] = a.phi_u[i];
  for (int i = tid; i < NQ * NP; i += nt) spsi[i] = a.phi_p[i];
  for (int i = tid; i < ND; i += nt) {
    const int f = a.local_field[i], bs = a.local_base[i];
    if (f < DIM) sys_u[f * NU + bs] = i; else sys_p[bs] = i;
  }
  __syncthreads();
  const double nu = a.prm.dt * a.prm.inv_re;
  const bool do_rhs = a.system && a.rhs != nullptr;
  const long long* rp00 = A.rowptr[0][0];
  const long long* rp01 = A.rowptr[0][1];
  const long long* rp10 = a.system ? A.rowptr[1][0] : A.rowptr[1][1];
  double* v00 = A.val[0][0];
  double* v01 = A.val[0][1];
  double* v10 = a.system ? A.val[1][0] : A.val[1][1];

  for (long long w = blockIdx.x; w < a.n_fast; w += gridDim.x) {
    const long long cell = a.cells[w];
    const double* g = a.geom + cell * GS;
    for (int i = tid; i < GS; i += nt) sgeo[i] = g[i];
    for (int i = tid; i < ND; i += nt) {
      const int gi = a.l2g[cell * ND + i];
      sidx[i] = gi;
      if (do_rhs) sU[i] = a.old_nse[gi];
    }
    {
      const unsigned short* p = a.pos + w * (NE * NE);
      for (int i = tid; i < NE * NE; i += nt) spos[i] = p[i];
    }
    __syncthreads();
    for (int i = tid; i < DIM * NU; i += nt) {
      const int gi = sidx[sys_u[i]];
      rb00[i] = rp00[gi];
      if (a.system) rb01[i] = rp01[gi];
    }
    for (int i = tid; i < NP; i += nt) rb10[i] = rp10[sidx[sys_p[i]] - a.n_u];
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
    // velocity-velocity
    for (int pr = tid; pr < NU * NU; pr += nt) {
      const int na = pr / NU, nb = pr - na * NU;
      double m = 0.0, gg[DIM][DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int c = 0; c < DIM; ++c) gg[d][c] = 0.0;
#pragma unroll 3
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
      const long long off = spos[na * NE + nb];
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
        double* row = v00 + rb00[c * NU + na] + off;
        if (a.system) {
#pragma unroll
          for (int d = 0; d < DIM; ++d) red_add_f64(row + d, (c == d ? diag : 0.0) + nu * gg[d][c]);
        } else
          red_add_f64(row, diag);
      }
    }
    if (a.system) {
      for (int it = tid; it < NU * NP * DIM; it += nt) {
        const int c = it % DIM, r2 = it / DIM, nb = r2 % NP, na = r2 / NP;
        double s = 0.0;
        for (int q = 0; q < NQ; ++q) s += sgeo[q] * G[(q * NU + na) * DIM + c] * spsi[q * NP + nb];
        red_add_f64(v01 + rb01[c * NU + na] + spos[na * NE + NU + nb], -s);
        red_add_f64(v10 + rb10[nb] + spos[(NU + nb) * NE + na] + c, -s);
      }
    } else {
      for (int it = tid; it < NP * NP; it += nt) {
        const int na = it / NP, nb = it - na * NP;
        double s = 0.0;
        for (int q = 0; q < NQ; ++q) s += sgeo[q] * spsi[q * NP + na] * spsi[q * NP + nb];
        red_add_f64(v10 + rb10[na] + spos[(NU + na) * NE + NU + nb], s);
      }
    }
    if (do_rhs) {
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
        fast_gravity<DIM>(a.prm, x, grav);
        const double rho = 1.0 - a.prm.beta * (sT[q] - a.prm.T_ref);
        double ct[3] = {0.0, 0.0, 0.0};
        if (DIM == 2) {
          ct[0] = -2.0 * u[1];
          ct[1] = 2.0 * u[0];
        } else {
          const double cz = a.prm.cuboid ? a.prm.cor_scale * a.prm.omega : 0.0;
          ct[0] = 2.0 * (-cz * u[1]);
          ct[1] = 2.0 * (cz * u[0]);
          ct[2] = 0.0;
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
      for (int i = tid; i < DIM * NU; i += nt) {
        const int c = i / NU, n = i - c * NU;
        double s = 0.0;
        for (int q = 0; q < NQ; ++q) s += sphi[q * NU + n] * sF[q * DIM + c];
        red_add_f64(a.rhs + sidx[sys_u[i]], s);
      }
    }
    __syncthreads();
  }
}

template <int DIM>
size_t fast_smem_bytes() {
  using D = FDims<DIM>;
  size_t dbl = (size_t)D::NQ * D::NU * DIM + D::NQ * D::NU + D::NQ * D::NP + D::GS + D::NQ * DIM + D::ND + 32;
  size_t i64 = 2 * DIM * D::NU + D::NP;
  size_t ints = D::ND + DIM * D::NU + D::NP;
  size_t u16 = (size_t)D::NE * D::NE;
  return dbl * 8 + i64 * 8 + ints * 4 + u16 * 2 + 16;
}

// ---- host-side plan builder ---------------------------------------------------------------------------
struct HostCsr {
  int64_t n_rows = 0;
  const int64_t* rp = nullptr;
  const int32_t* col = nullptr;
};

inline int64_t find_col(const HostCsr& A, int64_t row, int32_t c) {
  if (!A.rp || row < 0 || row >= A.n_rows) return -1;
  const int32_t* b = A.col + A.rp[row];
  const int32_t* e = A.col + A.rp[row + 1];
  const int32_t* p = std::lower_bound(b, e, c);
  if (p == e || *p != c) return -1;
  return p - b;
}
inline bool col_at(const HostCsr& A, int64_t row, int64_t off, int32_t c) {
  if (!A.rp || row < 0 || row >= A.n_rows) return false;
  if (off < 0 || A.rp[row] + off >= A.rp[row + 1]) return false;
  return A.col[A.rp[row] + off] == c;
}

}  // namespace

void dcp_fast_plan_free(FastPlan* p) {
  if (!p) return;
  cudaFree(p->fast_cells);
  cudaFree(p->general_cells);
  cudaFree(p->pos);
  delete p;
}

// Builds the position table for all unconstrained cells.  `system`: nse pattern (else preconditioner pattern).
int dcp_fast_plan_build(dcp_model* m, const dcp_model_desc* d, bool system, FastPlan** out) {
  *out = nullptr;
  const int dim = d->dim;
  const int NU = dim == 3 ? 27 : 9, NP = dim == 3 ? 8 : 4, ND = dim * NU + NP, NE = NU + NP;
  const int64_t n_u = d->nse_block_size[0];
  const dcp_csr_desc(*pat)[DCP_MAX_BLOCKS] = system ? d->nse_pattern : d->pre_pattern;
  HostCsr A00{pat[0][0].n_rows, pat[0][0].rowptr, pat[0][0].col};
  HostCsr A01{pat[0][1].n_rows, pat[0][1].rowptr, pat[0][1].col};
  HostCsr A10{pat[1][0].n_rows, pat[1][0].rowptr, pat[1][0].col};
  HostCsr A11{pat[1][1].n_rows, pat[1][1].rowptr, pat[1][1].col};
  std::vector<int> sys_u(dim * NU), sys_p(NP);
  for (int i = 0; i < ND; ++i) {
    int f = d->nse_local_field[i], b = d->nse_local_base[i];
    if (f < dim) sys_u[f * NU + b] = i; else sys_p[b] = i;
  }
  std::vector<int32_t> lod((size_t)d->nse_cs.n_dofs, -1);
  for (int64_t l = 0; l < d->nse_cs.n_lines; ++l) lod[d->nse_cs.line_dof[l]] = (int32_t)l;
  const int64_t nc = d->n_cells;
  std::vector<uint8_t> ok((size_t)nc, 0);
  std::vector<uint16_t> pos_all((size_t)nc * NE * NE, 0xFFFF);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t c = 0; c < nc; ++c) {
    const int32_t* idx = d->nse_l2g + c * ND;
    bool good = true;
    for (int i = 0; i < ND && good; ++i) good = lod[idx[i]] < 0;
    // node-blocked numbering
    for (int a = 0; a < NU && good; ++a)
      for (int k = 1; k < dim && good; ++k) good = idx[sys_u[k * NU + a]] == idx[sys_u[a]] + k;
    for (int a = 0; a < NP && good; ++a) good = idx[sys_p[a]] >= n_u;
    if (!good) continue;
    uint16_t* P = &pos_all[(size_t)c * NE * NE];
    for (int a = 0; a < NU && good; ++a) {
      const int64_t r0 = idx[sys_u[a]];
      for (int b = 0; b < NU && good; ++b) {
        const int32_t c0 = idx[sys_u[b]];
        if (system) {
          int64_t off = find_col(A00, r0, c0);
          good = off >= 0 && off + dim - 1 < 65535;
          for (int k = 0; k < dim && good; ++k)
            for (int dd = 0; dd < dim && good; ++dd) good = col_at(A00, r0 + k, off + dd, c0 + dd);
          if (good) P[a * NE + b] = (uint16_t)off;
        } else {
          // preconditioner: row (a,k) holds column (b,k) at offset off_k (the rank of node b among the
          // coupled nodes); the kernel uses one offset for all k, so off_k must not depend on k
          int64_t off0 = find_col(A00, r0, c0);
          good = off0 >= 0 && off0 < 65535;
          for (int k = 1; k < dim && good; ++k) good = col_at(A00, r0 + k, off0, c0 + k);
          if (good) P[a * NE + b] = (uint16_t)off0;
        }
      }
      if (system)
        for (int b = 0; b < NP && good; ++b) {
          const int32_t cp = (int32_t)(idx[sys_p[b]] - n_u);
          int64_t off = find_col(A01, r0, cp);
          good = off >= 0 && off < 65535;
          for (int k = 1; k < dim && good; ++k) good = col_at(A01, r0 + k, off, cp);
          if (good) P[a * NE + NU + b] = (uint16_t)off;
        }
    }
    for (int a = 0; a < NP && good; ++a) {
      const int64_t rp = idx[sys_p[a]] - n_u;
      if (system) {
        for (int b = 0; b < NU && good; ++b) {
          const int32_t c0 = idx[sys_u[b]];
          int64_t off = find_col(A10, rp, c0);
          good = off >= 0 && off + dim - 1 < 65535;
          for (int dd = 1; dd < dim && good; ++dd) good = col_at(A10, rp, off + dd, c0 + dd);
          if (good) P[(NU + a) * NE + b] = (uint16_t)off;
        }
      } else {
        for (int b = 0; b < NP && good; ++b) {
          int64_t off = find_col(A11, rp, (int32_t)(idx[sys_p[b]] - n_u));
          good = off >= 0 && off < 65535;
          if (good) P[(NU + a) * NE + NU + b] = (uint16_t)off;
        }
      }
    }
    ok[c] = good ? 1 : 0;
  }
  std::vector<int32_t> fast, general;
  for (int64_t c = 0; c < nc; ++c) (ok[c] ? fast : general).push_back((int32_t)c);
  std::vector<uint16_t> pos((size_t)fast.size() * NE * NE);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)fast.size(); ++i)
    std::copy(&pos_all[(size_t)fast[i] * NE * NE], &pos_all[(size_t)fast[i] * NE * NE] + NE * NE, &pos[(size_t)i * NE * NE]);
  FastPlan* p = new FastPlan;
  p->ne = NE;
  p->n_fast = (int64_t)fast.size();
  p->n_general = (int64_t)general.size();
  dcp_ctx* ctx = m->ctx;
  int rc = dcp_upload(ctx, &p->fast_cells, fast.data(), (int64_t)fast.size());
  if (rc == DCP_OK) rc = dcp_upload(ctx, &p->general_cells, general.data(), (int64_t)general.size());
  if (rc == DCP_OK && !pos.empty()) {
    if (cudaMalloc((void**)&p->pos, pos.size() * sizeof(uint16_t)) != cudaSuccess ||
        cudaMemcpyAsync(p->pos, pos.data(), pos.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
      dcp_set_error("fast plan: device allocation / copy of the position table failed");
      rc = DCP_ERR_CUDA;
    }
  }
  cudaStreamSynchronize(ctx->stream);
  if (rc != DCP_OK) {
    dcp_fast_plan_free(p);
    return rc;
  }
  *out = p;
  return DCP_OK;
}

int64_t dcp_fast_plan_counts(const FastPlan* p, int64_t* n_general) {
  if (n_general) *n_general = p ? p->n_general : 0;
  return p ? p->n_fast : 0;
}
const int32_t* dcp_fast_plan_general_cells(const FastPlan* p) { return p ? p->general_cells : nullptr; }

template <int DIM>
static int launch_fast(dcp_model* m, const FastArgs& args, const BlockMat& mat) {
  dcp_ctx* ctx = m->ctx;
  const size_t smem = fast_smem_bytes<DIM>();
  static bool attr_set[4] = {false, false, false, false};
  if (!attr_set[DIM]) {
    DCP_CUDA(cudaFuncSetAttribute(th_fast_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[DIM] = true;
  }
  if (args.n_fast == 0) return DCP_OK;
  int per_sm = 1;
  DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, th_fast_kernel<DIM>, 256, smem));
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)ctx->sm_count * per_sm;
  if (grid > args.n_fast) grid = args.n_fast;
  th_fast_kernel<DIM><<<(unsigned)grid, 256, smem, ctx->stream>>>(args, make_view(mat));
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

int dcp_launch_th_fast(dcp_model* m, const dcp_params& p, bool system, const FastPlan* plan, const double* old_nse,
                       const double* old_temp) {
  FastArgs a;
  a.n_fast = plan->n_fast;
  a.cells = plan->fast_cells;
  a.pos = plan->pos;
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
  a.n_u = m->nse.start[1];
  a.prm = p;
  const BlockMat& mat = system ? m->nse : m->pre;
  if (m->dim == 3) return launch_fast<3>(m, a, mat);
  return launch_fast<2>(m, a, mat);
}
