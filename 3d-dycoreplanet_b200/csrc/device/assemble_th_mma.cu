// Classic 3-D Taylor-Hood NSE system / preconditioner on the cells without constrained dofs: the per-cell
// contraction on the FP64 tensor cores (mma.sync.m8n8k4.f64, SASS DMMA), scatter through the position table.
//
// Same integrals as assemble_th_fast.cu (reference: include/core/boussinesq_model.tpp:421-464, 550-687).  ncu of
// the DFMA version showed the kernel bound by shared-memory operand traffic (MIO throttle: 8 LDS per 10 DFMA);
// B200's DMMA runs at the full FP64 rate (37.0 vs 33.9 TFLOP/s DFMA, profiles/r01_fp64_peaks.json) and needs one
// operand double per lane per 256 FMA, so the contraction moves to the tensor pipe:
//
//   X[q][4b+beta] = (d_0 phi_b, d_1 phi_b, d_2 phi_b, phi_b)(x_q)  for the 27 Q2 nodes b, then the 8 psi_b'(x_q)
//   D = X^T diag(w) X      (116 x 116 Gram matrix, K = 27 quadrature points padded to 28)
//
// One 8x8 DMMA tile holds complete 4x4 blocks of 2x2 node pairs, so the epilogue is local to 8 lanes:
//   diag_ab = D[(a,3),(b,3)] + nu * sum_e D[(a,e),(b,e)]                         (three xor-shuffles)
//   L[(a,c),(b,d)] = nu * D[(a,d),(b,c)] + delta_cd diag_ab                      (:626-632)
//   L[(a,c),p_b'] = L[p_b',(a,c)] = -D[(a,c),psi_b']                             (:633-635)
//   preconditioner: L[(a,c),(b,c)] = diag_ab, L[p_a,p_b] = D[psi_a,psi_b]        (:455-462)
// The 6 of 16 cross terms (phi_a d phi_b) of each block are computed and dropped: the tensor pipe has the
// headroom, the scatter (red.global.add.f64, ~240 G/s per-lane issue bound) is what limits this strategy.
#include <omp.h>

#include <algorithm>

#include "scatter.cuh"

namespace {

using namespace dcpdev;

constexpr int NU = 27, NP = 8, NQ = 27, ND = 89, NE = 35, GS = NQ * 13;
constexpr int LDB = 132;        // row stride of X (132 mod 16 == 4: conflict-free fragment loads)
constexpr int KQ = 28;          // quadrature points padded to a multiple of 4
constexpr int PSI0 = 112;       // first psi column (node columns 0..107, 108..111 padding)
constexpr int MTHREADS = 256;

struct MmaArgs {
  long long n_fast;
  const int* cells;
  const unsigned short* pos;
  const unsigned char* nmask;       // [n][36]: unconstrained-component mask of the 27 velocity nodes, 8 pressure flags
  const int* wide_idx;              // preconditioner: -1 or index into pos_wide
  const unsigned short* pos_wide;   // [n_wide][3][27][27]
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
  long long n_u;
  dcp_params prm;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <bool SYSTEM>
__global__ void __launch_bounds__(MTHREADS, 3) th_mma_kernel(MmaArgs a, BlockView A, CsView cs) {
  extern __shared__ __align__(16) double smem[];
  double* X = smem;                     // KQ * LDB
  double* wq = X + KQ * LDB;            // KQ (+4)
  double* sgeo = wq + 32;               // GS
  double* sF = sgeo + GS;               // NQ*3
  double* sU = sF + NQ * 3;             // ND
  double* sT = sU + ND + 1;             // 32
  double* swt_ = sT + 32;                      // 3*NU master weights of the constrained component
  long long* rb00 = (long long*)(swt_ + 3 * NU + 1);  // 3*NU
  long long* rb01 = rb00 + 3 * NU;          // 3*NU
  long long* rb10 = rb01 + 3 * NU;          // NP
  int* sidx = (int*)(rb10 + NP);            // ND
  int* sys_u = sidx + ND;                   // 3*NU
  int* sys_p = sys_u + 3 * NU;              // NP
  unsigned short* spos = (unsigned short*)(sys_p + NP);  // NE*NE
  unsigned short* swide = spos + NE * NE + 1;              // 3*NU*NU (preconditioner, non-uniform cells)
  unsigned char* snm = (unsigned char*)(swide + (SYSTEM ? 0 : 3 * NU * NU) + 1);  // 36
  unsigned char* skc = snm + 36;   // 28: component of the node that is constrained WITH masters (no-normal-flux), 3 = none
  double* swt = swt_;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;

  for (int i = tid; i < ND; i += nt) {
    const int f = a.local_field[i], bs = a.local_base[i];
    if (f < 3) sys_u[f * NU + bs] = i; else sys_p[bs] = i;
  }
  for (int i = tid; i < LDB; i += nt) X[(KQ - 1) * LDB + i] = 0.0;   // padded quadrature point
  for (int i = tid; i < KQ * 4; i += nt) X[(i >> 2) * LDB + 108 + (i & 3)] = 0.0;  // padding node columns
  if (tid < 4) wq[NQ + tid] = 0.0;
  __syncthreads();
  const double nu = a.prm.dt * a.prm.inv_re;
  const bool do_rhs = SYSTEM && a.rhs != nullptr;
  const long long* rp00 = A.rowptr[0][0];
  const long long* rp01 = A.rowptr[0][1];
  const long long* rp10 = SYSTEM ? A.rowptr[1][0] : A.rowptr[1][1];
  double* v00 = A.val[0][0];
  double* v01 = A.val[0][1];
  double* v10 = SYSTEM ? A.val[1][0] : A.val[1][1];

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
      if (tid < 36) snm[tid] = a.nmask[w * 36 + tid];
    }
    const int cflag = a.nmask[w * 36 + 35];  // 1: the cell holds constrained velocity dofs
    __syncthreads();
    if (tid < NU) {
      // no-normal-flux lines: u_k = sum_{c != k} w_c u_c on the same node (verified when the plan was built)
      int kc = 3;
      double wv[3] = {0.0, 0.0, 0.0};
      if (cflag && snm[tid] != 7) {
        const int g0 = sidx[sys_u[tid]];
        for (int c = 0; c < 3; ++c) {
          const int li = cs.line_of_dof[g0 + c];
          if (li >= 0 && cs.line_ptr[li + 1] > cs.line_ptr[li]) {
            kc = c;
            for (int k = cs.line_ptr[li]; k < cs.line_ptr[li + 1]; ++k) wv[cs.entry_dof[k] - g0] = cs.entry_w[k];
          }
        }
      }
      skc[tid] = (unsigned char)kc;
      swt[tid * 3] = wv[0];
      swt[tid * 3 + 1] = wv[1];
      swt[tid * 3 + 2] = wv[2];
    }
    int wide = -1;
    if (!SYSTEM) {
      wide = a.wide_idx[w];
      if (wide >= 0) {
        const unsigned short* p = a.pos_wide + (size_t)wide * (3 * NU * NU);
        for (int i = tid; i < 3 * NU * NU; i += nt) swide[i] = p[i];
      }
    }
    __syncthreads();
    for (int i = tid; i < 3 * NU; i += nt) {
      const int gi = sidx[sys_u[i]];
      rb00[i] = rp00[gi];
      if (SYSTEM) rb01[i] = rp01[gi];
    }
    for (int i = tid; i < NP; i += nt) rb10[i] = rp10[sidx[sys_p[i]] - a.n_u];
    if (tid < NQ) wq[tid] = sgeo[tid];
    // operand table: physical gradients and values of the Q2 functions, then the Q1 pressure functions
    for (int i = tid; i < NQ * NU; i += nt) {
      const int q = i / NU, b = i - q * NU;
      const double r0 = __ldg(a.dphi_u + i * 3), r1 = __ldg(a.dphi_u + i * 3 + 1), r2 = __ldg(a.dphi_u + i * 3 + 2);
      double* x = X + q * LDB + 4 * b;
#pragma unroll
      for (int d = 0; d < 3; ++d) x[d] = sgeo[NQ * (1 + d) + q] * r0 + sgeo[NQ * (4 + d) + q] * r1 + sgeo[NQ * (7 + d) + q] * r2;
      x[3] = __ldg(a.phi_u + i);
    }
    for (int i = tid; i < NQ * NP; i += nt) X[(i / NP) * LDB + PSI0 + (i % NP)] = __ldg(a.phi_p + i);
    __syncthreads();

    // ---- Gram matrix on the tensor cores: 14 node row tiles x (14 node + 1 psi) column tiles, 7 k-steps
    // task = (row tile mt, group of 5 column tiles); preconditioner adds the psi row tile against the psi column tile
    const int n_tasks = SYSTEM ? 42 : 43;
    for (int t = warp; t < n_tasks; t += nwarps) {
      const int mt = t < 42 ? t / 3 : 14, nt0 = t < 42 ? (t % 3) * 5 : 14, ntn = t < 42 ? 5 : 1;
      double acc[5][2];
#pragma unroll
      for (int j = 0; j < 5; ++j) acc[j][0] = acc[j][1] = 0.0;
      const int frow = lane >> 2, fk = lane & 3;
#pragma unroll
      for (int ks = 0; ks < KQ / 4; ++ks) {
        const int q = 4 * ks + fk;
        const double* xr = X + q * LDB;
        const double av = wq[q] * xr[8 * mt + frow];
#pragma unroll
        for (int j = 0; j < 5; ++j)
          if (j < ntn) dmma(acc[j][0], acc[j][1], av, xr[8 * (nt0 + j) + frow]);
      }
      // ---- epilogue: lane holds D[row = lane/4][col = 2*(lane%4) + {0,1}] of every tile
      const int alpha = (lane >> 2) & 3, na = 2 * mt + (lane >> 4);
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        if (j >= ntn) continue;
        const int ntile = nt0 + j;
        if (ntile < 14 && mt < 14) {
          const int nb = 2 * ntile + ((lane >> 1) & 1);
          // diag_ab = D[phi,phi] + nu * trace, gathered over the 8 lanes that share the node pair (a,b)
          const double own = (alpha & 1) ? acc[j][1] : acc[j][0];
          double dg = ((lane & 1) == (alpha >> 1)) ? own * (alpha < 3 ? nu : 1.0) : 0.0;
          dg += __shfl_xor_sync(0xffffffffu, dg, 1);
          dg += __shfl_xor_sync(0xffffffffu, dg, 4);
          dg += __shfl_xor_sync(0xffffffffu, dg, 8);
          const bool validp = na < NU && nb < NU;
          const int ma = validp ? snm[na] : 7, mb = validp ? snm[nb] : 7;
          if (SYSTEM) {
            const int ka = validp ? skc[na] : 3, kb = validp ? skc[nb] : 3;
            // full block value F = L[(a, c = beta_jj), (b, d = alpha)] before constraints
            double F[2];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int beta = 2 * (lane & 1) + jj;
              F[jj] = (alpha < 3 && beta < 3) ? nu * acc[j][jj] + (alpha == beta ? dg : 0.0) : 0.0;
            }
            if (__any_sync(0xffffffffu, ka != 3 || kb != 3)) {
              // rows / columns of a constrained component are redistributed to the other components of the same node:
              // R[c][d] = F[c][d] + wa_c F[ka][d] + wb_d F[c][kb] + wa_c wb_d F[ka][kb]   (C^T L C on the 3x3 block)
              const int kas = ka == 3 ? 0 : ka, kbs = kb == 3 ? 0 : kb;
              const int srcA = (lane & 0x1e) | (kas >> 1), srcB = (lane & 0x13) | (kbs << 2);
              const int srcC = (lane & 0x12) | (kbs << 2) | (kas >> 1);
              const double a0 = __shfl_sync(0xffffffffu, F[0], srcA), a1 = __shfl_sync(0xffffffffu, F[1], srcA);
              const double b0 = __shfl_sync(0xffffffffu, F[0], srcB), b1 = __shfl_sync(0xffffffffu, F[1], srcB);
              const double c0 = __shfl_sync(0xffffffffu, F[0], srcC), c1 = __shfl_sync(0xffffffffu, F[1], srcC);
              const double FA = (kas & 1) ? a1 : a0, FC = (kas & 1) ? c1 : c0;
              const double wb = (kb != 3 && alpha < 3) ? swt[nb * 3 + alpha] : 0.0;
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                const int beta = 2 * (lane & 1) + jj;
                const double wa = (ka != 3 && beta < 3) ? swt[na * 3 + beta] : 0.0;
                F[jj] += wa * FA + wb * (jj ? b1 : b0) + wa * wb * FC;
              }
            }
            if (validp && alpha < 3) {
              const long long off = (long long)spos[na * NE + nb] + __popc(mb & ((1 << alpha) - 1));
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                const int beta = 2 * (lane & 1) + jj;
                if (beta >= 3) continue;
                if ((ma & (1 << beta)) && (mb & (1 << alpha)))
                  red_add_f64(v00 + rb00[beta * NU + na] + off, F[jj]);
                else if (na == nb && alpha == beta && !(ma & (1 << alpha)))
                  // constrained dof: |L_ii| on its own diagonal (the row holds nothing else)
                  red_add_f64(v00 + rb00[alpha * NU + na], fabs(nu * acc[j][jj] + dg));
              }
            }
          } else if (validp && alpha < 3 && (lane & 1) == 0) {
            if (ma & mb & (1 << alpha)) {
              const unsigned off = wide >= 0 ? swide[(alpha * NU + na) * NU + nb] : spos[na * NE + nb];
              if (off != 0xffffu) red_add_f64(v00 + rb00[alpha * NU + na] + off, dg);
            } else if (na == nb && !(ma & (1 << alpha)))
              red_add_f64(v00 + rb00[alpha * NU + na], fabs(dg));
          }
        } else if (ntile == 14 && mt < 14) {
          if (SYSTEM) {
            const bool va = na < NU;
            const int ma = va ? snm[na] : 7, ka = va ? skc[na] : 3;
            double sv[2] = {-acc[j][0], -acc[j][1]};
            if (__any_sync(0xffffffffu, ka != 3)) {
              const int src = (lane & 0x13) | ((ka == 3 ? 0 : ka) << 2);
              const double k0 = __shfl_sync(0xffffffffu, sv[0], src), k1 = __shfl_sync(0xffffffffu, sv[1], src);
              const double wa = (ka != 3 && alpha < 3) ? swt[na * 3 + alpha] : 0.0;
              sv[0] += wa * k0;
              sv[1] += wa * k1;
            }
            if (va && alpha < 3 && (ma & (1 << alpha))) {
              const int rank = __popc(ma & ((1 << alpha) - 1));
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                const int pb = 2 * (lane & 3) + jj;
                if (!snm[NU + pb]) continue;
                red_add_f64(v01 + rb01[alpha * NU + na] + spos[na * NE + NU + pb], sv[jj]);
                red_add_f64(v10 + rb10[pb] + spos[(NU + pb) * NE + na] + rank, sv[jj]);
              }
            }
          }
        } else if (ntile == 14 && mt == 14 && !SYSTEM) {
          const int pa = lane >> 2;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int pb = 2 * (lane & 3) + jj;
            if (snm[NU + pa] && snm[NU + pb]) red_add_f64(v10 + rb10[pa] + spos[(NU + pa) * NE + NU + pb], acc[j][jj]);
          }
        }
      }
    }
    if (do_rhs) {
      for (int q = tid; q < NQ; q += nt) {
        double tq = 0.0;
        for (int k = 0; k < a.ndt; ++k) tq += a.old_temp[a.l2g_t[cell * a.ndt + k]] * __ldg(a.phi_t + q * a.ndt + k);
        sT[q] = tq;
      }
      __syncthreads();
      for (int q = tid; q < NQ; q += nt) {
        double u[3] = {0, 0, 0}, gu[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int n = 0; n < NU; ++n) {
          const double* x = X + q * LDB + 4 * n;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const double U = sU[sys_u[c * NU + n]];
            u[c] += U * x[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) gu[c][d] += U * x[d];
          }
        }
        double xq[3], grav[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) xq[d] = sgeo[NQ * (10 + d) + q];
        if (a.prm.cuboid) {
          grav[0] = grav[1] = 0.0;
          grav[2] = -a.prm.g_const;
        } else {
          const double r = sqrt(xq[0] * xq[0] + xq[1] * xq[1] + xq[2] * xq[2]);
          const double sc = r > 1.0 ? r : sqrt(r);
#pragma unroll
          for (int d = 0; d < 3; ++d) grav[d] = -a.prm.g_const * xq[d] / sc;
        }
        const double rho = 1.0 - a.prm.beta * (sT[q] - a.prm.T_ref);
        const double cz = a.prm.cuboid ? a.prm.cor_scale * a.prm.omega : 0.0;
        const double ct[3] = {2.0 * (-cz * u[1]), 2.0 * (cz * u[0]), 0.0};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double adv = u[0] * gu[c][0] + u[1] * gu[c][1] + u[2] * gu[c][2];
          sF[q * 3 + c] = (u[c] + a.prm.dt * rho * (a.prm.g_scale * grav[c]) - a.prm.dt * adv - a.prm.dt * ct[c]) * sgeo[q];
        }
      }
      __syncthreads();
      for (int i = tid; i < 3 * NU; i += nt) {
        const int c = i / NU, n = i - c * NU;
        double s = 0.0;
        for (int q = 0; q < NQ; ++q) s += X[q * LDB + 4 * n + 3] * sF[q * 3 + c];
        const int gi = sidx[sys_u[i]];
        if (snm[n] & (1 << c))
          red_add_f64(a.rhs + gi, s);
        else {  // constrained dof: its share goes to the masters (none for Dirichlet lines)
          const int li = cs.line_of_dof[gi];
          for (int k = cs.line_ptr[li]; k < cs.line_ptr[li + 1]; ++k) red_add_f64(a.rhs + cs.entry_dof[k], cs.entry_w[k] * s);
        }
      }
    }
    __syncthreads();
  }
}

constexpr size_t mma_smem_bytes(bool system) {
  return sizeof(double) * (KQ * LDB + 32 + GS + NQ * 3 + ND + 1 + 32 + 3 * NU + 1) + sizeof(long long) * (6 * NU + NP) +
         sizeof(int) * (ND + 3 * NU + NP) + sizeof(unsigned short) * (NE * NE + 2 + (system ? 0 : 3 * NU * NU)) + 36 + 28 + 32;
}

}  // namespace

// ---- host-side plan: masked position tables for every node-blocked cell ------------------------------------------
namespace {
struct HCsrM {
  int64_t n_rows = 0;
  const int64_t* rp = nullptr;
  const int32_t* col = nullptr;
  int64_t find(int64_t r, int32_t c) const {
    if (!rp || r < 0 || r >= n_rows) return -1;
    const int32_t* b = col + rp[r];
    const int32_t* e = col + rp[r + 1];
    const int32_t* p = std::lower_bound(b, e, c);
    return (p == e || *p != c) ? -1 : p - b;
  }
  bool at(int64_t r, int64_t off, int32_t c) const {
    return rp && r >= 0 && r < n_rows && off >= 0 && rp[r] + off < rp[r + 1] && col[rp[r] + off] == c;
  }
};
template <class T>
int upm(dcp_ctx* ctx, T** dst, const std::vector<T>& v) {
  *dst = nullptr;
  if (v.empty()) return DCP_OK;
  if (cudaMalloc((void**)dst, v.size() * sizeof(T)) != cudaSuccess ||
      cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
    dcp_set_error("masked plan: device allocation / copy failed");
    return DCP_ERR_CUDA;
  }
  return DCP_OK;
}
}  // namespace

void dcp_masked_plan_free(MaskedPlan* p) {
  if (!p) return;
  cudaFree(p->cells);
  cudaFree(p->other_cells);
  cudaFree(p->pos);
  cudaFree(p->nmask);
  cudaFree(p->wide_idx);
  cudaFree(p->pos_wide);
  delete p;
}

// Position tables with constraint masks for the classic 3-D family.  Every cell whose numbering is node-blocked and
// whose rows verify gets an entry; the few that do not are listed in other_cells and go through the general kernel.
int dcp_masked_plan_build(dcp_model* m, const dcp_model_desc* d, bool system, MaskedPlan** out) {
  *out = nullptr;
  const int64_t nc = d->n_cells, n_u = d->nse_block_size[0];
  const dcp_csr_desc(*pat)[DCP_MAX_BLOCKS] = system ? d->nse_pattern : d->pre_pattern;
  const HCsrM A00{pat[0][0].n_rows, pat[0][0].rowptr, pat[0][0].col}, A01{pat[0][1].n_rows, pat[0][1].rowptr, pat[0][1].col};
  const HCsrM A10{pat[1][0].n_rows, pat[1][0].rowptr, pat[1][0].col}, A11{pat[1][1].n_rows, pat[1][1].rowptr, pat[1][1].col};
  std::vector<int> sys_u(3 * NU), sys_p(NP);
  for (int i = 0; i < ND; ++i) {
    const int f = d->nse_local_field[i], b = d->nse_local_base[i];
    if (f < 3) sys_u[f * NU + b] = i; else sys_p[b] = i;
  }
  std::vector<int32_t> lod((size_t)d->nse_cs.n_dofs, -1);
  for (int64_t l = 0; l < d->nse_cs.n_lines; ++l) lod[d->nse_cs.line_dof[l]] = (int32_t)l;
  std::vector<uint8_t> ok((size_t)nc, 0), is_wide((size_t)nc, 0);
  std::vector<uint16_t> pos_all((size_t)nc * NE * NE, 0xFFFF);
  std::vector<uint8_t> mask_all((size_t)nc * 36, 0);
  std::vector<std::vector<uint16_t>> wide_rows((size_t)nc);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t c = 0; c < nc; ++c) {
    const int32_t* idx = d->nse_l2g + c * ND;
    bool good = true;
    for (int a = 0; a < NU && good; ++a)
      for (int k = 1; k < 3 && good; ++k) good = idx[sys_u[k * NU + a]] == idx[sys_u[a]] + k;
    for (int a = 0; a < NP && good; ++a) good = idx[sys_p[a]] >= n_u;
    if (!good) continue;
    uint16_t* P = &pos_all[(size_t)c * NE * NE];
    uint8_t* mk = &mask_all[(size_t)c * 36];
    int first[NU];
    for (int a = 0; a < NU; ++a) {
      int mm = 0;
      first[a] = -1;
      for (int k = 0; k < 3; ++k)
        if (lod[idx[sys_u[a]] + k] < 0) {
          mm |= 1 << k;
          if (first[a] < 0) first[a] = k;
        }
      mk[a] = (uint8_t)mm;
    }
    for (int a = 0; a < NP; ++a) {
      mk[NU + a] = lod[idx[sys_p[a]]] < 0 ? 1 : 0;
      if (!mk[NU + a]) good = false;  // constrained pressure dofs: general path
    }
    // supported constraint lines: homogeneous, masters (if any) are the other components of the same node, at most one
    // such line per node (Dirichlet and no-normal-flux lines of boussinesq_model.tpp:313-329); anything else (periodic,
    // hanging nodes, inhomogeneous) sends the cell to the general kernel
    bool any_cs = false;
    for (int a = 0; a < NU && good; ++a) {
      int n_master_lines = 0;
      const int32_t g0 = idx[sys_u[a]];
      for (int k = 0; k < 3 && good; ++k) {
        const int32_t li = lod[g0 + k];
        if (li < 0) continue;
        any_cs = true;
        if (d->nse_cs.inhom[li] != 0.0) good = false;
        const int32_t e0 = d->nse_cs.line_ptr[li], e1 = d->nse_cs.line_ptr[li + 1];
        if (e1 > e0) {
          ++n_master_lines;
          if (!system) good = false;  // preconditioner: redistributed entries leave the same-component pattern
          for (int32_t e = e0; e < e1 && good; ++e) {
            const int32_t md = d->nse_cs.entry_dof[e];
            good = md >= g0 && md < g0 + 3 && md != g0 + k && lod[md] < 0;
          }
        }
      }
      if (n_master_lines > 1) good = false;
    }
    mk[35] = any_cs ? 1 : 0;
    if (!good) continue;
    if (system) {
      for (int a = 0; a < NU && good; ++a) {
        if (!mk[a]) continue;
        const int64_t r0 = idx[sys_u[a]];
        for (int b = 0; b < NU && good; ++b) {
          if (!mk[b]) continue;
          const int32_t c0 = idx[sys_u[b]];
          const int64_t off = A00.find(r0 + first[a], c0 + first[b]);
          good = off >= 0 && off + 2 < 65535;
          for (int k = 0; k < 3 && good; ++k)
            if (mk[a] & (1 << k)) {
              int rank = 0;
              for (int dd = 0; dd < 3 && good; ++dd)
                if (mk[b] & (1 << dd)) good = A00.at(r0 + k, off + rank++, c0 + dd);
            }
          if (good) P[a * NE + b] = (uint16_t)off;
        }
        for (int b = 0; b < NP && good; ++b) {
          if (!mk[NU + b]) continue;
          const int32_t cp = (int32_t)(idx[sys_p[b]] - n_u);
          const int64_t off = A01.find(r0 + first[a], cp);
          good = off >= 0 && off < 65535;
          for (int k = 0; k < 3 && good; ++k)
            if (mk[a] & (1 << k)) good = A01.at(r0 + k, off, cp);
          if (good) P[a * NE + NU + b] = (uint16_t)off;
        }
      }
      for (int a = 0; a < NP && good; ++a) {
        if (!mk[NU + a]) continue;
        const int64_t rp = idx[sys_p[a]] - n_u;
        for (int b = 0; b < NU && good; ++b) {
          if (!mk[b]) continue;
          const int32_t c0 = idx[sys_u[b]];
          const int64_t off = A10.find(rp, c0 + first[b]);
          good = off >= 0 && off + 2 < 65535;
          int rank = 0;
          for (int dd = 0; dd < 3 && good; ++dd)
            if (mk[b] & (1 << dd)) good = A10.at(rp, off + rank++, c0 + dd);
          if (good) P[(NU + a) * NE + b] = (uint16_t)off;
        }
      }
    } else {
      // preconditioner: row (a,k) holds column (b,k); the offset is usually the same for the three k
      std::vector<uint16_t> W(3 * NU * NU, 0xFFFF);
      bool uniform = true;
      for (int a = 0; a < NU && good; ++a)
        for (int b = 0; b < NU && good; ++b) {
          int64_t common = -1;
          for (int k = 0; k < 3 && good; ++k) {
            if (!(mk[a] & mk[b] & (1 << k))) continue;
            const int64_t off = A00.find((int64_t)idx[sys_u[a]] + k, idx[sys_u[b]] + k);
            good = off >= 0 && off < 65535;
            if (!good) break;
            W[(k * NU + a) * NU + b] = (uint16_t)off;
            if (common < 0) common = off; else if (common != off) uniform = false;
          }
          if (good && common >= 0) P[a * NE + b] = (uint16_t)common;
        }
      for (int a = 0; a < NP && good; ++a) {
        if (!mk[NU + a]) continue;
        for (int b = 0; b < NP && good; ++b) {
          if (!mk[NU + b]) continue;
          const int64_t off = A11.find(idx[sys_p[a]] - n_u, (int32_t)(idx[sys_p[b]] - n_u));
          good = off >= 0 && off < 65535;
          if (good) P[(NU + a) * NE + NU + b] = (uint16_t)off;
        }
      }
      if (good && !uniform) {
        is_wide[c] = 1;
        wide_rows[c].swap(W);
      }
    }
    ok[c] = good ? 1 : 0;
  }
  std::vector<int32_t> cells, other, wide_idx;
  std::vector<uint16_t> pos, pos_wide;
  std::vector<uint8_t> nmask;
  for (int64_t c = 0; c < nc; ++c) {
    if (!ok[c]) {
      other.push_back((int32_t)c);
      continue;
    }
    cells.push_back((int32_t)c);
    if (is_wide[c]) {
      wide_idx.push_back((int32_t)(pos_wide.size() / (3 * NU * NU)));
      pos_wide.insert(pos_wide.end(), wide_rows[c].begin(), wide_rows[c].end());
    } else
      wide_idx.push_back(-1);
  }
  pos.resize(cells.size() * (size_t)(NE * NE));
  nmask.resize(cells.size() * (size_t)36);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)cells.size(); ++i) {
    std::copy(&pos_all[(size_t)cells[i] * NE * NE], &pos_all[(size_t)cells[i] * NE * NE] + NE * NE, &pos[(size_t)i * NE * NE]);
    std::copy(&mask_all[(size_t)cells[i] * 36], &mask_all[(size_t)cells[i] * 36] + 36, &nmask[(size_t)i * 36]);
  }
  MaskedPlan* P = new MaskedPlan;
  P->n = (int64_t)cells.size();
  P->n_other = (int64_t)other.size();
  P->n_wide = (int64_t)(pos_wide.size() / (3 * NU * NU));
  dcp_ctx* ctx = m->ctx;
  int rc = upm(ctx, &P->cells, cells);
  if (rc == DCP_OK) rc = upm(ctx, &P->other_cells, other);
  if (rc == DCP_OK) rc = upm(ctx, &P->pos, pos);
  if (rc == DCP_OK) rc = upm(ctx, &P->nmask, nmask);
  if (rc == DCP_OK) rc = upm(ctx, &P->wide_idx, wide_idx);
  if (rc == DCP_OK) rc = upm(ctx, &P->pos_wide, pos_wide);
  cudaStreamSynchronize(ctx->stream);
  if (rc != DCP_OK) {
    dcp_masked_plan_free(P);
    return rc;
  }
  *out = P;
  return DCP_OK;
}

int dcp_launch_th_mma(dcp_model* m, const dcp_params& p, bool system, const MaskedPlan* plan, const double* old_nse,
                      const double* old_temp) {
  dcp_ctx* ctx = m->ctx;
  MmaArgs a;
  a.n_fast = plan->n;
  a.cells = plan->cells;
  a.pos = plan->pos;
  a.nmask = plan->nmask;
  a.wide_idx = plan->wide_idx;
  a.pos_wide = plan->pos_wide;
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
  a.n_u = m->nse.start[1];
  a.prm = p;
  if (a.n_fast == 0) return DCP_OK;
  const size_t smem = mma_smem_bytes(system);
  static bool attr = false;
  if (!attr) {
    DCP_CUDA(cudaFuncSetAttribute(th_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma_smem_bytes(true)));
    DCP_CUDA(cudaFuncSetAttribute(th_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma_smem_bytes(false)));
    attr = true;
  }
  int per_sm = 1;
  if (system)
    DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, th_mma_kernel<true>, MTHREADS, smem));
  else
    DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, th_mma_kernel<false>, MTHREADS, smem));
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)ctx->sm_count * per_sm;
  if (grid > a.n_fast) grid = a.n_fast;
  const BlockMat& mat = system ? m->nse : m->pre;
  if (system)
    th_mma_kernel<true><<<(unsigned)grid, MTHREADS, smem, ctx->stream>>>(a, make_view(mat), make_view(m->nse_cs));
  else
    th_mma_kernel<false><<<(unsigned)grid, MTHREADS, smem, ctx->stream>>>(a, make_view(mat), make_view(m->nse_cs));
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}
