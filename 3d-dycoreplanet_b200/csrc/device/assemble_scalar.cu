// Temperature mass / stiffness matrices and right-hand side (scalar Q1 or Q2 Lagrange space).
//
// Replaces local_assemble_temperature_matrix + copier (/root/reference/include/core/boussinesq_model.tpp:
// 748-817) and local_assemble_temperature_rhs + copier (:873-964); identical integrands are used by the
// FEEC model (include/core/boussineq_model_FEEC.tpp:883-952, 1008-1099; there the advecting velocity is the
// Raviart-Thomas field, Piola-mapped and without the face sign).
//   M[i,j] += phi_i phi_j JxW,   K[i,j] += (1/Pe) grad phi_i . grad phi_j JxW      (:786-797)
//   r[i]   += (phi_i T_old - tau phi_i (u_new . grad T_old) - tau*0*phi_i) JxW,  tau = dt/n   (:928-937)
//   matrix_for_bc(j,i) = (phi_i phi_j + tau/Pe grad phi_i . grad phi_j) JxW  for inhomogeneous i (:939-949)
//
// One warp per cell.  Quadrature points are processed 32 at a time with ONE POINT PER LANE: the lane reads its
// column of the cell's mapping record (coalesced), maps all shape functions of its point and parks
// (phi, d_x phi, d_y phi, d_z phi) in a shared-memory row S[lane][.] of odd stride (conflict-free for the
// writer "lane = point" and for the readers "lane = matrix entry").  Lanes then own the symmetric local
// matrix entries.  8..27 dofs per cell: these kernels are HBM/latency bound, not FLOP bound.
#include <algorithm>
#include <cstdlib>

#include "scatter.cuh"

namespace {

using namespace dcpdev;

constexpr int MAX_ND = 27;

struct ScalarArgs {
  long long n_cells;
  int nd, nq;
  int gstride;          // doubles per cell record
  int feec;             // rhs: velocity is the Raviart-Thomas field (Piola-mapped, unsigned)
  const double* geom;
  const int* l2g;
  const unsigned short* tpos;  // [n_cells][nd*nd] positions inside the row, tpos[cell][0] == 0xffff: general scatter
  const double* phi;   // [nq][nd]
  const double* dphi;  // [nq][nd][dim]
  // point-fastest copies for the loads with lane = quadrature point (2 cache lines per warp-wide load instead of 27)
  const double* phiT;    // [nd][nq]
  const double* dphiT;   // [nd][dim][nq]
  const double* phi_uT;  // classic velocity base element on the temperature rule, [ndu][nq]
  dcp_params prm;
  // rhs only
  const int* l2g_nse;
  int nd_nse, ndu;
  const int* nse_field;
  const int* nse_base;
  const double* phi_u;  // classic: [nq][ndu] velocity base element on the temperature rule; FEEC: [nq][6][3]
  const double* old_temp;
  const double* nse_solution;
  double* rhs;
  const double* q2_tab;            // Q2 matrix kernel: reference gradients and values in table-build order
  const int* cell_list;            // rhs, full kernel: the cells to process (nullptr: all)
  const unsigned char* bc_flag;    // rhs, plain kernel: cells to skip (they hold inhomogeneously constrained dofs)
};

// per-warp shared scratch, carved from dynamic shared memory
struct ScalarLayout {
  int sv;        // stride of S rows (4*nd rounded up to odd)
  int doubles;   // doubles per warp
  int o_S, o_A, o_B, o_c, o_l, o_T, o_U, o_idx;
};
__host__ __device__ inline ScalarLayout scalar_layout(int nd, int nd_nse) {
  ScalarLayout L;
  L.sv = (4 * nd) | 1;
  int o = 0;
  L.o_S = o; o += 32 * L.sv;
  L.o_A = o; o += nd * nd;
  L.o_B = o; o += nd * nd;
  L.o_c = o; o += 32 * 2;          // per point: weight, rhs coefficient
  L.o_l = o; o += MAX_ND + 1;
  L.o_T = o; o += MAX_ND + 1;
  L.o_U = o; o += nd_nse + 1;      // velocity dofs of the cell (rhs)
  L.o_idx = o; o += MAX_ND + 5;    // ints: idx[nd], lines[nd] (2 ints per double slot)
  L.doubles = o;
  return L;
}

// lane = quadrature point q: map all shape functions at q into row[k*4 + {0:phi,1..3:grad}]
template <int DIM>
__device__ __forceinline__ void map_point(const ScalarArgs& a, const double* g, int q, double* row) {
  double K[DIM][DIM];
#pragma unroll
  for (int e = 0; e < DIM; ++e)
#pragma unroll
    for (int d = 0; d < DIM; ++d) K[e][d] = g[a.nq * (1 + e * DIM + d) + q];
  for (int k = 0; k < a.nd; ++k) {
    double r[DIM];
#pragma unroll
    for (int e = 0; e < DIM; ++e) r[e] = __ldg(a.dphiT + (size_t)(k * DIM + e) * a.nq + q);
    row[k * 4] = __ldg(a.phiT + (size_t)k * a.nq + q);
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double v = 0.0;
#pragma unroll
      for (int e = 0; e < DIM; ++e) v += K[e][d] * r[e];
      row[k * 4 + 1 + d] = v;
    }
    if (DIM == 2) row[k * 4 + 3] = 0.0;
  }
}

template <int DIM>
__global__ void __launch_bounds__(128) temperature_matrix_kernel(ScalarArgs a, CsView cs, BlockView Mass, BlockView Stiff,
                                                                 int* err) {
  extern __shared__ double smem_d[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const ScalarLayout L = scalar_layout(a.nd, 0);
  double* base = smem_d + (size_t)wid * L.doubles;
  double* S = base + L.o_S;
  double* A = base + L.o_A;
  double* B = base + L.o_B;
  double* wq = base + L.o_c;
  int* idx = reinterpret_cast<int*>(base + L.o_idx);
  int* lines = idx + MAX_ND + 1;
  const int nd = a.nd, nsym = nd * (nd + 1) / 2;
  for (long long ci = (long long)blockIdx.x * nwarps + wid; ci < a.n_cells; ci += (long long)gridDim.x * nwarps) {
    const long long cell = a.cell_list ? a.cell_list[ci] : ci;
    const double* g = a.geom + cell * a.gstride;
    if (lane < nd) idx[lane] = a.l2g[cell * nd + lane];
    const unsigned short* tp = a.tpos + cell * (long long)(nd * nd);
    if (DIM == 3 && nd == 8 && a.nq <= 28) {
      // Q1 temperature: the two 8x8 local matrices are Gram matrices over the quadrature points -> 28 DMMA per cell
      // instead of 36 entries x 27 points x 9 shared-memory loads (the shared-memory pipe bounded the loop below)
      if (lane < 28) {
        if (lane < a.nq) {
          map_point<DIM>(a, g, lane, S + lane * L.sv);
          wq[lane] = g[lane];
        } else {
          for (int k = 0; k < 32; ++k) S[lane * L.sv + k] = 0.0;
          wq[lane] = 0.0;
        }
      }
      __syncwarp();
      const int frow = lane >> 2, fk = lane & 3;
      double m0 = 0.0, m1 = 0.0, k0 = 0.0, k1 = 0.0;
#pragma unroll
      for (int ks = 0; ks < 7; ++ks) {
        const int q = 4 * ks + fk;
        const double* sp = S + q * L.sv + frow * 4;
        const double w = wq[q], x0 = sp[0], x1 = sp[1], x2 = sp[2], x3 = sp[3];
        dmma_m8n8k4(m0, m1, w * x0, x0);
        dmma_m8n8k4(k0, k1, w * x1, x1);
        dmma_m8n8k4(k0, k1, w * x2, x2);
        dmma_m8n8k4(k0, k1, w * x3, x3);
      }
      k0 *= a.prm.inv_pe;
      k1 *= a.prm.inv_pe;
      const int e0 = frow * 8 + 2 * fk;
      if (tp[0] != 0xffffu) {
        const long long r0 = Mass.rowptr[0][0][idx[frow]];
        const long long at0 = r0 + tp[e0], at1 = r0 + tp[e0 + 1];
        red_add_f64(Mass.val[0][0] + at0, m0);
        red_add_f64(Mass.val[0][0] + at1, m1);
        red_add_f64(Stiff.val[0][0] + at0, k0);
        red_add_f64(Stiff.val[0][0] + at1, k1);
      } else {
        A[e0] = m0;
        A[e0 + 1] = m1;
        B[e0] = k0;
        B[e0 + 1] = k1;
        __syncwarp();
        distribute_local_matrix<true>(cs, nd, nd, A, nullptr, idx, lines, Mass, nullptr, lane, 32, false, err);
        __syncwarp();
        distribute_local_matrix<true>(cs, nd, nd, B, nullptr, idx, lines, Stiff, nullptr, lane, 32, false, err);
      }
      __syncwarp();
      continue;
    }
    for (int q0 = 0; q0 < a.nq; q0 += 32) {
      const int q = q0 + lane;
      if (q < a.nq) {
        map_point<DIM>(a, g, q, S + lane * L.sv);
        wq[lane] = g[q];
      }
      __syncwarp();
      const int nqc = min(32, a.nq - q0);
      for (int e = lane; e < nsym; e += 32) {
        int i = 0, r = e;
        while (r >= nd - i) { r -= nd - i; ++i; }
        const int j = i + r;
        double m = 0.0, k = 0.0;
        for (int p = 0; p < nqc; ++p) {
          const double* si = S + p * L.sv + i * 4;
          const double* sj = S + p * L.sv + j * 4;
          const double w = wq[p];
          m += w * si[0] * sj[0];
          k += w * (si[1] * sj[1] + si[2] * sj[2] + si[3] * sj[3]);
        }
        k *= a.prm.inv_pe;
        if (q0 == 0) {
          A[i * nd + j] = m;
          B[i * nd + j] = k;
        } else {
          A[i * nd + j] += m;
          B[i * nd + j] += k;
        }
      }
      __syncwarp();
    }
    for (int e = lane; e < nd * nd; e += 32) {
      const int i = e / nd, j = e - i * nd;
      if (i > j) {
        A[e] = A[j * nd + i];
        B[e] = B[j * nd + i];
      }
    }
    __syncwarp();
    if (tp[0] != 0xffffu) {
      // no constrained dof in this cell: both matrices share one pattern, positions were found once per mesh
      const long long* rp = Mass.rowptr[0][0];
      for (int e = lane; e < nd * nd; e += 32) {
        const long long at = rp[idx[e / nd]] + tp[e];
        red_add_f64(Mass.val[0][0] + at, A[e]);
        red_add_f64(Stiff.val[0][0] + at, B[e]);
      }
    } else {
      distribute_local_matrix<true>(cs, nd, nd, A, nullptr, idx, lines, Mass, nullptr, lane, 32, false, err);
      __syncwarp();
      distribute_local_matrix<true>(cs, nd, nd, B, nullptr, idx, lines, Stiff, nullptr, lane, 32, false, err);
    }
    __syncwarp();
  }
}

// ---- Q1 temperature in 3-D: the two 8x8 Gram matrices on DMMA, with the cell's loads one cell ahead ---------------
// Same arithmetic as the Q1 branch of temperature_matrix_kernel.  That kernel was bound by the latency of its loads
// (ncu: long-scoreboard 6.5 of 13 stall cycles per issue, 17 warps per SM): mapping record -> table -> DMMA -> dof
// indices -> row starts -> reductions form one dependent chain per cell.  Here a warp keeps the next cell's mapping
// data and dof indices in registers while it works on the current one, and fetches the row starts / positions of the
// current cell before the table build.
__global__ void __launch_bounds__(128) temperature_matrix_q1_kernel(ScalarArgs a, CsView cs, BlockView Mass, BlockView Stiff, int* err) {
  extern __shared__ double smem_d[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  constexpr int nd = 8;
  const ScalarLayout L = scalar_layout(nd, 0);
  double* base = smem_d + (size_t)wid * L.doubles;
  double* S = base + L.o_S;
  double* A = base + L.o_A;
  double* B = base + L.o_B;
  double* wq = base + L.o_c;
  int* idx = reinterpret_cast<int*>(base + L.o_idx);
  int* lines = idx + MAX_ND + 1;
  const int frow = lane >> 2, fk = lane & 3, e0 = frow * 8 + 2 * fk, nq = a.nq;
  const long long stride = (long long)gridDim.x * nwarps;
  long long ci = (long long)blockIdx.x * nwarps + wid;
  auto load_cell = [&](long long c, double (&gq)[10], int& iv) {
    const double* g = a.geom + c * a.gstride;
#pragma unroll
    for (int k = 0; k < 10; ++k) gq[k] = lane < nq ? g[nq * k + lane] : 0.0;
    iv = lane < nd ? a.l2g[c * nd + lane] : 0;
  };
  double gc[10];
  int ic = 0;
  if (ci < a.n_cells) load_cell(ci, gc, ic);
  for (; ci < a.n_cells; ci += stride) {
    double gn[10];
    int in = 0;
    if (ci + stride < a.n_cells) load_cell(ci + stride, gn, in);   // in flight while this cell is worked on
    const unsigned short* tp = a.tpos + ci * (long long)(nd * nd);
    const unsigned short t00 = tp[0], tpa = tp[e0], tpb = tp[e0 + 1];
    const long long r0 = Mass.rowptr[0][0][__shfl_sync(0xffffffffu, ic, frow)];
    if (lane < 28) {
      double* row = S + lane * L.sv;
      if (lane < nq) {
        for (int k = 0; k < nd; ++k) {
          const double r0d = __ldg(a.dphiT + (size_t)(k * 3) * nq + lane), r1d = __ldg(a.dphiT + (size_t)(k * 3 + 1) * nq + lane),
                       r2d = __ldg(a.dphiT + (size_t)(k * 3 + 2) * nq + lane);
          row[k * 4] = __ldg(a.phiT + (size_t)k * nq + lane);
#pragma unroll
          for (int d = 0; d < 3; ++d) row[k * 4 + 1 + d] = gc[1 + d] * r0d + gc[4 + d] * r1d + gc[7 + d] * r2d;
        }
        wq[lane] = gc[0];
      } else {
        for (int k = 0; k < 32; ++k) row[k] = 0.0;
        wq[lane] = 0.0;
      }
    }
    __syncwarp();
    double m0 = 0.0, m1 = 0.0, k0 = 0.0, k1 = 0.0;
#pragma unroll
    for (int ks = 0; ks < 7; ++ks) {
      const int q = 4 * ks + fk;
      const double* sp = S + q * L.sv + frow * 4;
      const double w = wq[q], x0 = sp[0], x1 = sp[1], x2 = sp[2], x3 = sp[3];
      dmma_m8n8k4(m0, m1, w * x0, x0);
      dmma_m8n8k4(k0, k1, w * x1, x1);
      dmma_m8n8k4(k0, k1, w * x2, x2);
      dmma_m8n8k4(k0, k1, w * x3, x3);
    }
    k0 *= a.prm.inv_pe;
    k1 *= a.prm.inv_pe;
    if (t00 != 0xffffu) {
      const long long at0 = r0 + tpa, at1 = r0 + tpb;
      red_add_f64(Mass.val[0][0] + at0, m0);
      red_add_f64(Mass.val[0][0] + at1, m1);
      red_add_f64(Stiff.val[0][0] + at0, k0);
      red_add_f64(Stiff.val[0][0] + at1, k1);
    } else {
      if (lane < nd) idx[lane] = ic;
      A[e0] = m0;
      A[e0 + 1] = m1;
      B[e0] = k0;
      B[e0 + 1] = k1;
      __syncwarp();
      distribute_local_matrix<true>(cs, nd, nd, A, nullptr, idx, lines, Mass, nullptr, lane, 32, false, err);
      __syncwarp();
      distribute_local_matrix<true>(cs, nd, nd, B, nullptr, idx, lines, Stiff, nullptr, lane, 32, false, err);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 10; ++k) gc[k] = gn[k];
    ic = in;
  }
}

template <int DIM>
__global__ void __launch_bounds__(128) temperature_rhs_kernel(ScalarArgs a, CsView cs) {
  extern __shared__ double smem_d[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const ScalarLayout L = scalar_layout(a.nd, a.nd_nse);
  double* base = smem_d + (size_t)wid * L.doubles;
  double* S = base + L.o_S;
  double* A = base + L.o_A;
  double* cq = base + L.o_c;  // [32][2]: weight, rhs coefficient
  double* l = base + L.o_l;
  double* T = base + L.o_T;
  double* U = base + L.o_U;
  int* idx = reinterpret_cast<int*>(base + L.o_idx);
  int* lines = idx + MAX_ND + 1;
  const int nd = a.nd, nn = nd * nd;
  const double tau = a.prm.dt / a.prm.nse_interval;
  for (long long it = (long long)blockIdx.x * nwarps + wid; it < a.n_cells; it += (long long)gridDim.x * nwarps) {
    const long long cell = a.cell_list ? a.cell_list[it] : it;
    const double* g = a.geom + cell * a.gstride;
    bool inhom = false;
    if (lane < nd) {
      const int gi = a.l2g[cell * nd + lane];
      idx[lane] = gi;
      T[lane] = a.old_temp[gi];
      l[lane] = 0.0;
      const int li = cs.line_of_dof[gi];
      lines[lane] = li;
      inhom = li >= 0 && cs.inhom[li] != 0.0;
    }
    const bool need_bc = __any_sync(0xffffffffu, inhom);
    for (int k = lane; k < a.nd_nse; k += 32) {
      const double v = a.nse_solution[a.l2g_nse[cell * a.nd_nse + k]];
      if (a.feec)
        U[k] = v;
      else {  // classic: store component-major, U[c * ndu + node]
        const int f = __ldg(a.nse_field + k);
        if (f < DIM) U[f * a.ndu + __ldg(a.nse_base + k)] = v;
      }
    }
    if (need_bc)
      for (int i = lane; i < nn; i += 32) A[i] = 0.0;
    __syncwarp();
    for (int q0 = 0; q0 < a.nq; q0 += 32) {
      const int q = q0 + lane;
      if (q < a.nq) {
        double* row = S + lane * L.sv;
        map_point<DIM>(a, g, q, row);
        double oldT = 0.0, gT[3] = {0.0, 0.0, 0.0}, u[3] = {0.0, 0.0, 0.0};
        for (int k = 0; k < nd; ++k) {
          oldT += T[k] * row[k * 4];
#pragma unroll
          for (int d = 0; d < DIM; ++d) gT[d] += T[k] * row[k * 4 + 1 + d];
        }
        if (a.feec) {
          // u(q) = sum_k U_k J phi_hat_k / det J over the six face dofs (cell dofs 12..17), no face sign
          // (get_function_values, boussineq_model_FEEC.tpp:1039-1040)
          const double det = g[a.nq * 22 + q];
          for (int k = 0; k < 6; ++k) {
            const double* ph = a.phi_u + ((size_t)q * 6 + k) * 3;
            const double p0 = __ldg(ph), p1 = __ldg(ph + 1), p2 = __ldg(ph + 2), Uk = U[12 + k];
#pragma unroll
            for (int d = 0; d < 3; ++d)
              u[d] += Uk * ((g[a.nq * (13 + d * 3) + q] * p0 + g[a.nq * (14 + d * 3) + q] * p1 +
                             g[a.nq * (15 + d * 3) + q] * p2) / det);
          }
        } else {
          for (int n = 0; n < a.ndu; ++n) {
            const double ph = __ldg(a.phi_uT + (size_t)n * a.nq + q);
#pragma unroll
            for (int d = 0; d < DIM; ++d) u[d] += U[d * a.ndu + n] * ph;
          }
        }
        double ugT = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) ugT += u[d] * gT[d];
        const double w = g[q];
        const double gamma = 0.0;  // heat source multiplied by literal 0 in the reference (:922-926)
        cq[lane * 2] = w;
        cq[lane * 2 + 1] = (oldT - tau * ugT - tau * gamma) * w;
      }
      __syncwarp();
      const int nqc = min(32, a.nq - q0);
      if (lane < nd) {
        double s = 0.0;
        for (int p = 0; p < nqc; ++p) s += S[p * L.sv + lane * 4] * cq[p * 2 + 1];
        l[lane] += s;
      }
      if (need_bc)
        for (int e = lane; e < nn; e += 32) {
          const int j = e / nd, i = e - j * nd;
          double s = 0.0;
          for (int p = 0; p < nqc; ++p) {
            const double* si = S + p * L.sv + i * 4;
            const double* sj = S + p * L.sv + j * 4;
            s += cq[p * 2] * (si[0] * sj[0] + tau * a.prm.inv_pe * (si[1] * sj[1] + si[2] * sj[2] + si[3] * sj[3]));
          }
          A[e] += s;
        }
      __syncwarp();
    }
    distribute_local_vector_bc<true>(cs, nd, nd, l, A, idx, lines, a.rhs, lane, 32);
    __syncwarp();
  }
}

// Right-hand side on the cells without inhomogeneously constrained dofs (all but the Dirichlet boundary layers): no
// matrix_for_bc, so nothing has to be shared between the quadrature points except the eight coefficients -- the
// mapped gradients live in registers, the per-warp scratch shrinks from 11 kB to 2 kB and the SM holds twice the
// warps (the kernel is latency-bound: ncu showed 47 % long-scoreboard stalls at 25 % occupancy).
template <int DIM>
__global__ void __launch_bounds__(128, 8) temperature_rhs_plain_kernel(ScalarArgs a, CsView cs) {
  extern __shared__ double smem_d[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int per_warp = 32 + 2 * (MAX_ND + 1) + a.nd_nse + 1 + MAX_ND + 5;
  double* base = smem_d + (size_t)wid * per_warp;
  double* cq = base;                       // 32 rhs coefficients of the current batch of points
  double* l = cq + 32;
  double* T = l + MAX_ND + 1;
  double* U = T + MAX_ND + 1;
  int* idx = reinterpret_cast<int*>(U + a.nd_nse + 1);
  int* lines = idx + MAX_ND + 1;
  const int nd = a.nd;
  const double tau = a.prm.dt / a.prm.nse_interval;
  for (long long cell = (long long)blockIdx.x * nwarps + wid; cell < a.n_cells; cell += (long long)gridDim.x * nwarps) {
    if (a.bc_flag[cell]) continue;
    const double* g = a.geom + cell * a.gstride;
    if (lane < nd) {
      const int gi = a.l2g[cell * nd + lane];
      idx[lane] = gi;
      T[lane] = a.old_temp[gi];
      l[lane] = 0.0;
      lines[lane] = cs.line_of_dof[gi];
    }
    for (int k = lane; k < a.nd_nse; k += 32) {
      const double v = a.nse_solution[a.l2g_nse[cell * a.nd_nse + k]];
      if (a.feec)
        U[k] = v;
      else {
        const int f = __ldg(a.nse_field + k);
        if (f < DIM) U[f * a.ndu + __ldg(a.nse_base + k)] = v;
      }
    }
    __syncwarp();
    for (int q0 = 0; q0 < a.nq; q0 += 32) {
      const int q = q0 + lane;
      if (q < a.nq) {
        double K[DIM][DIM];
#pragma unroll
        for (int e = 0; e < DIM; ++e)
#pragma unroll
          for (int d = 0; d < DIM; ++d) K[e][d] = g[a.nq * (1 + e * DIM + d) + q];
        // grad T = K^T (sum_k T_k grad_ref phi_k): the reference gradient is summed over the nodes first, the mapping is
        // applied once per point (9 multiply-adds per node less than mapping every shape function)
        double oldT = 0.0, gT[3] = {0.0, 0.0, 0.0}, u[3] = {0.0, 0.0, 0.0}, gr[DIM];
#pragma unroll
        for (int e = 0; e < DIM; ++e) gr[e] = 0.0;
        for (int k = 0; k < nd; ++k) {
          const double Tk = T[k];
          oldT += Tk * __ldg(a.phiT + (size_t)k * a.nq + q);
#pragma unroll
          for (int e = 0; e < DIM; ++e) gr[e] += Tk * __ldg(a.dphiT + (size_t)(k * DIM + e) * a.nq + q);
        }
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
          for (int e = 0; e < DIM; ++e) gT[d] += K[e][d] * gr[e];
        if (a.feec) {
          const double det = g[a.nq * 22 + q];
          for (int k = 0; k < 6; ++k) {
            const double* ph = a.phi_u + ((size_t)q * 6 + k) * 3;
            const double p0 = __ldg(ph), p1 = __ldg(ph + 1), p2 = __ldg(ph + 2), Uk = U[12 + k];
#pragma unroll
            for (int d = 0; d < 3; ++d)
              u[d] += Uk * ((g[a.nq * (13 + d * 3) + q] * p0 + g[a.nq * (14 + d * 3) + q] * p1 +
                             g[a.nq * (15 + d * 3) + q] * p2) / det);
          }
        } else {
          for (int n = 0; n < a.ndu; ++n) {
            const double ph = __ldg(a.phi_uT + (size_t)n * a.nq + q);
#pragma unroll
            for (int d = 0; d < DIM; ++d) u[d] += U[d * a.ndu + n] * ph;
          }
        }
        double ugT = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) ugT += u[d] * gT[d];
        const double gamma = 0.0;  // heat source multiplied by literal 0 in the reference (:922-926)
        cq[lane] = (oldT - tau * ugT - tau * gamma) * g[q];
      }
      __syncwarp();
      const int nqc = min(32, a.nq - q0);
      if (lane < nd) {
        double s = 0.0;
        for (int p = 0; p < nqc; ++p) s += __ldg(a.phi + (size_t)(q0 + p) * nd + lane) * cq[p];
        l[lane] += s;
      }
      __syncwarp();
    }
    distribute_local_vector_bc<true>(cs, nd, nd, l, nullptr, idx, lines, a.rhs, lane, 32);
    __syncwarp();
  }
}

// ---- Q1 / Q2 temperature in 3-D, classic family: the plain right-hand side with its two dense contractions on DMMA --
// Same integrand as temperature_rhs_plain_kernel (boussinesq_model.tpp:928-937).  ncu on that kernel (Q1): 56 % of the
// issue slots, of which 29 % interpolate the velocity at the quadrature points (a lane per point looping over the 27 Q2
// nodes) and 19 % contract the point coefficients with the test functions on a few lanes; with runtime sizes its index
// arithmetic costs as much again (Q2: 4 600 warp instructions per cell).  Both are small GEMMs with a constant operand:
//   u[p][c] = sum_n phi_u[p][n] U[n][c]  (4 row tiles of points x 7 k-steps per batch of 32 points)
//   l[i]    = sum_p phi_t[p][i] cq[p]    (row tiles of test functions x 8 k-steps per batch)
// with the constant tables in shared memory once per CTA and compile-time sizes (ND dofs, NQ points per cell).
constexpr int RQ_LD = 28;   // 27 velocity nodes padded to 28 (k-steps of 4)
template <int ND, int NQ>
struct RhsDmma {
  static constexpr int NB = (NQ + 31) / 32;          // batches of 32 points
  static constexpr int NQP = NB * 32;                // padded points
  static constexpr int NT = (ND + 7) / 8;            // row tiles of test functions
  static constexpr int NDP = NT * 8 + 4;             // row stride of the phi_t table: padded test functions + 4 (the fragment
                                                     // loads "lane = (test function, point)" then fall on distinct banks)
  static constexpr int PER_WARP = 32 + 28 + 3 * RQ_LD + 32 * 4 + 28 + 28;   // cq, T, U, su, l, (idx, lines)
  static constexpr size_t smem_bytes(int warps) { return sizeof(double) * (size_t)(NQP * RQ_LD + NQP * NDP + warps * PER_WARP); }
};

template <int ND, int NQ>
__global__ void __launch_bounds__(128, ND == 8 ? 8 : 5) temperature_rhs_plain_dmma_kernel(ScalarArgs a, CsView cs) {
  using C = RhsDmma<ND, NQ>;
  extern __shared__ __align__(16) double smem_d[];
  constexpr int NDU = 27;
  double* tabU = smem_d;                      // [NQP points][28 nodes] phi_u, zero padded
  double* tabT = tabU + C::NQP * RQ_LD;       // [NQP points][NDP] phi_t, zero padded
  double* wbase = tabT + C::NQP * C::NDP;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  double* cq = wbase + (size_t)wid * C::PER_WARP;   // [32] rhs coefficient of the batch's points
  double* T = cq + 32;                              // [28]
  double* U = T + 28;                               // [3][28] velocity, component-major, entry 27 zero
  double* su = U + 3 * RQ_LD;                       // [32 points][4] velocity at the batch's points
  double* l = su + 32 * 4;                          // [28]
  int* idx = reinterpret_cast<int*>(l + 28);        // [28]
  int* lines = idx + 28;                            // [28]
  const int frow = lane >> 2, fk = lane & 3;
  for (int i = threadIdx.x; i < C::NQP * RQ_LD; i += blockDim.x) {
    const int pnt = i / RQ_LD, n = i - pnt * RQ_LD;
    tabU[i] = pnt < NQ && n < NDU ? __ldg(a.phi_uT + (size_t)n * NQ + pnt) : 0.0;
  }
  for (int i = threadIdx.x; i < C::NQP * C::NDP; i += blockDim.x) {
    const int pnt = i / C::NDP, k = i - pnt * C::NDP;
    tabT[i] = pnt < NQ && k < ND ? __ldg(a.phi + (size_t)pnt * ND + k) : 0.0;
  }
  for (int i = lane; i < 3 * RQ_LD; i += 32) U[i] = 0.0;
  __syncthreads();
  const double tau = a.prm.dt / a.prm.nse_interval;
  for (long long cell = (long long)blockIdx.x * nwarps + wid; cell < a.n_cells; cell += (long long)gridDim.x * nwarps) {
    if (a.bc_flag[cell]) continue;
    const double* g = a.geom + cell * a.gstride;
    if (lane < ND) {
      const int gi = a.l2g[cell * ND + lane];
      idx[lane] = gi;
      T[lane] = a.old_temp[gi];
      lines[lane] = cs.line_of_dof[gi];
    }
    for (int k = lane; k < a.nd_nse; k += 32) {
      const int f = __ldg(a.nse_field + k);
      if (f < 3) U[f * RQ_LD + __ldg(a.nse_base + k)] = a.nse_solution[a.l2g_nse[cell * a.nd_nse + k]];
    }
    double acc[C::NT][2];
#pragma unroll
    for (int t = 0; t < C::NT; ++t) acc[t][0] = acc[t][1] = 0.0;
    __syncwarp();
    double bfr[7];   // the velocity operand of the interpolation: U[c = frow][node 4 ks + fk]
#pragma unroll
    for (int ks = 0; ks < 7; ++ks) bfr[ks] = frow < 3 ? U[frow * RQ_LD + 4 * ks + fk] : 0.0;
#pragma unroll 1
    for (int bt = 0; bt < C::NB; ++bt) {
      const int q = 32 * bt + lane;
      const bool on = q < NQ;
      // the lane's quadrature point: d xi / d x and the weight (in flight during the velocity interpolation)
      double K[3][3], w = 0.0;
#pragma unroll
      for (int e = 0; e < 3; ++e)
#pragma unroll
        for (int d = 0; d < 3; ++d) K[e][d] = on ? g[NQ * (1 + e * 3 + d) + q] : 0.0;
      if (on) w = g[q];
      // u[p][c] = sum_n phi_u[p][n] U[n][c] for the 32 points of the batch
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        double c0 = 0.0, c1 = 0.0;
        const double* ta = tabU + (size_t)(32 * bt + 8 * t + frow) * RQ_LD + fk;
#pragma unroll
        for (int ks = 0; ks < 7; ++ks) dmma_m8n8k4(c0, c1, ta[4 * ks], bfr[ks]);
        if (fk < 2) {
          su[(8 * t + frow) * 4 + 2 * fk] = c0;
          su[(8 * t + frow) * 4 + 2 * fk + 1] = c1;
        }
      }
      __syncwarp();
      {
        // grad T = K^T (sum_k T_k grad_ref phi_k): reference gradient first, the mapping once per point
        double oldT = 0.0, gr[3] = {0.0, 0.0, 0.0};
        if (on) {
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            const double Tk = T[k];
            oldT += Tk * __ldg(a.phiT + (size_t)k * NQ + q);   // point-fastest copy: coalesced (the shared table is [point][i])
#pragma unroll
            for (int e = 0; e < 3; ++e) gr[e] += Tk * __ldg(a.dphiT + (size_t)(k * 3 + e) * NQ + q);
          }
        }
        double ugT = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) ugT += su[lane * 4 + d] * (K[0][d] * gr[0] + K[1][d] * gr[1] + K[2][d] * gr[2]);
        const double gamma = 0.0;  // heat source multiplied by literal 0 in the reference (:922-926)
        cq[lane] = (oldT - tau * ugT - tau * gamma) * w;   // w = 0 for the padding points
      }
      __syncwarp();
      // l[i] += sum_p phi_t[p][i] cq[p]
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int pl = 4 * ks + fk;
        const double bq = frow == 0 ? cq[pl] : 0.0;
        const double* tt = tabT + (size_t)(32 * bt + pl) * C::NDP + frow;
#pragma unroll
        for (int t = 0; t < C::NT; ++t) dmma_m8n8k4(acc[t][0], acc[t][1], tt[8 * t], bq);
      }
      __syncwarp();
    }
    if (fk == 0) {
#pragma unroll
      for (int t = 0; t < C::NT; ++t)
        if (8 * t + frow < ND) l[8 * t + frow] = acc[t][0];
    }
    __syncwarp();
    distribute_local_vector_bc<true>(cs, ND, ND, l, nullptr, idx, lines, a.rhs, lane, 32);
    __syncwarp();
  }
}

struct ScalarLaunch {
  int warps;
  size_t smem;
  unsigned grid;
};
ScalarLaunch scalar_launch(dcp_ctx* ctx, long long n_cells, int nd, int nd_nse) {
  const ScalarLayout L = scalar_layout(nd, nd_nse);
  ScalarLaunch s;
  s.warps = nd <= 8 ? 4 : 2;
  s.smem = (size_t)L.doubles * sizeof(double) * s.warps;
  long long b = (n_cells + s.warps - 1) / s.warps;
  long long cap = (long long)ctx->sm_count * 12;
  s.grid = (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
  return s;
}


// reference table of the Q2 kernel in the order its table build reads it: [half][node of the group j][e: 3 gradients,
// value][thread = 4 * point + group]
__global__ void q2_tab_kernel(const double* __restrict__ phi, const double* __restrict__ dphi, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * 7 * 4 * 128) return;
  const int t = i & 127, e = (i >> 7) & 3, j = (i >> 9) % 7, half = i / (7 * 4 * 128);
  const int q = half * 32 + (t >> 2), b = (t & 3) * 7 + j;
  out[i] = b < 27 ? (e < 3 ? dphi[((size_t)q * 27 + b) * 3 + e] : phi[(size_t)q * 27 + b]) : 0.0;
}

// ---- Q2 temperature in 3-D: mass and stiffness matrices on the FP64 tensor cores ---------------------------------
// 27 dofs and 64 quadrature points per cell (QGauss(temperature_degree + 2), boussinesq_model.tpp:834): the symmetric
// DFMA loop above spends 378 entries x 64 points x 8 shared-memory loads per cell.  Here, like the Stokes kernel,
//   X[q][32 alpha + b] = (d_0 phi_b, d_1 phi_b, d_2 phi_b, phi_b)(x_q)            (27 nodes padded to 32)
//   M_ab = sum_q w_q X[q][96 + a] X[q][96 + b],   K_ab = (1/Pe) sum_q w_q sum_e X[q][32 e + a] X[q][32 e + b]
// with mma.sync.m8n8k4.f64: one CTA (4 warps) per cell, the points in two halves of 32 (33 kB operand table), the 10
// node-block pairs (ta <= tb) spread over the warps with both accumulators kept across the halves; lane-local scatter
// through the position row (cells with constrained dofs take the general kernel).
constexpr int TQ_LDB = 132;   // 132 mod 16 == 4: conflict-free fragment loads
constexpr int TQ_HALF = 32;

// operand table of 32 points of one cell: thread = (point, group of 7 nodes)
__device__ __forceinline__ void q2_build_table(const ScalarArgs& a, const double* g, int half, double* X, double* wq, int tid) {
  const int ql = tid >> 2, bg = tid & 3, q = half * TQ_HALF + ql;
  double kinv[3][3];
#pragma unroll
  for (int e = 0; e < 3; ++e)
#pragma unroll
    for (int d = 0; d < 3; ++d) kinv[e][d] = g[a.nq * (1 + 3 * e + d) + q];
  if (bg == 0) wq[ql] = g[q];
  double* x = X + ql * TQ_LDB + bg * 7;
  const int nb = bg == 3 ? 6 : 7;
  const double* tb = a.q2_tab + (size_t)half * (7 * 4 * 128) + tid;   // [half][j][e][thread]: coalesced
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    if (j >= nb) break;
    const double r0 = __ldg(tb + (j * 4) * 128), r1 = __ldg(tb + (j * 4 + 1) * 128), r2 = __ldg(tb + (j * 4 + 2) * 128);
#pragma unroll
    for (int d = 0; d < 3; ++d) x[32 * d + j] = kinv[0][d] * r0 + kinv[1][d] * r1 + kinv[2][d] * r2;
    x[96 + j] = __ldg(tb + (j * 4 + 3) * 128);
  }
}

// The 10 node-block pairs (ta <= tb) of a cell are grouped so that the pairs of one warp share operand fragments:
// warp 0: (0,0) (0,1) (0,2); warp 1: (0,3) (1,3) (2,3); warp 2: (1,1) (1,2); warp 3: (2,2) (3,3) -- 11 fragment loads per
// 10 DMMA and product instead of 20 (the kernel was bound by the LSU pipe: one pair of loads per DMMA).
__constant__ unsigned char c_q2_ta[4][3] = {{0, 0, 0}, {0, 1, 2}, {1, 1, 9}, {2, 3, 9}};
__constant__ unsigned char c_q2_tb[4][3] = {{0, 1, 2}, {3, 3, 3}, {1, 2, 9}, {2, 3, 9}};

template <int W>
__device__ __forceinline__ void q2_mma_half_w(const double* X, const double* wq, int frow, int fk, double (&am)[3][2], double (&ak)[3][2]) {
  constexpr int NT = W < 2 ? 3 : 2;
  constexpr int TA[4][3] = {{0, 0, 0}, {0, 1, 2}, {1, 1, 9}, {2, 3, 9}};
  constexpr int TB[4][3] = {{0, 1, 2}, {3, 3, 3}, {1, 2, 9}, {2, 3, 9}};
  constexpr int T0 = W == 0 ? 0 : (W == 1 ? 0 : (W == 2 ? 1 : 2));   // first operand tile this warp reads
  constexpr int NTILE = W == 0 ? 3 : (W == 1 ? 4 : 2);               // ... and how many consecutive ones
#pragma unroll
  for (int ks = 0; ks < TQ_HALF / 4; ++ks) {
    const int ql = 4 * ks + fk;
    const double* xr = X + ql * TQ_LDB + frow;
    const double wv = wq[ql];
#pragma unroll
    for (int al = 0; al < 4; ++al) {
      double xt[4];
#pragma unroll
      for (int t = 0; t < NTILE; ++t) xt[t] = xr[32 * al + 8 * (T0 + t)];
#pragma unroll
      for (int s = 0; s < NT; ++s) {
        const double af = wv * xt[TA[W][s] - T0], bf = xt[TB[W][s] - T0];
        if (al < 3) dmma_m8n8k4(ak[s][0], ak[s][1], af, bf); else dmma_m8n8k4(am[s][0], am[s][1], af, bf);
      }
    }
  }
}
__device__ __forceinline__ void q2_mma_half(const double* X, const double* wq, int warp, int frow, int fk, double (&am)[3][2], double (&ak)[3][2]) {
  if (warp == 0) q2_mma_half_w<0>(X, wq, frow, fk, am, ak);
  else if (warp == 1) q2_mma_half_w<1>(X, wq, frow, fk, am, ak);
  else if (warp == 2) q2_mma_half_w<2>(X, wq, frow, fk, am, ak);
  else q2_mma_half_w<3>(X, wq, frow, fk, am, ak);
}

__global__ void __launch_bounds__(128, 4) temperature_matrix_q2_kernel(ScalarArgs a, CsView cs, BlockView Mass, BlockView Stiff, int* err) {
  extern __shared__ __align__(16) double smem_d[];
  double* X = smem_d;                      // [32][TQ_LDB]; after the contraction: the two local matrices of a cell with constrained dofs
  double* wq = X + TQ_HALF * TQ_LDB;       // [32]
  int* idx = reinterpret_cast<int*>(wq + TQ_HALF);   // [28]
  int* lines = idx + 28;                             // [28]
  long long* rstart = reinterpret_cast<long long*>(lines + 28);   // [28] row starts of the cell's dofs
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frow = lane >> 2, fk = lane & 3;
  constexpr int ND2 = 27;
  for (int i = tid; i < TQ_HALF * TQ_LDB; i += 128) X[i] = 0.0;   // padding nodes 27..31 stay zero
  const long long* rp = Mass.rowptr[0][0];
  double* vm = Mass.val[0][0];
  double* vk = Stiff.val[0][0];
  for (long long ci = blockIdx.x; ci < a.n_cells; ci += gridDim.x) {
    const long long cell = a.cell_list ? a.cell_list[ci] : ci;
    const double* g = a.geom + cell * a.gstride;
    const unsigned short* tp = a.tpos + cell * (long long)(ND2 * ND2);
    __syncthreads();   // every warp is done with the previous cell
    if (tid < ND2) {
      const int gi = a.l2g[cell * ND2 + tid];
      idx[tid] = gi;
      rstart[tid] = rp[gi];
    }
    // pairs of this warp: c_q2_ta / c_q2_tb
    double am[3][2], ak[3][2];
#pragma unroll
    for (int s = 0; s < 3; ++s) am[s][0] = am[s][1] = ak[s][0] = ak[s][1] = 0.0;
    for (int half = 0; half < 2; ++half) {
      if (half) __syncthreads();
      q2_build_table(a, g, half, X, wq, tid);
      __syncthreads();
      q2_mma_half(X, wq, warp, frow, fk, am, ak);
    }
    if (tp[0] == 0xffffu) {
      // the cell holds constrained dofs: the local matrices go through shared memory (in place of the operand table) and
      // the full distribute_local_to_global
      double* A = X;
      double* B = X + ND2 * ND2 + 1;
      __syncthreads();
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ta = c_q2_ta[warp][s], tb = c_q2_tb[warp][s];
        if (ta < 4) {
          const int na = 8 * ta + frow;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int nb = 8 * tb + 2 * fk + jj;
            if (na >= ND2 || nb >= ND2) continue;
            const double mv = am[s][jj], kv = ak[s][jj] * a.prm.inv_pe;
            A[na * ND2 + nb] = mv;
            B[na * ND2 + nb] = kv;
            A[nb * ND2 + na] = mv;
            B[nb * ND2 + na] = kv;
          }
        }
      }
      __syncthreads();
      distribute_local_matrix<false>(cs, ND2, ND2, A, nullptr, idx, lines, Mass, nullptr, tid, 128, false, err);
      __syncthreads();
      distribute_local_matrix<false>(cs, ND2, ND2, B, nullptr, idx, lines, Stiff, nullptr, tid, 128, false, err);
      __syncthreads();
      for (int i = tid; i < 2 * ND2 * ND2 + 2; i += 128) X[i] = 0.0;   // the padding columns of the table are zero again
      continue;
    }
    // scatter: lane holds (na, nb = 8 tb + 2 fk + jj); the transposed pair comes from the same registers.  All positions
    // are fetched before the first reduction (the reductions order memory: a load behind one would wait for its turn).
    unsigned short pd[3][2], pt[3][2];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int ta = c_q2_ta[warp][s], tb = c_q2_tb[warp][s];
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int na = 8 * ta + frow, nb = 8 * tb + 2 * fk + jj;
        const bool on = ta < 4 && na < ND2 && nb < ND2;
        pd[s][jj] = on ? tp[na * ND2 + nb] : (unsigned short)0;
        pt[s][jj] = on && ta != tb ? tp[nb * ND2 + na] : (unsigned short)0;
      }
    }
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int ta = c_q2_ta[warp][s], tb = c_q2_tb[warp][s];
      if (ta < 4) {
        const int na = 8 * ta + frow;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int nb = 8 * tb + 2 * fk + jj;
          if (na >= ND2 || nb >= ND2) continue;
          const double mv = am[s][jj], kv = ak[s][jj] * a.prm.inv_pe;
          const long long at = rstart[na] + pd[s][jj];
          red_add_f64(vm + at, mv);
          red_add_f64(vk + at, kv);
          if (ta != tb) {
            const long long at2 = rstart[nb] + pt[s][jj];
            red_add_f64(vm + at2, mv);
            red_add_f64(vk + at2, kv);
          }
        }
      }
    }
  }
}

// ---- Q2 temperature in 3-D, classic family: right-hand side of the cells with inhomogeneously constrained dofs ------
// (boussinesq_model.tpp:928-963).  matrix_for_bc = M_local + tau/Pe K_local is the same Gram contraction as the matrices
// above, so it runs on the tensor cores (the DFMA loop of temperature_rhs_kernel spends 729 entries x 64 points x 8
// shared-memory loads per cell -- 1.7 ms for the 6 % boundary cells of the refine-5 shell against 1.5 ms for all the
// others); the local vector comes from the same operand table.
__global__ void __launch_bounds__(128, 4) temperature_rhs_bc_q2_kernel(ScalarArgs a, CsView cs) {
  extern __shared__ __align__(16) double smem_d[];
  double* X = smem_d;                      // [32][TQ_LDB]; after the contraction: matrix_for_bc
  double* wq = X + TQ_HALF * TQ_LDB;       // [32]
  int* idx = reinterpret_cast<int*>(wq + TQ_HALF);   // [28]
  int* lines = idx + 28;                             // [28]
  double* cq = reinterpret_cast<double*>(lines + 28);   // [32] rhs coefficient of the points
  double* sT = cq + 32;                    // [28] old temperature
  double* sU = sT + 28;                    // [3][28] velocity, component-major
  double* sl = sU + 3 * 28;                // [28] local vector
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frow = lane >> 2, fk = lane & 3;
  constexpr int ND2 = 27;
  const double tau = a.prm.dt / a.prm.nse_interval;
  for (int i = tid; i < TQ_HALF * TQ_LDB; i += 128) X[i] = 0.0;   // padding nodes 27..31 stay zero
  for (long long ci = blockIdx.x; ci < a.n_cells; ci += gridDim.x) {
    const long long cell = a.cell_list ? a.cell_list[ci] : ci;
    const double* g = a.geom + cell * a.gstride;
    __syncthreads();   // every warp is done with the previous cell
    if (tid < ND2) {
      const int gi = a.l2g[cell * ND2 + tid];
      idx[tid] = gi;
      sT[tid] = a.old_temp[gi];
      sl[tid] = 0.0;
      lines[tid] = cs.line_of_dof[gi];
    }
    for (int k = tid; k < a.nd_nse; k += 128) {
      const int f = __ldg(a.nse_field + k);
      if (f < 3) sU[f * 28 + __ldg(a.nse_base + k)] = a.nse_solution[a.l2g_nse[cell * a.nd_nse + k]];
    }
    double am[3][2], ak[3][2];
#pragma unroll
    for (int s = 0; s < 3; ++s) am[s][0] = am[s][1] = ak[s][0] = ak[s][1] = 0.0;
    for (int half = 0; half < 2; ++half) {
      if (half) __syncthreads();
      q2_build_table(a, g, half, X, wq, tid);
      __syncthreads();
      {
        // old temperature, its gradient and the velocity at the point: the four threads of a point share the nodes
        const int ql = tid >> 2, part = tid & 3, q = half * TQ_HALF + ql;
        const double* xr = X + ql * TQ_LDB;
        double oldT = 0.0, gT[3] = {0.0, 0.0, 0.0}, u[3] = {0.0, 0.0, 0.0};
        const int k1 = part == 3 ? ND2 : 7 * part + 7;
        for (int k = 7 * part; k < k1; ++k) {
          const double Tk = sT[k], ph = __ldg(a.phi_uT + (size_t)k * a.nq + q);
          oldT += Tk * xr[96 + k];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            gT[d] += Tk * xr[32 * d + k];
            u[d] += sU[d * 28 + k] * ph;
          }
        }
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
          oldT += __shfl_xor_sync(0xffffffffu, oldT, o);
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            gT[d] += __shfl_xor_sync(0xffffffffu, gT[d], o);
            u[d] += __shfl_xor_sync(0xffffffffu, u[d], o);
          }
        }
        const double gamma = 0.0;  // heat source multiplied by literal 0 in the reference (:922-926)
        if (part == 0) cq[ql] = (oldT - tau * (u[0] * gT[0] + u[1] * gT[1] + u[2] * gT[2]) - tau * gamma) * wq[ql];
      }
      q2_mma_half(X, wq, warp, frow, fk, am, ak);
      __syncthreads();
      if (tid < ND2) {
        double s = 0.0;
        for (int p = 0; p < TQ_HALF; ++p) s += X[p * TQ_LDB + 96 + tid] * cq[p];
        sl[tid] += s;
      }
    }
    double* A = X;
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int ta = c_q2_ta[warp][s], tb = c_q2_tb[warp][s];
      if (ta < 4) {
        const int na = 8 * ta + frow;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int nb = 8 * tb + 2 * fk + jj;
          if (na >= ND2 || nb >= ND2) continue;
          const double v = am[s][jj] + tau * a.prm.inv_pe * ak[s][jj];
          A[na * ND2 + nb] = v;
          A[nb * ND2 + na] = v;
        }
      }
    }
    __syncthreads();
    distribute_local_vector_bc<false>(cs, ND2, ND2, sl, A, idx, lines, a.rhs, tid, 128);
    __syncthreads();
    for (int i = tid; i < ND2 * ND2; i += 128) X[i] = 0.0;   // the padding columns of the table are zero again
  }
}

constexpr size_t q2_smem_bytes() { return sizeof(double) * (TQ_HALF * TQ_LDB + TQ_HALF + 32 + 28 + 3 * 28 + 28 + 28) + sizeof(int) * 56; }

int dcp_q2_table(dcp_model* m) {
  if (m->temp_q2_tab) return DCP_OK;
  DCP_CUDA(cudaMalloc((void**)&m->temp_q2_tab, sizeof(double) * 2 * 7 * 4 * 128));
  q2_tab_kernel<<<(2 * 7 * 4 * 128 + 255) / 256, 256, 0, m->ctx->stream>>>(m->phi_t_qt, m->dphi_t_qt, m->temp_q2_tab);
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
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
  a.tpos = m->temp_pos;
  a.phi = m->phi_t_qt;
  a.dphi = m->dphi_t_qt;
  a.phiT = m->phi_t_qt_T;
  a.dphiT = m->dphi_t_qt_T;
  a.phi_uT = m->phi_u_qt_T;
  a.prm = p;
  if (m->n_cells == 0) return DCP_OK;
  if (m->dim == 3 && a.nd == 27 && a.nq == 64 && !std::getenv("DCP_NO_Q2_DMMA")) {
    // Q2 temperature: tensor-core kernel; cells without a position row (constrained dofs) end in the general scatter
    DCP_TRY(dcp_q2_table(m));
    a.q2_tab = m->temp_q2_tab;
    a.cell_list = nullptr;
    const size_t smem = q2_smem_bytes();
    DCP_CUDA(cudaFuncSetAttribute(temperature_matrix_q2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = std::min<long long>((long long)ctx->sm_count * 4, a.n_cells);
    temperature_matrix_q2_kernel<<<(unsigned)grid, 128, smem, ctx->stream>>>(a, make_view(m->temp_cs), make_view(m->tmass), make_view(m->tstiff),
                                                                           ctx->d_err);
    ctx->launches++;
    DCP_CUDA(cudaGetLastError());
    return DCP_OK;
  }
  const ScalarLaunch s = scalar_launch(ctx, m->n_cells, a.nd, 0);
  if (m->dim == 3 && a.nd == 8 && a.nq <= 28 && !std::getenv("DCP_NO_Q1_PREFETCH")) {
    DCP_CUDA(cudaFuncSetAttribute(temperature_matrix_q1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem));
    temperature_matrix_q1_kernel<<<s.grid, 32 * s.warps, s.smem, ctx->stream>>>(a, make_view(m->temp_cs), make_view(m->tmass), make_view(m->tstiff),
                                                                               ctx->d_err);
    ctx->launches++;
    DCP_CUDA(cudaGetLastError());
    return DCP_OK;
  }
  if (m->dim == 3) {
    DCP_CUDA(cudaFuncSetAttribute(temperature_matrix_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem));
    temperature_matrix_kernel<3><<<s.grid, 32 * s.warps, s.smem, ctx->stream>>>(
        a, make_view(m->temp_cs), make_view(m->tmass), make_view(m->tstiff), ctx->d_err);
  } else {
    DCP_CUDA(cudaFuncSetAttribute(temperature_matrix_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem));
    temperature_matrix_kernel<2><<<s.grid, 32 * s.warps, s.smem, ctx->stream>>>(
        a, make_view(m->temp_cs), make_view(m->tmass), make_view(m->tstiff), ctx->d_err);
  }
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
  a.phiT = m->phi_t_qt_T;
  a.dphiT = m->dphi_t_qt_T;
  a.phi_uT = m->phi_u_qt_T;
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
  // (1) all cells without inhomogeneously constrained dofs: plain kernel, small scratch
  {
    a.cell_list = nullptr;
    a.bc_flag = m->temp_bc_flag;
    const int warps = 4;
    const size_t smem = sizeof(double) * warps * (size_t)(32 + 2 * (MAX_ND + 1) + a.nd_nse + 1 + MAX_ND + 5);
    long long b = (m->n_cells + warps - 1) / warps;
    const long long cap = (long long)ctx->sm_count * 16;
    const unsigned grid = (unsigned)(b > cap ? cap : b);
    if (m->dim == 3 && !a.feec && a.ndu == 27 && ((a.nd == 8 && a.nq == 27) || (a.nd == 27 && a.nq == 64)) && !std::getenv("DCP_NO_DMMA_RHS")) {
      if (a.nd == 8) {
        using C = RhsDmma<8, 27>;
        const size_t smem_q = C::smem_bytes(warps);
        DCP_CUDA(cudaFuncSetAttribute(temperature_rhs_plain_dmma_kernel<8, 27>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q));
        const unsigned grid_q = (unsigned)std::min<long long>(b, (long long)ctx->sm_count * 8);
        temperature_rhs_plain_dmma_kernel<8, 27><<<grid_q, 32 * warps, smem_q, ctx->stream>>>(a, make_view(m->temp_cs));
      } else {
        using C = RhsDmma<27, 64>;
        const size_t smem_q = C::smem_bytes(warps);
        DCP_CUDA(cudaFuncSetAttribute(temperature_rhs_plain_dmma_kernel<27, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q));
        const unsigned grid_q = (unsigned)std::min<long long>(b, (long long)ctx->sm_count * 5);
        temperature_rhs_plain_dmma_kernel<27, 64><<<grid_q, 32 * warps, smem_q, ctx->stream>>>(a, make_view(m->temp_cs));
      }
    } else if (m->dim == 3)
      temperature_rhs_plain_kernel<3><<<grid, 32 * warps, smem, ctx->stream>>>(a, make_view(m->temp_cs));
    else
      temperature_rhs_plain_kernel<2><<<grid, 32 * warps, smem, ctx->stream>>>(a, make_view(m->temp_cs));
    ctx->launches++;
    DCP_CUDA(cudaGetLastError());
  }
  // (2) the cells that need matrix_for_bc (boussinesq_model.tpp:939-949)
  if (m->n_temp_bc_cells == 0) return DCP_OK;
  a.cell_list = m->temp_bc_cells;
  a.bc_flag = nullptr;
  a.n_cells = m->n_temp_bc_cells;
  if (m->dim == 3 && !a.feec && a.nd == 27 && a.nq == 64 && a.ndu == 27 && !std::getenv("DCP_NO_Q2_DMMA")) {
    DCP_TRY(dcp_q2_table(m));
    a.q2_tab = m->temp_q2_tab;
    const size_t smem = q2_smem_bytes();
    DCP_CUDA(cudaFuncSetAttribute(temperature_rhs_bc_q2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = std::min<long long>((long long)ctx->sm_count * 4, a.n_cells);
    temperature_rhs_bc_q2_kernel<<<(unsigned)grid, 128, smem, ctx->stream>>>(a, make_view(m->temp_cs));
    ctx->launches++;
    DCP_CUDA(cudaGetLastError());
    return DCP_OK;
  }
  const ScalarLaunch s = scalar_launch(ctx, a.n_cells, a.nd, a.nd_nse);
  if (m->dim == 3) {
    DCP_CUDA(cudaFuncSetAttribute(temperature_rhs_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem));
    temperature_rhs_kernel<3><<<s.grid, 32 * s.warps, s.smem, ctx->stream>>>(a, make_view(m->temp_cs));
  } else {
    DCP_CUDA(cudaFuncSetAttribute(temperature_rhs_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem));
    temperature_rhs_kernel<2><<<s.grid, 32 * s.warps, s.smem, ctx->stream>>>(a, make_view(m->temp_cs));
  }
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}
