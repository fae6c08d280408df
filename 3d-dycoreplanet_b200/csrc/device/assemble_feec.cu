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
// Scatter: cells without constrained dofs add their nonzero entries at `row start + position`, positions found once per
// mesh (19 x 19 uint16 per cell and matrix); the binary searches of the general path cost ~9 dependent global loads
// per entry and dominated the first version of this kernel.  Cells with constrained dofs take the general
// AffineConstraints path of scatter.cuh.
// Preconditioner (quirk Q5, :558-569): JxW multiplies only p*p; dt/Re curl.curl and the sign(u.w) terms are
// summed unweighted over the 8 points.
#include <algorithm>
#include <vector>

#include "scatter.cuh"

namespace {

using namespace dcpdev;

constexpr int NW = 12, NU = 6, ND = 19;
constexpr int PROW = 364;  // plan row: 19 * 19 offsets padded to a multiple of 8 bytes (cp.async granularity)
constexpr int SV = 97;      // preconditioner: values per quadrature point (96) padded to an odd stride
constexpr int SV_SYS = 61;  // system: the larger of the two table halves (36 + 18 + 6) padded to an odd stride
constexpr int NQMAX = 27;
constexpr int FWARPS = 1;   // one warp per CTA: the scratch of one warp (19 kB system, 11 kB preconditioner) is the occupancy unit

struct FeecArgs {
  long long n_cells;
  int nq, ndt, gstride;
  const double* geom;
  const double* sign;
  const int* l2g;
  const int* l2g_t;
  const double *tw, *tc, *tu, *td;  // reference tables on this rule, [q][function][component]
  const double *tw_t, *tc_t, *tu_t;  // system rule, point-fastest copies [function][component][q] (CTA kernel)
  const double* phi_t;
  const double* old_nse;
  const double* old_temp;
  double* rhs;
  const unsigned short* pos;  // [n_cells][19*19] position of entry (i,j) inside row i of its block; pos[cell][0] == 0xfffe: general path
  const int* cell_list;       // warp kernel: the cells to process (nullptr: all)
  dcp_params prm;
};

template <int NQM, int STRIDE>
struct FeecScratch {
  double S[NQM * STRIDE];
  double L[ND * ND];
  double l[ND];
  double F[NQM * 4];    // rhs integrand per point: A[3] (dotted with u_i) and B (times div u_i), JxW folded in
  double wq[NQM];
  double U[ND];
  double sg[ND];
  int idx[ND + 1];
  int lines[ND + 1];
  long long rs[ND * 3];  // start of local row i inside column block bj
  double Tn[28];         // nodal old temperature of the cell
  unsigned short pos[ND * ND + 3];
};

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// NQM: capacity of the per-warp table (27 for the system rule QGauss(3), 8 for the preconditioner rule QGauss(2)).
// Shared memory per warp decides how many warps an SM holds, and this kernel is latency-bound (ncu: 12 % occupancy,
// long-scoreboard + fixed-latency stalls), so the system pass builds its table in two halves that share one buffer:
// first the vorticity values (for the w-w block), then curl / velocity / divergence (for everything else).
template <bool SYSTEM, int NQM>
__global__ void __launch_bounds__(32 * FWARPS) feec_kernel(FeecArgs a, CsView cs, BlockView A, int* err) {
  extern __shared__ __align__(16) unsigned char raw_smem[];
  using Scratch = FeecScratch<NQM, SYSTEM ? SV_SYS : SV>;
  Scratch* all = reinterpret_cast<Scratch*>(raw_smem);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  Scratch& s = all[wid];
  constexpr int ST = SYSTEM ? SV_SYS : SV;                  // row stride of S
  constexpr int P_W = 0;                                    // vorticity values (system: first half)
  constexpr int P_C = SYSTEM ? 0 : 36;                      // curls (system: second half)
  constexpr int P_U = SYSTEM ? 36 : 72;                     // velocity values
  constexpr int P_D = SYSTEM ? 54 : 90;                     // divergences
  const int nq = a.nq;
  const double nu = a.prm.dt * a.prm.inv_re;
  // Local matrix by 3 x 3 register tiles: one lane owns a block of three row functions x three column functions and
  // loads their 18 mapped components once per quadrature point (2 shared-memory loads per entry and point instead
  // of 6; all tiles of a stage run in one round of the warp).
  auto tile = [&](int orow, int ocol, bool weighted, bool sign_only, double acc[3][3]) {
#pragma unroll
    for (int x = 0; x < 3; ++x)
#pragma unroll
      for (int y = 0; y < 3; ++y) acc[x][y] = 0.0;
    for (int q = 0; q < nq; ++q) {
      const double* ra = s.S + q * ST + orow;
      const double* rb = s.S + q * ST + ocol;
      double va[9], vb[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        va[k] = ra[k];
        vb[k] = rb[k];
      }
      const double w = weighted ? s.wq[q] : 1.0;
#pragma unroll
      for (int x = 0; x < 3; ++x)
#pragma unroll
        for (int y = 0; y < 3; ++y) {
          const double dt3 = va[3 * x] * vb[3 * y] + va[3 * x + 1] * vb[3 * y + 1] + va[3 * x + 2] * vb[3 * y + 2];
          if (sign_only)
            acc[x][y] += fabs(dt3) > 1.0e-9 ? (signbit(dt3) ? -1.0 : 1.0) : 0.0;
          else
            acc[x][y] += w * dt3;
        }
    }
  };
  // upper-triangular block pairs (I <= J) of an n x n block grid, t-th pair
  auto sym_pair = [](int t, int n, int& I, int& J) {
    I = 0;
    while (t >= n - I) {
      t -= n - I;
      ++I;
    }
    J = I + t;
  };
  for (long long it = (long long)blockIdx.x * FWARPS + wid; it < a.n_cells; it += (long long)gridDim.x * FWARPS) {
    const long long cell = a.cell_list ? a.cell_list[it] : it;
    const double* g = a.geom + cell * a.gstride;
    if (lane < ND) {
      const int gi = a.l2g[cell * ND + lane];
      s.idx[lane] = gi;
      s.sg[lane] = a.sign[cell * ND + lane];
      if (SYSTEM) s.U[lane] = a.old_nse[gi];
    }
    if (SYSTEM && lane < a.ndt) s.Tn[lane] = a.old_temp[a.l2g_t[cell * a.ndt + lane]];
    {  // the plan row, staged now so that its latency overlaps the quadrature loop
      const unsigned short* pp = a.pos + cell * (long long)PROW;
      for (int i = lane; i < ND * ND; i += 32) s.pos[i] = pp[i];
    }
    for (int i = lane; i < ND * ND; i += 32) s.L[i] = 0.0;
    __syncwarp();
    // ---- mapping at this lane's quadrature point (kept in registers across the two table halves)
    const int q = lane;
    const bool has_q = lane < nq;
    double J[3][3], K[3][3], idet = 0.0, ow[3] = {0, 0, 0}, ou[3] = {0, 0, 0};
    if (has_q) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          K[i][j] = g[nq * (1 + i * 3 + j) + q];
          J[i][j] = g[nq * (13 + i * 3 + j) + q];
        }
      idet = 1.0 / g[nq * 22 + q];
      s.wq[q] = g[q];
      // vorticity functions: covariant map K^T phi_hat
      double* row = s.S + q * ST;
      for (int k = 0; k < NW; ++k) {
        const double* ph = a.tw + ((size_t)q * NW + k) * 3;
        const double p0 = __ldg(ph), p1 = __ldg(ph + 1), p2 = __ldg(ph + 2);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double vw = K[0][d] * p0 + K[1][d] * p1 + K[2][d] * p2;
          row[P_W + k * 3 + d] = vw;
          if (SYSTEM) ow[d] += s.U[k] * vw;
        }
      }
    }
    double acc[3][3];
    if (SYSTEM) {
      __syncwarp();
      if (lane < 10) {  // (w,w) tiles
        int I, Jb;
        sym_pair(lane, 4, I, Jb);
        tile(P_W + 9 * I, P_W + 9 * Jb, true, false, acc);
#pragma unroll
        for (int x = 0; x < 3; ++x)
#pragma unroll
          for (int y = 0; y < 3; ++y) {
            s.L[(3 * I + x) * ND + 3 * Jb + y] = acc[x][y];
            s.L[(3 * Jb + y) * ND + 3 * I + x] = acc[x][y];
          }
      }
      __syncwarp();  // the second half of the table overwrites the vorticity values
    }
    if (has_q) {
      double* row = s.S + q * ST;
      for (int k = 0; k < NW; ++k) {  // curls: J c_hat / det
        const double* ch = a.tc + ((size_t)q * NW + k) * 3;
        const double c0 = __ldg(ch), c1 = __ldg(ch + 1), c2 = __ldg(ch + 2);
#pragma unroll
        for (int d = 0; d < 3; ++d) row[P_C + k * 3 + d] = (J[d][0] * c0 + J[d][1] * c1 + J[d][2] * c2) * idet;
      }
      for (int k = 0; k < NU; ++k) {  // velocity functions: contravariant Piola J phi_hat / det, times the face sign
        const double* ph = a.tu + ((size_t)q * NU + k) * 3;
        const double p0 = __ldg(ph), p1 = __ldg(ph + 1), p2 = __ldg(ph + 2);
        const double sgk = s.sg[NW + k];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double ru = (J[d][0] * p0 + J[d][1] * p1 + J[d][2] * p2) * idet;
          row[P_U + k * 3 + d] = sgk * ru;
          if (SYSTEM) ou[d] += s.U[NW + k] * ru;  // get_function_values: no face sign (:705-708)
        }
        row[P_D + k] = sgk * __ldg(a.td + k) * idet;
      }
      if (SYSTEM) {
        const double w = s.wq[q];
        double T = 0.0;
        for (int k = 0; k < a.ndt; ++k) T += s.Tn[k] * __ldg(a.phi_t + q * a.ndt + k);
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
      // lanes 0-7: (curl w,u) tiles, 8-10: (u,u); lane 11: the 6 (u,p) entries; lanes 12-17: right-hand side
      if (lane < 11) {
        const bool cu = lane < 8;
        int I, Jb;
        if (cu) {
          I = lane >> 1;
          Jb = lane & 1;
        } else
          sym_pair(lane - 8, 2, I, Jb);
        tile((cu ? P_C : P_U) + 9 * I, P_U + 9 * Jb, true, false, acc);
        const int r0 = (cu ? 0 : NW) + 3 * I, c0 = NW + 3 * Jb;
#pragma unroll
        for (int x = 0; x < 3; ++x)
#pragma unroll
          for (int y = 0; y < 3; ++y) {
            const double v = acc[x][y];
            s.L[(r0 + x) * ND + c0 + y] = cu ? -v : v;
            s.L[(c0 + y) * ND + r0 + x] = cu ? nu * v : v;
          }
      } else if (lane == 11) {
        double d[NU] = {0, 0, 0, 0, 0, 0};
        for (int p = 0; p < nq; ++p)
#pragma unroll
          for (int i = 0; i < NU; ++i) d[i] += s.wq[p] * s.S[p * ST + P_D + i];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          s.L[(NW + i) * ND + NW + NU] = -d[i];
          s.L[(NW + NU) * ND + NW + i] = -d[i];
        }
      } else if (lane < 12 + NU) {
        const int i = lane - 12;
        double v = 0.0;
        for (int p = 0; p < nq; ++p)
          v += dot3(s.S + p * ST + P_U + i * 3, s.F + p * 4) + s.S[p * ST + P_D + i] * s.F[p * 4 + 3];
        s.l[NW + i] = v;
      }
      if (lane >= 20 && lane < 20 + NW) s.l[lane - 20] = 0.0;
      if (lane == 19) s.l[NW + NU] = 0.0;
    } else {
      // preconditioner: lanes 0-9 curl-curl tiles, 10-17 sign tiles (u rows x w columns), lane 18 the pressure mass
      int I = 0, Jb = 0;
      const bool cc = lane < 10;
      if (cc)
        sym_pair(lane, 4, I, Jb);
      else {
        I = (lane - 10) >> 2;
        Jb = (lane - 10) & 3;
      }
      if (lane < 18) {
        tile((cc ? P_C : P_U) + 9 * I, (cc ? P_C : P_W) + 9 * Jb, false, !cc, acc);
        const int r0 = (cc ? 0 : NW) + 3 * I, c0 = 3 * Jb;
#pragma unroll
        for (int x = 0; x < 3; ++x)
#pragma unroll
          for (int y = 0; y < 3; ++y) {
            const double v = cc ? nu * acc[x][y] : acc[x][y];
            s.L[(r0 + x) * ND + c0 + y] = v;
            s.L[(c0 + y) * ND + r0 + x] = v;
          }
      } else if (lane == 18) {
        double v = 0.0;
        for (int p = 0; p < nq; ++p) v += s.wq[p];
        s.L[(NW + NU) * ND + NW + NU] = v;
      }
    }
    __syncwarp();
    const unsigned short* pp = s.pos;
    if (pp[0] != 0xfffeu) {
      for (int t = lane; t < ND * 3; t += 32) {
        const int i = t / 3, bj = t - 3 * i, bi = i < NW ? 0 : (i < NW + NU ? 1 : 2);
        const long long* rp = A.rowptr[bi][bj];
        s.rs[t] = rp ? rp[s.idx[i] - A.start[bi]] : 0;
      }
      __syncwarp();
      for (int e = lane; e < ND * ND; e += 32) {
        const double v = s.L[e];
        if (v == 0.0) continue;  // distribute_local_to_global elides exact zeros
        const int i = e / ND, j = e - i * ND;
        const int bi = i < NW ? 0 : (i < NW + NU ? 1 : 2), bj = j < NW ? 0 : (j < NW + NU ? 1 : 2);
        const unsigned o = pp[e];
        if (o == 0xffffu)
          atomicAdd(err, 1);  // entry missing from the sparsity pattern
        else
          red_add_f64(A.val[bi][bj] + s.rs[i * 3 + bj] + o, v);
      }
      if (SYSTEM && lane < ND) red_add_f64(a.rhs + s.idx[lane], s.l[lane]);
    } else {
      distribute_local_matrix<true>(cs, ND, ND, s.L, SYSTEM ? s.l : nullptr, s.idx, s.lines, A, SYSTEM ? a.rhs : nullptr,
                                    lane, 32, false, err);
    }
    __syncwarp();
  }
}


// ---- system pass, one CTA (4 warps) per cell -------------------------------------------------------------------------
// The warp-per-cell kernel above is a chain of dependent latencies (ncu: >90 % of the cycles stalled at 11 warps per
// SM).  Four warps per cell shorten every link of the chain -- the table is built by (point, function group) threads,
// every 3 x 3 tile by four threads that split the quadrature points -- and 27 kB of scratch per CTA put 32 warps on an SM.
// Cells with constrained dofs are skipped here and handled by the warp kernel over a cell list.
constexpr int CT = 128;
constexpr int SWS = 37;   // row stride of the vorticity table (36 values)
struct FeecCta {
  double G[2][NQMAX * 23];     // mapping records of the current and the next cell (cp.async double buffer)
  double SGN[2][ND + 1];
  double SW[NQMAX * SWS];
  double SR[NQMAX * SV_SYS];   // curls (36), velocity values (18), divergences (6)
  double L[ND * ND];
  double l[ND];
  double F[NQMAX * 4];
  double wq[NQMAX];
  double ow[NQMAX * 4 * 3];    // partial vorticity at the points, per function group
  double ou[NQMAX * 2 * 3];    // partial velocity at the points
  double U[ND];
  double Tn[28];
  long long rs[ND * 3];
  unsigned short P[2][PROW];   // plan rows
  int IDX[2][ND + 1];
};

__global__ void __launch_bounds__(CT) feec_system_cta_kernel(FeecArgs a, BlockView A, int* err) {
  extern __shared__ __align__(16) unsigned char raw_smem[];
  FeecCta& s = *reinterpret_cast<FeecCta*>(raw_smem);
  const int t = threadIdx.x;
  const int nq = a.nq;
  const double nu = a.prm.dt * a.prm.inv_re;
  // the inputs of the next cell travel while the current one is processed (ncu of the unpipelined version: 29 % of the
  // stall samples waited for exactly these loads, another 20 % at the barriers behind them)
  auto issue = [&](int buf, long long cell) {
    const double* g = a.geom + cell * a.gstride;
    for (int i = t; i < a.gstride; i += CT) cp_async8(&s.G[buf][i], g + i);
    const unsigned short* pp = a.pos + cell * (long long)PROW;
    for (int i = t; i < PROW / 4; i += CT) cp_async8(&s.P[buf][4 * i], pp + 4 * i);
    if (t < ND) {
      cp_async4(&s.IDX[buf][t], a.l2g + cell * ND + t);
      cp_async8(&s.SGN[buf][t], a.sign + cell * ND + t);
    }
  };
  if ((long long)blockIdx.x < a.n_cells) issue(0, a.cell_list[blockIdx.x]);
  cp_async_commit();
  int cur = 0;
  for (long long it = blockIdx.x; it < a.n_cells; it += gridDim.x, cur ^= 1) {
    const long long cell = a.cell_list[it];
    cp_async_wait<0>();
    __syncthreads();  // this cell's inputs have landed; the previous cell's scatter is done with the scratch
    if (it + gridDim.x < a.n_cells) issue(cur ^ 1, a.cell_list[it + gridDim.x]);
    cp_async_commit();
    const double* g = s.G[cur];
    const double* sgn = s.SGN[cur];
    const int* idx = s.IDX[cur];
    const unsigned short* pos = s.P[cur];
    if (t < ND)
      s.U[t] = a.old_nse[idx[t]];
    else if (t >= 32 && t < 32 + a.ndt)
      s.Tn[t - 32] = a.old_temp[a.l2g_t[cell * a.ndt + t - 32]];
    for (int i = t; i < ND * ND; i += CT) s.L[i] = 0.0;
    __syncthreads();
    // ---- tables: thread = (point q, group of three functions)
    if (t < 4 * nq) {
      const int q = t % nq, grp = t / nq;
      double J[3][3], K[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          K[i][j] = g[nq * (1 + i * 3 + j) + q];
          J[i][j] = g[nq * (13 + i * 3 + j) + q];
        }
      const double idet = 1.0 / g[nq * 22 + q];
      if (grp == 0) s.wq[q] = g[q];
      double ow[3] = {0, 0, 0};
      for (int kk = 0; kk < 3; ++kk) {
        const int k = 3 * grp + kk;
        const double* ph = a.tw_t + (size_t)(k * 3) * nq + q;
        const double* ch = a.tc_t + (size_t)(k * 3) * nq + q;
        const double p0 = __ldg(ph), p1 = __ldg(ph + nq), p2 = __ldg(ph + 2 * nq);
        const double c0 = __ldg(ch), c1 = __ldg(ch + nq), c2 = __ldg(ch + 2 * nq);
        const double Uk = s.U[k];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double vw = K[0][d] * p0 + K[1][d] * p1 + K[2][d] * p2;
          s.SW[q * SWS + k * 3 + d] = vw;
          s.SR[q * SV_SYS + k * 3 + d] = (J[d][0] * c0 + J[d][1] * c1 + J[d][2] * c2) * idet;
          ow[d] += Uk * vw;
        }
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) s.ow[(q * 4 + grp) * 3 + d] = ow[d];
      if (grp < 2) {  // velocity functions 3 grp .. 3 grp + 2 and their divergences
        double ou[3] = {0, 0, 0};
        for (int kk = 0; kk < 3; ++kk) {
          const int k = 3 * grp + kk;
          const double* ph = a.tu_t + (size_t)(k * 3) * nq + q;
          const double p0 = __ldg(ph), p1 = __ldg(ph + nq), p2 = __ldg(ph + 2 * nq);
          const double sgk = sgn[NW + k], Uk = s.U[NW + k];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const double ru = (J[d][0] * p0 + J[d][1] * p1 + J[d][2] * p2) * idet;
            s.SR[q * SV_SYS + 36 + k * 3 + d] = sgk * ru;
            ou[d] += Uk * ru;  // get_function_values: no face sign (:705-708)
          }
          s.SR[q * SV_SYS + 54 + k] = sgk * __ldg(a.td + k) * idet;
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) s.ou[(q * 2 + grp) * 3 + d] = ou[d];
      }
    }
    __syncthreads();
    // ---- warp 0: right-hand-side integrand at the points; warps 1-3: the 21 tiles (four threads each) and the six
    // (u,p) entries
    if (t < 32) {
      if (t < nq) {
        const int q = t;
        double ow[3], ou[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          ow[d] = s.ow[(q * 4) * 3 + d] + s.ow[(q * 4 + 1) * 3 + d] + s.ow[(q * 4 + 2) * 3 + d] + s.ow[(q * 4 + 3) * 3 + d];
          ou[d] = s.ou[(q * 2) * 3 + d] + s.ou[(q * 2 + 1) * 3 + d];
        }
        const double w = s.wq[q];
        double T = 0.0;
        for (int k = 0; k < a.ndt; ++k) T += s.Tn[k] * __ldg(a.phi_t + q * a.ndt + k);
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
    } else {
      const int u = t - 32;            // 0 .. 95
      const int tl = u >> 2, split = u & 3;
      const bool is_tile = tl < 21;
      // tile kinds: 0-9 (w,w), 10-17 (curl w,u), 18-20 (u,u)
      int I = 0, Jb = 0, kind = 0;
      if (tl < 10) {
        int r = tl;
        while (r >= 4 - I) { r -= 4 - I; ++I; }
        Jb = I + r;
      } else if (tl < 18) {
        kind = 1;
        I = (tl - 10) >> 1;
        Jb = (tl - 10) & 1;
      } else if (is_tile) {
        kind = 2;
        int r = tl - 18;
        while (r >= 2 - I) { r -= 2 - I; ++I; }
        Jb = I + r;
      }
      double acc[3][3];
#pragma unroll
      for (int x = 0; x < 3; ++x)
#pragma unroll
        for (int y = 0; y < 3; ++y) acc[x][y] = 0.0;
      if (is_tile) {
        const double* ta = kind == 0 ? s.SW + 9 * I : (kind == 1 ? s.SR + 9 * I : s.SR + 36 + 9 * I);
        const double* tb = kind == 0 ? s.SW + 9 * Jb : s.SR + 36 + 9 * Jb;
        const int st = kind == 0 ? SWS : SV_SYS;
        const int q0 = split * 7, q1 = min(nq, q0 + 7);
        for (int q = q0; q < q1; ++q) {
          double va[9], vb[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            va[k] = ta[q * st + k];
            vb[k] = tb[q * st + k];
          }
          const double w = s.wq[q];
#pragma unroll
          for (int x = 0; x < 3; ++x)
#pragma unroll
            for (int y = 0; y < 3; ++y)
              acc[x][y] += w * (va[3 * x] * vb[3 * y] + va[3 * x + 1] * vb[3 * y + 1] + va[3 * x + 2] * vb[3 * y + 2]);
        }
      }
      // the four threads of a tile hold partial sums over their points (all lanes of the warp take part)
#pragma unroll
      for (int x = 0; x < 3; ++x)
#pragma unroll
        for (int y = 0; y < 3; ++y) {
          double v = acc[x][y];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          acc[x][y] = v;
        }
      if (is_tile && split == 0) {
        const int r0 = (kind == 2 ? NW : 0) + 3 * I, c0 = (kind == 0 ? 0 : NW) + 3 * Jb;
#pragma unroll
        for (int x = 0; x < 3; ++x)
#pragma unroll
          for (int y = 0; y < 3; ++y) {
            const double v = acc[x][y];
            s.L[(r0 + x) * ND + c0 + y] = kind == 1 ? -v : v;
            s.L[(c0 + y) * ND + r0 + x] = kind == 1 ? nu * v : v;
          }
      }
      if (!is_tile && u - 84 < NU) {   // the six (u,p) entries
        const int i = u - 84;
        double dsum = 0.0;
        for (int q = 0; q < nq; ++q) dsum += s.wq[q] * s.SR[q * SV_SYS + 54 + i];
        s.L[(NW + i) * ND + NW + NU] = -dsum;
        s.L[(NW + NU) * ND + NW + i] = -dsum;
      }
    }
    __syncthreads();
    // ---- right-hand side (needs F) and the row starts
    if (t < NU) {
      double v = 0.0;
      for (int q = 0; q < nq; ++q)
        v += dot3(s.SR + q * SV_SYS + 36 + t * 3, s.F + q * 4) + s.SR[q * SV_SYS + 54 + t] * s.F[q * 4 + 3];
      s.l[NW + t] = v;
    } else if (t >= 32 && t < 32 + ND * 3) {
      const int tt = t - 32, i = tt / 3, bj = tt - 3 * i, bi = i < NW ? 0 : (i < NW + NU ? 1 : 2);
      const long long* rp = A.rowptr[bi][bj];
      s.rs[tt] = rp ? rp[idx[i] - A.start[bi]] : 0;
    }
    __syncthreads();
    for (int e = t; e < ND * ND; e += CT) {
      const double v = s.L[e];
      if (v == 0.0) continue;  // distribute_local_to_global elides exact zeros
      const int i = e / ND, j = e - i * ND;
      const int bi = i < NW ? 0 : (i < NW + NU ? 1 : 2), bj = j < NW ? 0 : (j < NW + NU ? 1 : 2);
      const unsigned o = pos[e];
      if (o == 0xffffu)
        atomicAdd(err, 1);  // entry missing from the sparsity pattern
      else
        red_add_f64(A.val[bi][bj] + s.rs[i * 3 + bj] + o, v);
    }
    if (t >= NW && t < NW + NU) red_add_f64(a.rhs + idx[t], s.l[t]);  // w and p rows of the right-hand side are zero
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
  a.tw_t = m->feec_w_qn_t;
  a.tc_t = m->feec_c_qn_t;
  a.tu_t = m->feec_u_qn_t;
  a.phi_t = m->phi_t_qn;
  a.old_nse = old_nse;
  a.old_temp = old_temp;
  a.rhs = system ? m->nse_rhs : nullptr;
  a.pos = system ? m->feec_pos_nse : m->feec_pos_pre;
  a.cell_list = nullptr;
  a.prm = p;
  if (!a.pos) {
    dcp_set_error("FEEC: scatter positions missing (model not fully created)");
    return DCP_ERR_STATE;
  }
  if (a.nq > NQMAX) {
    dcp_set_error("FEEC: more than 27 quadrature points per cell is not supported");
    return DCP_ERR_ARG;
  }
  if (m->n_cells == 0) return DCP_OK;
  const BlockMat& mat = system ? m->nse : m->pre;
  auto launch = [&](auto kernel, size_t smem) -> int {
    DCP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * FWARPS, smem));
    if (per_sm < 1) per_sm = 1;
    long long blocks = (a.n_cells + FWARPS - 1) / FWARPS;
    const long long cap = (long long)ctx->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    kernel<<<(unsigned)blocks, 32 * FWARPS, smem, ctx->stream>>>(a, make_view(m->nse_cs), make_view(mat), ctx->d_err);
    return DCP_OK;
  };
  if (system) {
    // cells without constrained dofs: one CTA per cell; the rest: warp kernel over the list
    {
      const size_t smem = sizeof(FeecCta);
      DCP_CUDA(cudaFuncSetAttribute(feec_system_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 1;
      DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, feec_system_cta_kernel, CT, smem));
      if (per_sm < 1) per_sm = 1;
      long long blocks = m->n_feec_fast;
      const long long cap = (long long)ctx->sm_count * per_sm;
      if (blocks > cap) blocks = cap;
      if (blocks > 0) {
        a.cell_list = m->feec_fast_cells;
        a.n_cells = m->n_feec_fast;
        feec_system_cta_kernel<<<(unsigned)blocks, CT, smem, ctx->stream>>>(a, make_view(mat), ctx->d_err);
        ctx->launches++;
        DCP_CUDA(cudaGetLastError());
      }
    }
    if (m->n_feec_general == 0) return DCP_OK;
    a.cell_list = m->feec_general_cells;
    a.n_cells = m->n_feec_general;
    DCP_TRY(launch(feec_kernel<true, NQMAX>, sizeof(FeecScratch<NQMAX, SV_SYS>) * FWARPS));
  } else if (a.nq <= 8)
    DCP_TRY(launch(feec_kernel<false, 8>, sizeof(FeecScratch<8, SV>) * FWARPS));
  else
    DCP_TRY(launch(feec_kernel<false, NQMAX>, sizeof(FeecScratch<NQMAX, SV>) * FWARPS));
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}

// Positions of the 19 x 19 local entries inside their CSR rows, for the cells without constrained dofs.
int dcp_feec_positions_build(dcp_model* m, const dcp_model_desc* d, bool system, uint16_t** out) {
  *out = nullptr;
  const int64_t nc = d->n_cells;
  const dcp_csr_desc(*pat)[DCP_MAX_BLOCKS] = system ? d->nse_pattern : d->pre_pattern;
  int64_t start[4] = {0, d->nse_block_size[0], d->nse_block_size[0] + d->nse_block_size[1],
                      d->nse_block_size[0] + d->nse_block_size[1] + d->nse_block_size[2]};
  std::vector<int32_t> lod((size_t)d->nse_cs.n_dofs, -1);
  for (int64_t l = 0; l < d->nse_cs.n_lines; ++l) lod[d->nse_cs.line_dof[l]] = (int32_t)l;
  std::vector<uint16_t> pos((size_t)nc * PROW, 0xffff);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nc; ++c) {
    const int32_t* idx = d->nse_l2g + c * ND;
    uint16_t* P = &pos[(size_t)c * PROW];
    bool fast = true;
    for (int i = 0; i < ND && fast; ++i) {
      const int bi = i < NW ? 0 : (i < NW + NU ? 1 : 2);
      fast = lod[idx[i]] < 0 && idx[i] >= start[bi] && idx[i] < start[bi + 1];
    }
    for (int i = 0; i < ND && fast; ++i) {
      const int bi = i < NW ? 0 : (i < NW + NU ? 1 : 2);
      for (int j = 0; j < ND; ++j) {
        const int bj = j < NW ? 0 : (j < NW + NU ? 1 : 2);
        const dcp_csr_desc& B = pat[bi][bj];
        if (!B.rowptr) continue;
        const int64_t r = idx[i] - start[bi];
        const int32_t col = (int32_t)(idx[j] - start[bj]);
        const int32_t* b = B.col + B.rowptr[r];
        const int32_t* e = B.col + B.rowptr[r + 1];
        const int32_t* p = std::lower_bound(b, e, col);
        if (p != e && *p == col && p - b < 0xfffe) P[i * ND + j] = (uint16_t)(p - b);
        if (p != e && *p == col && p - b >= 0xfffe) fast = false;  // row too long for 16-bit offsets
      }
    }
    if (!fast) P[0] = 0xfffe;
  }
  int rc = dcp_upload(m->ctx, out, pos.data(), (int64_t)pos.size());
  if (rc == DCP_OK && system) {
    std::vector<int32_t> general, fast;
    for (int64_t c = 0; c < nc; ++c) (pos[(size_t)c * PROW] == 0xfffe ? general : fast).push_back((int32_t)c);
    m->n_feec_general = (int64_t)general.size();
    m->n_feec_fast = (int64_t)fast.size();
    if (!general.empty()) rc = dcp_upload(m->ctx, &m->feec_general_cells, general.data(), (int64_t)general.size());
    if (rc == DCP_OK && !fast.empty()) rc = dcp_upload(m->ctx, &m->feec_fast_cells, fast.data(), (int64_t)fast.size());
  }
  cudaStreamSynchronize(m->ctx->stream);
  return rc;
}
