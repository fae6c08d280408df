// Classic 3-D Taylor-Hood NSE system: write-once assembly (DCP_STRATEGY_POSITIONS when every cell is a plan cell).
//
// Same integrals and the same tensor-core contraction as assemble_th_mma.cu (reference:
// include/core/boussinesq_model.tpp:550-687), but the scatter of copy_local_to_global_nse_system (:677-687) is split
// in two so that no CSR value is ever reduced with red.global.add.f64 and the matrix needs no zero-fill:
//
//   1. th_stage_kernel: per cell, the constraint-resolved 3x3 blocks L[(a,.),(b,.)] and the velocity-pressure
//      coupling leave the DMMA registers through a per-warp shared-memory transposition and are written, coalesced,
//      to a cell-major staging record (7 892 doubles: 27 velocity-node rows [r = 3c+d][b] + [c][p], 8 pressure-node
//      rows [c][b]).  The symmetric half is computed, both orientations are staged.
//   2. th_gather_kernel: one warp per (chunk, node).  It sums the staged rows of the node's cells inside the chunk in
//      shared-memory accumulators at the plan's row positions and writes the node's CSR rows once (first chunk that
//      touches the node: plain store of the whole row; later chunks: read-modify-write of the touched entries; the
//      chunks are stream-ordered, rows are owned by one warp, so no atomics are needed).
//
// The cells are processed in chunks of the plan order; a chunk's staging is consumed before the next chunk overwrites
// it.  Right-hand side: as in assemble_th_mma.cu (reductions into nse_rhs, 81 per cell).
#include <omp.h>

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>

#include "th_mma_common.cuh"

namespace {

using namespace dcpdev;
using namespace thmma;

constexpr int BW = 28;                      // column nodes per staged row segment (27 + pad: 32-byte sectors stay aligned)
constexpr int VROW = 9 * BW + 3 * NP;       // staged velocity-node row: [r = 3c+d][28] then [c][8 pressure nodes]  (276)
constexpr int VPRS = 9 * BW;                // start of the pressure columns inside a velocity-node row
constexpr int PROW = 3 * BW;                // staged pressure-node row: [c][28]
constexpr int OFF_DG = NU * VROW + NP * PROW;   // fused preconditioner: m + nu k of every node pair, [27][28]
constexpr int OFF_PP = OFF_DG + NU * BW;        // pressure mass block [8][8]
constexpr int REC = OFF_PP + NP * NP;           // doubles per cell record (8 944)
constexpr int LROW = 384;                   // longest block(0,0) / block(1,0) row the gather accumulators hold
constexpr int ASTR = 387;                   // accumulator row stride (387 mod 16 == 3: the three components of one column land in different banks)
constexpr int L01 = 32;                     // longest block(0,1) row
constexpr int GACC = 3 * ASTR + 3 * L01 + 7;  // per-warp accumulators (+3 diagonal slots of constrained components), 1264
constexpr int GWARPS = 8;                  // warps per CTA of the preconditioner's gather
constexpr int GSW = 6;                     // warps per CTA of the system gather (each with a ring of staged rows)
constexpr int GD = 3;                      // depth of that ring: staged rows in flight per warp
constexpr int GSWARP = GACC + GD * VROW + 4;   // doubles per warp: accumulators, ring, mbarriers
static_assert(GSWARP % 2 == 0 && GACC % 2 == 0 && VROW % 2 == 0, "16-byte alignment of the ring slots");
static_assert(GACC % 2 == 0, "accumulator alignment");
static_assert((REC * 8) % 32 == 0 && (VROW * 8) % 32 == 0 && (PROW * 8) % 32 == 0 && (BW * 8) % 32 == 0, "sector alignment of the staged segments");

// ---- stage: contraction + constraint epilogue, coalesced write of the cell record ---------------------------------
// The DMMA accumulator layout does the transposition: lane (frow = lane / 4, fk = lane % 4) holds the 3x3 blocks of
// (row node 8 ta + frow, column nodes 8 tb + 2 fk + {0,1}).  Direct orientation: for every r = 3c+d the warp writes
// 8 rows x 64 contiguous bytes with one 16-byte store per lane; transposed orientation (row node b, column node a,
// entry [d][c]): for every (r, jj) 4 rows x 64 contiguous bytes with one 8-byte store per lane.  All segments start
// on 32-byte sector boundaries (BW = 28).
__constant__ unsigned char c_task_ta[10] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3};
__constant__ unsigned char c_task_tb[10] = {0, 1, 2, 3, 1, 2, 3, 2, 3, 3};

// `sched` names the cells of this CTA (cell(k, w, slot): its k-th plan cell and staging slot) and, in the persistent
// kernel, waits for the slot to be free (wait_slot) and publishes the finished record (signal).
template <class Sched>
__device__ __forceinline__ void stage_cells(const MmaArgs& a, const CsView& cs, const double* __restrict__ dphi_lane, double* __restrict__ stage,
                                            const Sched sched) {
  extern __shared__ __align__(16) double smem[];
  double* X = smem;                     // KQ * LDB
  double* wq = X + KQ * LDB;            // KQ (+4)
  double* sgeo2 = wq + 32;              // 2 x (GS + 1): mapping record of this cell and of the CTA's next one
  double* swt = sgeo2 + 2 * (GS + 1);   // 3*NU
  double* sF = swt + 3 * NU + 1;        // NQ*3
  double* sUc = sF + NQ * 3;            // 3*BW: old velocity, component-major, entry 27 of every row stays zero
  double* sT = sUc + 3 * BW;            // 32
  double* sTn = sT + 32;                // 28
  double* sGU = sTn + 28;               // NQ*12
  unsigned char* snm2 = (unsigned char*)(sGU + NQ * 12);       // 2 x MSTR
  int* sidx2 = (int*)(snm2 + 2 * MSTR);                        // 2 x IDS
  int* sidt2 = sidx2 + 2 * IDS;                                // 2 x 28
  int* sys_u = sidt2 + 2 * 28;                                 // 3*NU
  int* sys_p = sys_u + 3 * NU;                                 // NP
  unsigned char* skc = (unsigned char*)(sys_p + NP);           // 28
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < ND; i += nt) {
    const int f = a.local_field[i], bs = a.local_base[i];
    if (f < 3) sys_u[f * NU + bs] = i; else sys_p[bs] = i;
  }
  for (int i = tid; i < KQ * LDB; i += nt) X[i] = 0.0;
  for (int i = tid; i < 3 * BW; i += nt) sUc[i] = 0.0;
  if (tid < 4) wq[NQ + tid] = 0.0;
  __syncthreads();
  // cell-independent operand columns: reference values of the Q2 functions (alpha = 3) and of the Q1 functions
  for (int i = tid; i < NQ * NU; i += nt) X[(i / NU) * LDB + 96 + (i % NU)] = __ldg(a.phi_u + i);
  for (int i = tid; i < NQ * NP; i += nt) X[(i / NP) * LDB + PSI0 + (i % NP)] = __ldg(a.phi_p + i);
  const double nu = a.prm.dt * a.prm.inv_re;
  const bool do_rhs = a.rhs != nullptr;
  const double* sgeo = sgeo2;
  const unsigned char* snm = snm2;
  const int* sidx = sidx2;
  const int* sidt = sidt2;
  // mapping record, masks and dof indices of a cell, copied asynchronously into buffer `b`
  auto issue_raw = [&](long long w, int b) {
    const long long cell = a.cells[w];
    const double* g = a.geom + cell * GS;
    for (int i = tid; i < GS; i += nt) cp_async8(sgeo2 + b * (GS + 1) + i, g + i);
    if (tid < MSTR / 8) cp_async8(snm2 + b * MSTR + 8 * tid, a.nmask + w * MSTR + 8 * tid);
    for (int i = tid; i < ND; i += nt) cp_async4(sidx2 + b * IDS + i, a.l2g + cell * ND + i);
    if (do_rhs)
      for (int i = tid; i < a.ndt; i += nt) cp_async4(sidt2 + b * 28 + i, a.l2g_t + cell * a.ndt + i);
  };
  auto node_cs = [&](int n) {
    NodeCs c;
    c.mask = snm[n];
    c.k = skc[n];
    c.w0 = swt[3 * n];
    c.w1 = swt[3 * n + 1];
    c.w2 = swt[3 * n + 2];
    return c;
  };
  const int frow = lane >> 2, fk = lane & 3;

  int buf = 0, k = 0;
  long long w = 0, w_next = 0;
  int slot = 0, slot_next = 0;
  bool have = sched.cell(0, w, slot), have_next = false;
  if (have) issue_raw(w, 0);
  cp_async_commit();
  for (; have; ++k, buf ^= 1, w = w_next, slot = slot_next, have = have_next) {
    have_next = sched.cell(k + 1, w_next, slot_next);
    double* S = stage + (size_t)slot * REC;
    sgeo = sgeo2 + buf * (GS + 1);
    snm = snm2 + buf * MSTR;
    sidx = sidx2 + buf * IDS;
    sidt = sidt2 + buf * 28;
    cp_async_wait<0>();
    __syncthreads();   // this cell's raw inputs have landed; every warp is done with the previous cell
    if (Sched::FUSED && k > 0 && tid == 0) sched.signal(k - 1);   // ... and has fenced its stores of that cell
    if (do_rhs) {
      for (int i = tid; i < 3 * NU; i += nt) {
        const int c = i / NU, n = i - c * NU;
        cp_async8(sUc + c * BW + n, a.old_nse + sidx[sys_u[i]]);
      }
      for (int i = tid; i < a.ndt; i += nt) cp_async8(sTn + i, a.old_temp + sidt[i]);
    }
    cp_async_commit();
    if (have_next) issue_raw(w_next, buf ^ 1);   // the CTA's next cell, while this one is computed
    cp_async_commit();
    const int cflag = snm[35];
    if (tid >= 96 && tid - 96 < NU) {   // warp 3 (the table build below keeps warps 0..3 busy with tid < 108 only partly)
      const int n = tid - 96;
      int kc = 3;
      double w0 = 0.0, w1 = 0.0, w2 = 0.0;
      if (cflag && snm[n] != 7) {
        const int g0 = sidx[sys_u[n]];
        for (int c = 0; c < 3; ++c) {
          const int li = cs.line_of_dof[g0 + c];
          if (li >= 0 && cs.line_ptr[li + 1] > cs.line_ptr[li]) {
            kc = c;
            for (int k = cs.line_ptr[li]; k < cs.line_ptr[li + 1]; ++k) {
              const int mc = cs.entry_dof[k] - g0;
              const double wv = cs.entry_w[k];
              if (mc == 0) w0 = wv; else if (mc == 1) w1 = wv; else w2 = wv;
            }
          }
        }
      }
      skc[n] = (unsigned char)kc;
      swt[n * 3] = w0;
      swt[n * 3 + 1] = w1;
      swt[n * 3 + 2] = w2;
    }
    if (tid < NQ) wq[tid] = sgeo[tid];
    // operand table, physical gradients: thread = (quadrature point, group of 7 nodes); the 9 entries of d xi / d x stay
    // in registers, the stores of one warp instruction fall into 16 different bank pairs ((4 q + 7 group) mod 16)
    if (tid < 4 * NQ) {
      const int q = tid >> 2, bg = tid & 3;
      double kinv[3][3];
#pragma unroll
      for (int e = 0; e < 3; ++e)
#pragma unroll
        for (int d = 0; d < 3; ++d) kinv[e][d] = sgeo[NQ * (1 + 3 * e + d) + q];
      double* x = X + q * LDB + bg * 7;
      const int nb = bg == 3 ? 6 : 7;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        if (j >= nb) break;
        const double r0 = __ldg(dphi_lane + (j * 3) * (4 * NQ) + tid), r1 = __ldg(dphi_lane + (j * 3 + 1) * (4 * NQ) + tid),
                     r2 = __ldg(dphi_lane + (j * 3 + 2) * (4 * NQ) + tid);
#pragma unroll
        for (int d = 0; d < 3; ++d) x[32 * d + j] = kinv[0][d] * r0 + kinv[1][d] * r1 + kinv[2][d] * r2;
      }
    }
    if (Sched::FUSED && tid == 0) sched.wait_slot(k);   // the gather pass is done with the record this cell overwrites
    cp_async_wait<1>();   // the gathered old solution (the next cell's raw inputs may still be in flight)
    __syncthreads();

    // ---- tasks: 10 node x node blocks (ta <= tb), 4 node x psi blocks; weights 10 : 3, so the four warps take
    // {0,1,2}, {3,4,5}, {6,7,10,11}, {8,9,12,13}
    for (int s = 0; s < 4; ++s) {
      int t;
      if (warp < 2) {
        if (s == 3) {
          if (warp == 1) {   // pressure mass block of the preconditioner, psi x psi (:455-462)
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < KQ / 4; ++ks) {
              const int q = 4 * ks + fk;
              const double pv = X[q * LDB + PSI0 + frow];
              dmma(c0, c1, wq[q] * pv, pv);
            }
            __stcg(reinterpret_cast<double2*>(S + OFF_PP + frow * NP + 2 * fk), make_double2(c0, c1));
          }
          break;
        }
        t = 3 * warp + s;
      } else
        t = s < 2 ? (warp == 2 ? 6 : 8) + s : (warp == 2 ? 10 : 12) + (s - 2);
      if (t < 10) {
        const int ta = c_task_ta[t], tb_ = c_task_tb[t];
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[i][k][0] = acc[i][k][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KQ / 4; ++ks) {
          const int q = 4 * ks + fk;
          const double* xr = X + q * LDB;
          const double wv = wq[q];
          double af[4], bf[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            af[i] = wv * xr[32 * i + 8 * ta + frow];
            bf[i] = xr[32 * i + 8 * tb_ + frow];
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if ((i < 3 && k < 3) || (i == 3 && k == 3)) dmma(acc[i][k][0], acc[i][k][1], af[i], bf[k]);   // the value x gradient cross terms are not needed
        }
        const int na = 8 * ta + frow;
        double Fj[2][9], dgj[2];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int nb = 8 * tb_ + 2 * fk + jj;
          const double dg = acc[3][3][jj] + nu * (acc[0][0][jj] + acc[1][1][jj] + acc[2][2][jj]);
          dgj[jj] = dg;   // m_ab + nu k_ab: the velocity block of the preconditioner (:455-462)
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int d = 0; d < 3; ++d) Fj[jj][c * 3 + d] = nu * acc[d][c][jj] + (c == d ? dg : 0.0);
          if (cflag && na < NU && nb < NU) {   // constraint lines in this cell: C^T F C and the diagonals of constrained dofs
            const NodeCs ca = node_cs(na), cb = node_cs(nb);
            double F[3][3];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int d = 0; d < 3; ++d) F[c][d] = Fj[jj][c * 3 + d];
            const double d00 = fabs(F[0][0]), d11 = fabs(F[1][1]), d22 = fabs(F[2][2]);
            if (ca.k != 3 || cb.k != 3) {
              const double wa[3] = {ca.w0, ca.w1, ca.w2}, wb[3] = {cb.w0, cb.w1, cb.w2};
              double Fa[3], Fb[3], Fab;
#pragma unroll
              for (int d = 0; d < 3; ++d) Fa[d] = ca.k == 0 ? F[0][d] : (ca.k == 1 ? F[1][d] : (ca.k == 2 ? F[2][d] : 0.0));
#pragma unroll
              for (int c = 0; c < 3; ++c) Fb[c] = cb.k == 0 ? F[c][0] : (cb.k == 1 ? F[c][1] : (cb.k == 2 ? F[c][2] : 0.0));
              Fab = cb.k == 0 ? Fa[0] : (cb.k == 1 ? Fa[1] : (cb.k == 2 ? Fa[2] : 0.0));
#pragma unroll
              for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int d = 0; d < 3; ++d) F[c][d] += wa[c] * Fa[d] + wb[d] * Fb[c] + wa[c] * wb[d] * Fab;
            }
            if (na == nb) {  // constrained dofs keep |L_ii| on their own diagonal: staged in the (unused) slot [c][c]
              if (!(ca.mask & 1)) F[0][0] = d00;
              if (!(ca.mask & 2)) F[1][1] = d11;
              if (!(ca.mask & 4)) F[2][2] = d22;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int d = 0; d < 3; ++d) Fj[jj][c * 3 + d] = F[c][d];
          }
        }
        // direct orientation: row 8ta + frow, columns 8tb + 2fk + {0,1} (padding nodes contribute exact zeros)
        if (na < NU && 8 * tb_ + 2 * fk < BW) {
          double* row = S + na * VROW + 8 * tb_ + 2 * fk;
#pragma unroll
          for (int r9 = 0; r9 < 9; ++r9) __stcg(reinterpret_cast<double2*>(row + r9 * BW), make_double2(Fj[0][r9], Fj[1][r9]));
          __stcg(reinterpret_cast<double2*>(S + OFF_DG + na * BW + 8 * tb_ + 2 * fk), make_double2(dgj[0], dgj[1]));
        }
        if (ta != tb_) {
          // transposed orientation: row node b, column node a, entry [d][c] = F[c][d]
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int nb = 8 * tb_ + 2 * fk + jj;
            if (nb < NU) {   // column 8ta + frow <= 23 is always a real node here (ta < tb)
              double* row = S + nb * VROW + 8 * ta + frow;
#pragma unroll
              for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int d = 0; d < 3; ++d) __stcg(row + (d * 3 + c) * BW, Fj[jj][c * 3 + d]);
              __stcg(S + OFF_DG + nb * BW + 8 * ta + frow, dgj[jj]);
            }
          }
        }
      } else {
        // velocity-pressure coupling: rows (a, c) of row block ta against the 8 psi columns   (:633-635)
        const int ta = t - 10;
        double acc[3][2] = {{0, 0}, {0, 0}, {0, 0}};
#pragma unroll
        for (int ks = 0; ks < KQ / 4; ++ks) {
          const int q = 4 * ks + fk;
          const double* xr = X + q * LDB;
          const double wv = wq[q], bf = xr[PSI0 + frow];
#pragma unroll
          for (int i = 0; i < 3; ++i) dmma(acc[i][0], acc[i][1], wv * xr[32 * i + 8 * ta + frow], bf);
        }
        const int na = 8 * ta + frow;
        if (na < NU) {
          double sv[2][3];
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            sv[jj][0] = -acc[0][jj];
            sv[jj][1] = -acc[1][jj];
            sv[jj][2] = -acc[2][jj];
          }
          if (cflag) {
            const NodeCs ca = node_cs(na);
            if (ca.k != 3) {
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                const double sk = ca.k == 0 ? sv[jj][0] : (ca.k == 1 ? sv[jj][1] : sv[jj][2]);
                sv[jj][0] += ca.w0 * sk;
                sv[jj][1] += ca.w1 * sk;
                sv[jj][2] += ca.w2 * sk;
              }
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            // velocity-node row: [c][p]; pressure-node rows: [c][a]
            __stcg(reinterpret_cast<double2*>(S + na * VROW + VPRS + c * 8 + 2 * fk), make_double2(sv[0][c], sv[1][c]));
            __stcg(S + NU * VROW + (2 * fk) * PROW + c * BW + na, sv[0][c]);
            __stcg(S + NU * VROW + (2 * fk + 1) * PROW + c * BW + na, sv[1][c]);
          }
        }
      }
    }
    if (do_rhs) {
      // old velocity and its gradient at the quadrature points on the tensor cores, too: for alpha = warp,
      // G[q][c] = sum_n X[q][32 alpha + n] u_c[n]  (4 row tiles of quadrature points x 7 k-steps over the nodes)
      {
        const int e = warp;
        double g4[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
#pragma unroll
        for (int ks = 0; ks < 7; ++ks) {
          const double bfr = frow < 3 ? sUc[frow * BW + 4 * ks + fk] : 0.0;
#pragma unroll
          for (int t4 = 0; t4 < 4; ++t4) {
            const int q = 8 * t4 + frow;
            const double afr = q < KQ ? X[q * LDB + 32 * e + 4 * ks + fk] : 0.0;
            dmma(g4[t4][0], g4[t4][1], afr, bfr);
          }
        }
#pragma unroll
        for (int t4 = 0; t4 < 4; ++t4) {
          const int q = 8 * t4 + frow;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int c = 2 * fk + jj;
            if (q < NQ && c < 3) sGU[q * 12 + c * 4 + e] = g4[t4][jj];
          }
        }
      }
      if (warp == 2 && lane < NQ) {
        double tq = 0.0;
        for (int k = 0; k < a.ndt; ++k) tq += sTn[k] * __ldg(a.phi_t + lane * a.ndt + k);
        sT[lane] = tq;
      }
      __syncthreads();
      for (int q = tid; q < NQ; q += nt) {
        double u[3], gu[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          u[c] = sGU[q * 12 + c * 4 + 3];
#pragma unroll
          for (int d = 0; d < 3; ++d) gu[c][d] = sGU[q * 12 + c * 4 + d];
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
        for (int q = 0; q < NQ; ++q) s += X[q * LDB + 96 + n] * sF[q * 3 + c];
        const int gi = sidx[sys_u[i]];
        if (snm[n] & (1 << c))
          red_add_f64(a.rhs + gi, s);
        else {
          const int li = cs.line_of_dof[gi];
          for (int k = cs.line_ptr[li]; k < cs.line_ptr[li + 1]; ++k) red_add_f64(a.rhs + cs.entry_dof[k], cs.entry_w[k] * s);
        }
      }
    }
    if (Sched::FUSED) __threadfence();   // staged record visible device-wide before the barrier that precedes signal()
  }
  if (Sched::FUSED) {
    __syncthreads();
    if (tid == 0 && k > 0) sched.signal(k - 1);
  }
}

// one launch per chunk: the CTAs stride over the chunk's cells
struct LaunchSched {
  static constexpr bool FUSED = false;
  long long w_begin, w_end;
  int slot_begin, ring;
  __device__ __forceinline__ bool cell(int k, long long& w, int& slot) const {
    w = w_begin + blockIdx.x + (long long)k * gridDim.x;
    if (w >= w_end) return false;
    slot = slot_begin + (int)(w - w_begin);
    if (slot >= ring) slot -= ring;
    return true;
  }
  __device__ __forceinline__ void wait_slot(int) const {}
  __device__ __forceinline__ void signal(int) const {}
};

__global__ void __launch_bounds__(MTHREADS, 4)
th_stage_kernel(MmaArgs a, CsView cs, const double* __restrict__ dphi_lane, double* __restrict__ stage, long long w_begin, long long w_end,
                int slot_begin, int ring) {
  stage_cells(a, cs, dphi_lane, stage, LaunchSched{w_begin, w_end, slot_begin, ring});
}

// reference gradients in the order the table build reads them: [node of the group j][e][thread = 4 q + group]
__global__ void dphi_lane_kernel(const double* __restrict__ dphi_u, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 7 * 3 * 4 * NQ) return;
  const int t = i % (4 * NQ), je = i / (4 * NQ), e = je % 3, j = je / 3;
  const int q = t >> 2, b = (t & 3) * 7 + j;
  out[i] = b < NU ? dphi_u[(size_t)(q * NU + b) * 3 + e] : 0.0;
}

constexpr size_t stage_smem_bytes() {
  return sizeof(double) * (KQ * LDB + 32 + 2 * (GS + 1) + 3 * NU + 1 + NQ * 3 + 3 * BW + 32 + 28 + NQ * 12) + 2 * MSTR +
         sizeof(int) * (2 * IDS + 2 * 28 + 3 * NU + NP) + 28 + 36;
}

// ---- gather: one warp per (chunk, node) -------------------------------------------------------------------------
struct GatherArgs {
  const int* v_g0;
  const unsigned* v_incptr;
  const unsigned char* v_flag;
  const unsigned* v_inc;
  const int* p_g0;
  const unsigned* p_incptr;
  const unsigned char* p_flag;
  const unsigned* p_inc;
  long long v_begin, v_end, p_begin, p_end;  // item ranges of this chunk
  long long w_base;                          // first plan cell of the chunk
  int slot_base, ring;                       // its staging slot; slots wrap at `ring`
  const unsigned short* pos;
  const unsigned char* nmask;
  const double* stage;
};

constexpr unsigned INC_CS = 0x80000000u;     // incidence word: the cell holds constrained velocity dofs

// what one incidence (cell of the chunk, local node) contributes to a velocity node: loaded one incidence ahead
// (positions and masks keep their 16- / 8-bit types until they are used: a widening right after the load would make the
// prefetch wait for it)
struct VInc {
  double v[9];   // lane = column node b: staged [r][b]
  double vp;     // lanes < 24: staged [c][p]
  unsigned short ob, op;
  unsigned char mb, maskA;
  int a;
};

__device__ __forceinline__ void load_vinc(VInc& I, const GatherArgs& g, unsigned e, int lane) {
  const int a = e & 31;
  const unsigned wl = (e & ~INC_CS) >> 5;
  const size_t w = (size_t)g.w_base + wl;
  int slot = g.slot_base + (int)wl;
  if (slot >= g.ring) slot -= g.ring;
  const unsigned short* prow = g.pos + w * PSTR + a * NE;
  const double* S = g.stage + (size_t)slot * REC + a * VROW;
  I.a = a;
  I.ob = 0;
  I.op = 0;
  if (lane < NU) I.ob = prow[lane];
  if (lane < NP) I.op = prow[NU + lane];
  I.mb = 7;
  I.maskA = 7;
  if (e & INC_CS) {
    const unsigned char* mrow = g.nmask + w * MSTR;
    I.mb = 0;
    if (lane < NU) I.mb = mrow[lane];
    I.maskA = mrow[a];
  }
#pragma unroll
  for (int r = 0; r < 9; ++r) I.v[r] = lane < NU ? __ldcg(S + r * BW + lane) : 0.0;
  I.vp = lane < 24 ? __ldcg(S + VPRS + lane) : 0.0;
}

// add one incidence into the warp's accumulators (the nine targets of a lane are distinct: load all, then store all)
__device__ __forceinline__ void add_vinc(const VInc& I, double* acc, double* acc01, double* accd, int lane) {
  const int ob = I.ob, mb = I.mb;
  if (I.maskA == 7 && __all_sync(0xffffffffu, lane >= NU || mb == 7)) {
    if (lane < NU) {
      double tv[9];
#pragma unroll
      for (int r = 0; r < 9; ++r) tv[r] = acc[(r / 3) * ASTR + ob + (r % 3)];
#pragma unroll
      for (int r = 0; r < 9; ++r) acc[(r / 3) * ASTR + ob + (r % 3)] = tv[r] + I.v[r];
    }
    const int o = __shfl_sync(0xffffffffu, (int)I.op, lane & 7);
    if (lane < 24) acc01[(lane >> 3) * L01 + o] += I.vp;
  } else {
    const int maskA = I.maskA;
    if (lane < NU) {
      int idx[9];
      double tv[9];
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const int c = r / 3, d = r - 3 * c;
        const bool on = ((maskA >> c) & 1) && ((mb >> d) & 1);
        idx[r] = on ? c * ASTR + ob + __popc(mb & ((1 << d) - 1)) : -1;
      }
#pragma unroll
      for (int r = 0; r < 9; ++r) tv[r] = idx[r] >= 0 ? acc[idx[r]] : 0.0;
#pragma unroll
      for (int r = 0; r < 9; ++r)
        if (idx[r] >= 0) acc[idx[r]] = tv[r] + I.v[r];
      if (lane == I.a && maskA != 7) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (!((maskA >> c) & 1)) accd[c] += I.v[4 * c];
      }
    }
    const int c = lane >> 3;
    const int o = __shfl_sync(0xffffffffu, (int)I.op, lane & 7);
    if (lane < 24 && ((maskA >> c) & 1)) acc01[c * L01 + o] += I.vp;
  }
  __syncwarp();
}

// write `len` accumulated entries to the row at `out` (store, or add to what earlier chunks left there) and leave the
// accumulators zero for the next item
__device__ __forceinline__ void flush_row(double* __restrict__ out, double* acc, int len, bool first, int lane) {
  if (first) {
#pragma unroll 4
    for (int k = lane; k < len; k += 32) {
      out[k] = acc[k];
      acc[k] = 0.0;
    }
  } else {
    for (int k0 = lane; k0 < len; k0 += 128) {
      double old[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) old[u] = k0 + 32 * u < len ? __ldcg(out + k0 + 32 * u) : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + 32 * u;
        if (k < len) {
          out[k] = old[u] + acc[k];
          acc[k] = 0.0;
        }
      }
    }
  }
}

constexpr unsigned FULLM = 0xffffffffu;
constexpr int WB = 32;   // items per block: a warp takes blocks round-robin, so expensive items (rows that earlier chunks
                         // already touched are read-modify-written) spread over all warps

// Walk the items [begin, end) block-cyclically.  Per block: the 32 item headers are read at once (lane t <-> item t);
// the incidence words of the block are read 32 at a time, one batch ahead; the loads of incidence i + 1 (`load`) and
// the row starts of item t + 1 (`rows`) are in flight while incidence i / item t are processed (`add`, `flush`).
// `aux` maps an incidence word to a second index that is fetched together with the words.
template <class Inc, int WBT, class AuxF, class RowF, class LoadF, class AddF, class FlushF>
__device__ __forceinline__ void walk_items(const int* __restrict__ g0s, const unsigned char* __restrict__ flags, const unsigned* __restrict__ incptr,
                                           const unsigned* __restrict__ incs, long long begin, long long end, long long gw, long long nw, int lane,
                                           AuxF aux, RowF rows, LoadF load, AddF add, FlushF flush) {
  for (long long j_lo = begin + gw * WBT; j_lo < end; j_lo += nw * WBT) {
    const int n_it = (int)((j_lo + WBT < end ? j_lo + WBT : end) - j_lo);
    int h_g0 = 0, h_first = 0;
    unsigned h_iend = 0;
    if (lane < n_it) {
      h_g0 = g0s[j_lo + lane];
      h_first = flags[j_lo + lane] & 1;
      h_iend = incptr[j_lo + lane + 1];
    }
    const unsigned i_lo = incptr[j_lo], i_hi = __shfl_sync(FULLM, h_iend, n_it - 1);
    unsigned ec = 0, en = 0, ib = i_lo;
    long long ac = -1, an = -1;
    if (ib + lane < i_hi) { ec = incs[ib + lane]; ac = aux(ec); }
    if (ib + 32 + lane < i_hi) { en = incs[ib + 32 + lane]; an = aux(en); }
    auto rotate = [&]() {
      ib += 32;
      ec = en; ac = an;
      en = 0; an = -1;
      if (ib + 32 + lane < i_hi) { en = incs[ib + 32 + lane]; an = aux(en); }
    };
    auto word = [&](unsigned i) {
      const unsigned k = i - ib;
      return k < 32 ? __shfl_sync(FULLM, ec, (int)k) : __shfl_sync(FULLM, en, (int)(k - 32));
    };
    auto auxw = [&](unsigned i) {
      const unsigned k = i - ib;
      return k < 32 ? __shfl_sync(FULLM, ac, (int)k) : __shfl_sync(FULLM, an, (int)(k - 32));
    };
    long long rnext = rows(__shfl_sync(FULLM, h_g0, 0));
    Inc bufA, bufB;
    unsigned i = i_lo;
    load(bufA, word(i), auxw(i));
    for (int t = 0; t < n_it; ++t) {
      const unsigned it_end = __shfl_sync(FULLM, h_iend, t);
      const bool first = __shfl_sync(FULLM, h_first, t) != 0;
      const long long rcur = rnext;
      if (t + 1 < n_it) rnext = rows(__shfl_sync(FULLM, h_g0, t + 1));
      // incidences of the item, two per trip: one buffer is consumed while the other one's loads are in flight
      while (i < it_end) {
        if (i - ib >= 32) rotate();
        if (i + 1 < i_hi) load(bufB, word(i + 1), auxw(i + 1));
        add(bufA);
        ++i;
        if (i < it_end) {
          if (i - ib >= 32) rotate();
          if (i + 1 < i_hi) load(bufA, word(i + 1), auxw(i + 1));
          add(bufB);
          ++i;
        } else {
          bufA = bufB;   // the prefetched incidence belongs to the next item
          break;
        }
      }
      flush(rcur, first);
    }
  }
}

// (A register ring with 3-4 incidences in flight per warp was measured slower than the two-buffer walk above: 1.5 ms
// against 1.16 ms per 65 536 cells for the preconditioner's gather -- these passes are bound by the DRAM efficiency of
// their 2 kB-granular accesses, not by the depth of the prefetch.)

// ---- TMA bulk copies into a per-warp ring ---------------------------------------------------------------------------
// A staged velocity-node row is one regular 2 208-byte tile: one elected lane fetches it with cp.async.bulk (SASS UBLKCP),
// completion is counted on an mbarrier of the slot.  GD rows are in flight per warp while one is added -- the gather is
// bound by the latency of these loads, not by their bytes.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load_row(double* dst, const double* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slot's previous readers (generic proxy) come first
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// velocity items of one chunk through the ring (same item / incidence walk as walk_items)
template <int WBT>
__device__ __forceinline__ void gather_system_velocity_bulk(const GatherArgs& g, const BlockView& A, double* acc, double* ring,
                                                            unsigned long long* bars, long long gw, long long nw, int lane) {
  double* acc01 = acc + 3 * ASTR;
  double* accd = acc01 + 3 * L01;
  const long long* rp00 = A.rowptr[0][0];
  const long long* rp01 = A.rowptr[0][1];
  double* v00 = A.val[0][0];
  double* v01 = A.val[0][1];
  unsigned phases = 0;   // parity of every slot's next completion
  for (long long j_lo = g.v_begin + gw * WBT; j_lo < g.v_end; j_lo += nw * WBT) {
    const int n_it = (int)((j_lo + WBT < g.v_end ? j_lo + WBT : g.v_end) - j_lo);
    int h_g0 = 0, h_first = 0;
    unsigned h_iend = 0;
    if (lane < n_it) {
      h_g0 = g.v_g0[j_lo + lane];
      h_first = g.v_flag[j_lo + lane] & 1;
      h_iend = g.v_incptr[j_lo + lane + 1];
    }
    const unsigned i_lo = g.v_incptr[j_lo], i_hi = __shfl_sync(FULLM, h_iend, n_it - 1);
    unsigned ec = 0, en = 0, ib = i_lo;
    if (ib + lane < i_hi) ec = g.v_inc[ib + lane];
    if (ib + 32 + lane < i_hi) en = g.v_inc[ib + 32 + lane];
    auto word = [&](unsigned i) {
      const unsigned k = i - ib;
      return k < 32 ? __shfl_sync(FULLM, ec, (int)k) : __shfl_sync(FULLM, en, (int)(k - 32));
    };
    auto rows = [&](int g0) { return lane < 4 ? rp00[g0 + lane] : (lane < 8 ? rp01[g0 + lane - 4] : 0ll); };
    // per-slot metadata (registers; the slot index is a compile-time constant everywhere below)
    unsigned short m_ob[GD], m_op[GD];
    unsigned char m_mb[GD], m_maskA[GD];
    int m_a[GD];
    auto issue = [&](unsigned i, int u, unsigned short& ob, unsigned char& mb, unsigned short& op, int& aa, unsigned char& maskA) {
      const unsigned e = word(i);
      const int a = e & 31;
      const unsigned wl = (e & ~INC_CS) >> 5;
      const size_t w = (size_t)g.w_base + wl;
      int slot = g.slot_base + (int)wl;
      if (slot >= g.ring) slot -= g.ring;
      if (lane == 0) bulk_load_row(ring + u * VROW, g.stage + (size_t)slot * REC + a * VROW, VROW * 8, bars + u);
      const unsigned short* prow = g.pos + w * PSTR + a * NE;
      aa = a;
      ob = 0;
      op = 0;
      if (lane < NU) ob = prow[lane];
      if (lane < NP) op = prow[NU + lane];
      mb = 7;
      maskA = 7;
      if (e & INC_CS) {
        const unsigned char* mrow = g.nmask + w * MSTR;
        mb = 0;
        if (lane < NU) mb = mrow[lane];
        maskA = mrow[a];
      }
    };
    long long rnext = rows(__shfl_sync(FULLM, h_g0, 0));
#pragma unroll
    for (int u = 0; u < GD; ++u)
      if (i_lo + u < i_hi) issue(i_lo + u, u, m_ob[u], m_mb[u], m_op[u], m_a[u], m_maskA[u]);
    unsigned i = i_lo;
    int t = 0, maskA_item = 7;
    unsigned it_end = __shfl_sync(FULLM, h_iend, 0);
    bool first = __shfl_sync(FULLM, h_first, 0) != 0;
    long long rcur = rnext;
    if (n_it > 1) rnext = rows(__shfl_sync(FULLM, h_g0, 1));
    while (i < i_hi) {
#pragma unroll
      for (int u = 0; u < GD; ++u) {
        if (i < i_hi) {
          mbar_wait(bars + u, (phases >> u) & 1u);
          phases ^= 1u << u;
          VInc I;
          const double* S = ring + u * VROW;
#pragma unroll
          for (int r = 0; r < 9; ++r) I.v[r] = lane < NU ? S[r * BW + lane] : 0.0;
          I.vp = lane < 24 ? S[VPRS + lane] : 0.0;
          I.ob = m_ob[u];
          I.mb = m_mb[u];
          I.op = m_op[u];
          I.a = m_a[u];
          I.maskA = m_maskA[u];
          maskA_item = I.maskA;
          add_vinc(I, acc, acc01, accd, lane);   // ends with __syncwarp: every lane has read the slot
          if (i - ib >= 32) {   // keep the word window ahead of the prefetch distance
            ib += 32;
            ec = en;
            en = 0;
            if (ib + 32 + lane < i_hi) en = g.v_inc[ib + 32 + lane];
          }
          if (i + GD < i_hi) issue(i + GD, u, m_ob[u], m_mb[u], m_op[u], m_a[u], m_maskA[u]);
          ++i;
          if (i == it_end) {   // the item is complete: write its rows, move to the next item
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const long long rs = __shfl_sync(FULLM, rcur, c), rs01 = __shfl_sync(FULLM, rcur, 4 + c);
              const int len = (int)(__shfl_sync(FULLM, rcur, c + 1) - rs), len01 = (int)(__shfl_sync(FULLM, rcur, 5 + c) - rs01);
              if ((maskA_item >> c) & 1) {
                flush_row(v00 + rs, acc + c * ASTR, len, first, lane);
                flush_row(v01 + rs01, acc01 + c * L01, len01, first, lane);
              } else if (lane == 0) {
                double* out = v00 + rs;   // constrained dof: the row holds its diagonal only
                if (first) {
                  out[0] = accd[c];
                  for (int k = 1; k < len; ++k) out[k] = 0.0;
                  for (int k = 0; k < len01; ++k) v01[rs01 + k] = 0.0;
                } else
                  out[0] = __ldcg(out) + accd[c];
                accd[c] = 0.0;
              }
            }
            __syncwarp();
            maskA_item = 7;
            ++t;
            if (t < n_it) {
              it_end = __shfl_sync(FULLM, h_iend, t);
              first = __shfl_sync(FULLM, h_first, t) != 0;
              rcur = rnext;
              if (t + 1 < n_it) rnext = rows(__shfl_sync(FULLM, h_g0, t + 1));
            }
          }
        }
      }
    }
  }
}

struct PInc {
  double v[3];
  unsigned short ob;
  unsigned char mb;
};

// the items of one chunk, system matrix: `acc` = this warp's zeroed accumulators [3][ASTR] + [3][L01] + diagonals
template <int WBT, bool VELOCITY = true>
__device__ __forceinline__ void gather_system(const GatherArgs& g, const BlockView& A, double* acc, long long gw, long long nw, int lane) {
  double* acc01 = acc + 3 * ASTR;           // [3][L01]
  double* accd = acc01 + 3 * L01;           // [3] constrained diagonals
  const long long* rp00 = A.rowptr[0][0];
  const long long* rp01 = A.rowptr[0][1];
  const long long* rp10 = A.rowptr[1][0];
  double* v00 = A.val[0][0];
  double* v01 = A.val[0][1];
  double* v10 = A.val[1][0];
  auto no_aux = [](unsigned) { return 0ll; };

  // velocity nodes: rows (g0 + c) of block(0,0) and block(0,1)
  if (VELOCITY) {
    int maskA = 7;
    auto rows = [&](int g0) { return lane < 4 ? rp00[g0 + lane] : (lane < 8 ? rp01[g0 + lane - 4] : 0ll); };
    auto load = [&](VInc& I, unsigned e, long long) { load_vinc(I, g, e, lane); };
    auto add = [&](const VInc& I) {
      maskA = I.maskA;
      add_vinc(I, acc, acc01, accd, lane);
    };
    auto flush = [&](long long rcur, bool first) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const long long rs = __shfl_sync(FULLM, rcur, c), rs01 = __shfl_sync(FULLM, rcur, 4 + c);
        const int len = (int)(__shfl_sync(FULLM, rcur, c + 1) - rs), len01 = (int)(__shfl_sync(FULLM, rcur, 5 + c) - rs01);
        if ((maskA >> c) & 1) {
          flush_row(v00 + rs, acc + c * ASTR, len, first, lane);
          flush_row(v01 + rs01, acc01 + c * L01, len01, first, lane);
        } else if (lane == 0) {
          // constrained dof: the row holds its diagonal only
          double* out = v00 + rs;
          if (first) {
            out[0] = accd[c];
            for (int k = 1; k < len; ++k) out[k] = 0.0;
            for (int k = 0; k < len01; ++k) v01[rs01 + k] = 0.0;
          } else
            out[0] = __ldcg(out) + accd[c];
          accd[c] = 0.0;
        }
      }
      maskA = 7;
      __syncwarp();
    };
    walk_items<VInc, WBT>(g.v_g0, g.v_flag, g.v_incptr, g.v_inc, g.v_begin, g.v_end, gw, nw, lane, no_aux, rows, load, add, flush);
  }

  // pressure nodes: row of block(1,0)
  {
    auto rows = [&](int pr) { return lane < 2 ? rp10[pr + lane] : 0ll; };
    auto load = [&](PInc& I, unsigned e, long long) {
      const int pn = e & 31;
      const unsigned wl = (e & ~INC_CS) >> 5;
      const size_t w = (size_t)g.w_base + wl;
      int slot = g.slot_base + (int)wl;
      if (slot >= g.ring) slot -= g.ring;
      I.ob = 0;
      I.mb = 0;
      I.v[0] = I.v[1] = I.v[2] = 0.0;
      if (lane < NU) {
        I.ob = g.pos[w * PSTR + (NU + pn) * NE + lane];
        I.mb = 7;
        if (e & INC_CS) I.mb = g.nmask[w * MSTR + lane];
        const double* S = g.stage + (size_t)slot * REC + NU * VROW + pn * PROW;
#pragma unroll
        for (int c = 0; c < 3; ++c) I.v[c] = __ldcg(S + c * BW + lane);
      }
    };
    auto add = [&](const PInc& I) {
      if (lane < NU) {
        int idx[3];
        double tv[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) idx[c] = ((I.mb >> c) & 1) ? (int)I.ob + __popc((int)I.mb & ((1 << c) - 1)) : -1;
#pragma unroll
        for (int c = 0; c < 3; ++c) tv[c] = idx[c] >= 0 ? acc[idx[c]] : 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (idx[c] >= 0) acc[idx[c]] = tv[c] + I.v[c];
      }
      __syncwarp();
    };
    auto flush = [&](long long rcur, bool first) {
      const long long rs = __shfl_sync(FULLM, rcur, 0);
      const int len = (int)(__shfl_sync(FULLM, rcur, 1) - rs);
      flush_row(v10 + rs, acc, len, first, lane);
      __syncwarp();
    };
    walk_items<PInc, WBT>(g.p_g0, g.p_flag, g.p_incptr, g.p_inc, g.p_begin, g.p_end, gw, nw, lane, no_aux, rows, load, add, flush);
  }
}

__global__ void __launch_bounds__(GWARPS * 32, 2) th_gather_kernel(GatherArgs g, BlockView A) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* acc = smem + warp * GACC;
  for (int k = lane; k < GACC; k += 32) acc[k] = 0.0;   // invariant: all accumulators are zero between items
  __syncwarp();
  gather_system<WB>(g, A, acc, (long long)blockIdx.x * GWARPS + warp, (long long)gridDim.x * GWARPS, lane);
}

// The same pass with the staged rows fetched by TMA bulk copies into a ring of GD rows per warp (DCP_GATHER_BULK=1).
// Measured slower than the register pipeline above (2.7 ms against 1.9 ms per 65 536 cells: the ring costs a quarter of
// the resident warps and the pass is bound by DRAM efficiency, not by the depth of the prefetch); kept selectable.
__global__ void __launch_bounds__(GSW * 32, 2) th_gather_bulk_kernel(GatherArgs g, BlockView A) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* acc = smem + warp * GSWARP;
  double* ring = acc + GACC;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ring + GD * VROW);
  for (int k = lane; k < GACC; k += 32) acc[k] = 0.0;   // invariant: all accumulators are zero between items
  if (lane == 0) {
    for (int u = 0; u < GD; ++u) mbar_init(bars + u, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const long long gw = (long long)blockIdx.x * GSW + warp, nw = (long long)gridDim.x * GSW;
  gather_system_velocity_bulk<WB>(g, A, acc, ring, bars, gw, nw, lane);
  gather_system<WB, false>(g, A, acc, gw, nw, lane);
}

// ---- fused preconditioner: second gather over the same (chunk, node) items ---------------------------------------
// nse_preconditioner_matrix (include/core/boussinesq_model.tpp:421-476): velocity rows hold m + nu k on the same
// component only, pressure rows the pressure mass matrix.  Positions and masks come from the preconditioner's own plan
// (pre_w: its index of a system-plan cell).  Cells with no-normal-flux lines -- their C^T (dg I) C spreads over other
// component pairs -- and cells outside that plan are skipped here and added afterwards by the reduction kernels; the
// first chunk that touches a node stores its whole rows, so those later additions start from a defined state.
struct PreArgs {
  const long long* pre_w;           // per system-plan cell: low word = index in the preconditioner plan (-1: skipped here),
                                    // high word = (index of its wide table + 1) << 1 | cell has constrained dofs
  const unsigned short* pos;
  const unsigned char* nmask;
  const unsigned short* pos_wide;
};

struct PVInc {
  double v;
  unsigned short o0, o1, o2;
  unsigned char mb, maskA;
  int a, skip;
};
struct PPInc {
  double v;
  unsigned short o;
  int skip;
};

// Accumulators of the preconditioner's gather: [3][PSTRIDE] + 3 diagonals.  The gathered contributions land on the
// same component only (at most 125 entries per row; checked when the plan is attached); rows of nodes with
// no-normal-flux lines are longer, but all their cells are skipped here, so the tail beyond PSTRIDE is stored as zeros.
constexpr int PSTRIDE = 131;                   // 131 mod 16 == 3
constexpr int PACC = 3 * PSTRIDE + 7;          // 400 doubles per warp

template <int WBT>
__device__ __forceinline__ void flush_row_capped(double* __restrict__ out, double* acc, int len, int cap, bool first, int lane) {
  const int n = len < cap ? len : cap;
  flush_row(out, acc, n, first, lane);
  if (first)
    for (int k = cap + lane; k < len; k += 32) out[k] = 0.0;
}

template <int WBT, int PS = PSTRIDE>
__device__ __forceinline__ void gather_pre(const GatherArgs& g, const PreArgs& pa, const BlockView& A, double* acc, long long gw, long long nw, int lane) {
  double* accd = acc + 3 * PS;
  const long long* rp00 = A.rowptr[0][0];
  const long long* rp11 = A.rowptr[1][1];
  double* v00 = A.val[0][0];
  double* v11 = A.val[1][1];
  auto aux = [&](unsigned e) { return pa.pre_w[(size_t)g.w_base + ((e & ~INC_CS) >> 5)]; };
  auto slot_of = [&](unsigned e) {
    int slot = g.slot_base + (int)((e & ~INC_CS) >> 5);
    return slot >= g.ring ? slot - g.ring : slot;
  };
  // velocity nodes
  {
    int maskA_item = 7;
    auto rows = [&](int g0) { return lane < 4 ? rp00[g0 + lane] : 0ll; };
    auto load = [&](PVInc& I, unsigned e, long long ax) {
      const int wp = (int)(ax & 0xffffffffll), hi = (int)(ax >> 32);
      I.skip = wp < 0;
      I.a = e & 31;
      I.v = 0.0;
      I.o0 = I.o1 = I.o2 = 0xffff;
      I.mb = I.maskA = 7;
      if (wp < 0) return;
      const unsigned char* nm = pa.nmask + (size_t)wp * MSTR;
      const int cflag = hi & 1, wide = (hi >> 1) - 1;
      if (lane < NU) {
        I.v = __ldcg(g.stage + (size_t)slot_of(e) * REC + OFF_DG + I.a * BW + lane);
        if (wide >= 0) {
          const unsigned short* b = pa.pos_wide + (size_t)wide * (3 * NU * NU) + I.a * NU + lane;
          I.o0 = b[0];
          I.o1 = b[NU * NU];
          I.o2 = b[2 * NU * NU];
        } else {
          const unsigned short o = pa.pos[(size_t)wp * PSTR + I.a * NE + lane];
          I.o0 = o;
          I.o1 = o;
          I.o2 = o;
        }
        if (cflag) I.mb = nm[lane];
      }
      if (cflag) I.maskA = nm[I.a];
    };
    auto add = [&](const PVInc& I) {
      if (!I.skip) {
        maskA_item = I.maskA;
        if (lane < NU) {
          const int mm = I.maskA & I.mb;
          if ((mm & 1) && I.o0 != 0xffff) acc[I.o0] += I.v;
          if ((mm & 2) && I.o1 != 0xffff) acc[PS + I.o1] += I.v;
          if ((mm & 4) && I.o2 != 0xffff) acc[2 * PS + I.o2] += I.v;
          if (lane == I.a && I.maskA != 7) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
              if (!((I.maskA >> c) & 1)) accd[c] += fabs(I.v);
          }
        }
      }
      __syncwarp();
    };
    auto flush = [&](long long rcur, bool first) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const long long rs = __shfl_sync(FULLM, rcur, c);
        const int len = (int)(__shfl_sync(FULLM, rcur, c + 1) - rs);
        if ((maskA_item >> c) & 1)
          flush_row_capped<WBT>(v00 + rs, acc + c * PS, len, PS, first, lane);
        else if (lane == 0) {
          double* out = v00 + rs;
          if (first) {
            out[0] = accd[c];
            for (int k = 1; k < len; ++k) out[k] = 0.0;
          } else
            out[0] = __ldcg(out) + accd[c];
          accd[c] = 0.0;
        }
      }
      maskA_item = 7;
      __syncwarp();
    };
    walk_items<PVInc, WBT>(g.v_g0, g.v_flag, g.v_incptr, g.v_inc, g.v_begin, g.v_end, gw, nw, lane, aux, rows, load, add, flush);
  }
  // pressure nodes: rows of block(1,1)
  {
    auto rows = [&](int pr) { return lane < 2 ? rp11[pr + lane] : 0ll; };
    auto load = [&](PPInc& I, unsigned e, long long ax) {
      const int wp = (int)(ax & 0xffffffffll);
      I.skip = wp < 0;
      I.v = 0.0;
      I.o = 0xffff;
      if (wp < 0 || lane >= NP) return;
      const int pn = e & 31;
      I.v = __ldcg(g.stage + (size_t)slot_of(e) * REC + OFF_PP + pn * NP + lane);
      I.o = pa.pos[(size_t)wp * PSTR + (NU + pn) * NE + NU + lane];
    };
    auto add = [&](const PPInc& I) {
      if (!I.skip && lane < NP && I.o != 0xffff) acc[I.o] += I.v;
      __syncwarp();
    };
    auto flush = [&](long long rcur, bool first) {
      const long long rs = __shfl_sync(FULLM, rcur, 0);
      const int len = (int)(__shfl_sync(FULLM, rcur, 1) - rs);
      flush_row_capped<WBT>(v11 + rs, acc, len, PS, first, lane);
      __syncwarp();
    };
    walk_items<PPInc, WBT>(g.p_g0, g.p_flag, g.p_incptr, g.p_inc, g.p_begin, g.p_end, gw, nw, lane, aux, rows, load, add, flush);
  }
}

__global__ void __launch_bounds__(GWARPS * 32, 3) th_pre_gather_kernel(GatherArgs g, PreArgs pa, BlockView A) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* acc = smem + warp * PACC;
  for (int k = lane; k < PACC; k += 32) acc[k] = 0.0;
  __syncwarp();
  gather_pre<WB>(g, pa, A, acc, (long long)blockIdx.x * GWARPS + warp, (long long)gridDim.x * GWARPS, lane);
}

// ---- persistent kernel: staging ring in L2 ------------------------------------------------------------------------
// One cooperative launch for the whole pass.  The first n_stage CTAs stage cells (chunk k = the k-th cell of every
// stage CTA), the other CTAs gather: chunk by chunk, as soon as all its records are staged.  The ring holds `ring_chunks`
// chunks (a few tens of MB: it lives in L2, the records never travel to HBM and back); a stage CTA overwrites a slot
// once every gather warp has finished the chunk that used it.  Gather warps finish chunk c before any of them starts
// c + 1, so the rows a node shares between chunks are read-modify-written in chunk order without atomics.
// Dependencies: stage(k) <- gathered(k - ring_chunks) <- staged(k - ring_chunks): no cycle; the launch is cooperative, so
// all CTAs are resident.  Every wait gives up after about a second and raises the abort flag (d_err[1]) instead of
// hanging the device.
struct FusedSync {
  unsigned* staged;    // [n_chunks] records published
  unsigned* gdone;     // [n_chunks] gather warps finished
  int* err;            // d_err: [1] = abort
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void spin_until(const unsigned* p, unsigned target, int* err) {
  for (unsigned it = 0; it < (1u << 23); ++it) {
    if (ld_acquire_u32(p) >= target) return;
    if ((it & 1023) == 1023 && *reinterpret_cast<volatile int*>(err + 1) != 0) return;
    __nanosleep(100);
  }
  atomicExch(err + 1, 1);   // give up: the results are invalid, the host reports it (dcp_check_device_errors)
}

struct FusedSched {
  static constexpr bool FUSED = true;
  long long n_cells;
  int n_stage, ring_chunks, n_gather_warps, rank;
  FusedSync sy;
  __device__ __forceinline__ bool cell(int k, long long& w, int& slot) const {
    w = (long long)k * n_stage + rank;
    if (w >= n_cells) return false;
    slot = (k % ring_chunks) * n_stage + rank;
    return true;
  }
  __device__ __forceinline__ void wait_slot(int k) const {
    if (k >= ring_chunks) spin_until(sy.gdone + (k - ring_chunks), (unsigned)n_gather_warps, sy.err);
  }
  __device__ __forceinline__ void signal(int k) const {
    __threadfence();
    atomicAdd(sy.staged + k, 1u);
  }
};

struct FusedArgs {
  MmaArgs a;
  CsView cs;
  const double* dphi_lane;
  double* stage;
  GatherArgs g;            // item arrays, plan, staging (per-chunk fields are set in the kernel)
  PreArgs pa;
  const long long* v_chunk_ptr;
  const long long* p_chunk_ptr;
  long long n_cells, n_chunks;
  int n_stage, n_gather_ctas, ring_chunks, fuse_pre;
  FusedSync sy;
};

constexpr size_t fused_smem_bytes() {
  return stage_smem_bytes() > sizeof(double) * GACC * (MTHREADS / 32) ? stage_smem_bytes() : sizeof(double) * GACC * (MTHREADS / 32);
}

constexpr int WBF = 2;   // items per block in the persistent kernel (a chunk has about 8 items per gather warp)

__global__ void __launch_bounds__(MTHREADS, 4) th_fused_kernel(const __grid_constant__ FusedArgs f, const __grid_constant__ BlockView A,
                                                               const __grid_constant__ BlockView Apre) {
  if ((int)blockIdx.x < f.n_stage) {
    FusedSched sc;
    sc.n_cells = f.n_cells;
    sc.n_stage = f.n_stage;
    sc.ring_chunks = f.ring_chunks;
    sc.n_gather_warps = f.n_gather_ctas * (MTHREADS / 32);
    sc.rank = (int)blockIdx.x;
    sc.sy = f.sy;
    stage_cells(f.a, f.cs, f.dphi_lane, f.stage, sc);
    return;
  }
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* acc = smem + warp * GACC;
  for (int k = lane; k < GACC; k += 32) acc[k] = 0.0;
  __syncwarp();
  const long long gw = (long long)((int)blockIdx.x - f.n_stage) * (MTHREADS / 32) + warp, nw = (long long)f.n_gather_ctas * (MTHREADS / 32);
  GatherArgs g = f.g;
  g.ring = f.ring_chunks * f.n_stage;
  for (long long c = 0; c < f.n_chunks; ++c) {
    const long long w0 = c * f.n_stage;
    const unsigned cells = (unsigned)(f.n_cells - w0 < f.n_stage ? f.n_cells - w0 : f.n_stage);
    if (lane == 0) {
      if (c > 0) spin_until(f.sy.gdone + (c - 1), (unsigned)nw, f.sy.err);   // rows shared with the previous chunk are final
      spin_until(f.sy.staged + c, cells, f.sy.err);
    }
    __syncwarp();
    g.v_begin = f.v_chunk_ptr[c];
    g.v_end = f.v_chunk_ptr[c + 1];
    g.p_begin = f.p_chunk_ptr[c];
    g.p_end = f.p_chunk_ptr[c + 1];
    g.w_base = w0;
    g.slot_base = (int)(c % f.ring_chunks) * f.n_stage;
    gather_system<WBF>(g, A, acc, gw, nw, lane);
    if (f.fuse_pre) gather_pre<WBF, ASTR>(g, f.pa, Apre, acc, gw, nw, lane);
    __syncwarp();
    if (lane == 0) {
      __threadfence();
      atomicAdd(f.sy.gdone + c, 1u);
    }
  }
}

template <class T>
int upg(dcp_ctx* ctx, T** dst, const std::vector<T>& v) {
  *dst = nullptr;
  if (v.empty()) return DCP_OK;
  if (cudaMalloc((void**)dst, v.size() * sizeof(T)) != cudaSuccess ||
      cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
    dcp_set_error("gather plan: device allocation / copy failed");
    return DCP_ERR_CUDA;
  }
  return DCP_OK;
}

}  // namespace

void dcp_gather_plan_free(GatherPlan* p) {
  if (!p) return;
  cudaFree(p->v_g0);
  cudaFree(p->v_incptr);
  cudaFree(p->v_flag);
  cudaFree(p->v_inc);
  cudaFree(p->p_g0);
  cudaFree(p->p_incptr);
  cudaFree(p->p_flag);
  cudaFree(p->p_inc);
  cudaFree(p->staging);
  cudaFree(p->dphi_lane);
  cudaFree(p->pre_w);
  cudaFree(p->pre_rest);
  if (p->stream2) cudaStreamDestroy(p->stream2);
  for (int i = 0; i < 2; ++i) {
    if (p->ev_staged[i]) cudaEventDestroy(p->ev_staged[i]);
    if (p->ev_gathered[i]) cudaEventDestroy(p->ev_gathered[i]);
  }
  cudaFree(p->d_v_chunk_ptr);
  cudaFree(p->d_p_chunk_ptr);
  cudaFree(p->sync);
  delete p;
}

// Items (chunk, node) with their incidences (cell of the chunk, local node) for the gather pass.  `cells` is the plan
// order of the masked plan.  Returns DCP_OK with *out == nullptr when the model does not qualify (row longer than the
// accumulators): the caller keeps the reduction path.
int dcp_gather_plan_build(dcp_model* m, const dcp_model_desc* d, const std::vector<int32_t>& cells,
                          const std::vector<uint8_t>& cell_has_constraints, GatherPlan** out) {
  *out = nullptr;
  const int64_t n = (int64_t)cells.size(), n_u = d->nse_block_size[0], n_p = d->nse_block_size[1];
  if (n == 0) return DCP_OK;
  const dcp_csr_desc(*pat)[DCP_MAX_BLOCKS] = d->nse_pattern;
  auto max_len = [](const dcp_csr_desc& P) {
    int64_t mx = 0;
#pragma omp parallel for reduction(max : mx)
    for (int64_t r = 0; r < P.n_rows; ++r) mx = std::max(mx, P.rowptr[r + 1] - P.rowptr[r]);
    return mx;
  };
  if (max_len(pat[0][0]) > LROW || max_len(pat[1][0]) > LROW || max_len(pat[0][1]) > L01) return DCP_OK;
  // Default: one stage + gather launch pair per chunk of DCP_GATHER_CHUNK cells (65 536), staging in HBM.
  // DCP_STAGED_MODE=persistent: one cooperative launch, chunk = one cell per staging CTA, the ring of `ring_chunks`
  // chunks stays in L2.  Measured 7x slower than the launch pairs at refine 5 and 6 (the gather needs two thirds of the
  // SMs at equal occupancy, and blocks of two items expose every load latency); kept as an experiment.
  int64_t chunk = 65536;
  bool fused = false;
  if (const char* e = std::getenv("DCP_STAGED_MODE")) fused = std::string(e) == "persistent";
  int n_stage = 0, n_gather_ctas = 0, ring_chunks = 2;
  if (const char* e = std::getenv("DCP_GATHER_CHUNK")) {
    chunk = std::max<int64_t>(1, std::atoll(e));
    fused = false;
  }
  if (fused) {
    int coop = 0, per_sm = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, m->ctx->device);
    const size_t smem_f = fused_smem_bytes();
    const cudaError_t e1 = cudaFuncSetAttribute(th_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
    const cudaError_t e2 = e1 == cudaSuccess ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, th_fused_kernel, MTHREADS, smem_f) : e1;
    if (!coop || e2 != cudaSuccess || per_sm < 2) {
      if (std::getenv("DCP_VERBOSE"))
        std::fprintf(stderr, "[dcp] persistent assembly kernel not used: cooperative launch %d, %s, %d CTAs per SM, %zu B shared memory\n", coop,
                     cudaGetErrorString(e2), per_sm, smem_f);
      cudaGetLastError();
      fused = false;
    } else {
      const int total = per_sm * m->ctx->sm_count;
      double frac = 0.25;   // share of the CTAs that gather
      if (const char* e = std::getenv("DCP_FUSED_GATHER_FRACTION")) frac = std::min(0.75, std::max(0.05, std::atof(e)));
      if (const char* e = std::getenv("DCP_FUSED_RING_CHUNKS")) ring_chunks = std::min(8, std::max(2, std::atoi(e)));
      n_gather_ctas = std::max(1, (int)(total * frac + 0.5));
      n_stage = total - n_gather_ctas;
      chunk = n_stage;
    }
  }
  chunk = std::min<int64_t>(chunk, n);
  if (fused && chunk < n_stage) {   // fewer cells than staging CTAs: one chunk
    n_stage = (int)chunk;
  }
  if (chunk >= (int64_t(1) << 26)) return DCP_OK;
  const int64_t n_chunks = (n + chunk - 1) / chunk;
  std::vector<int> sys_u(3 * NU), sys_p(NP);
  for (int i = 0; i < ND; ++i) {
    const int f = d->nse_local_field[i], b = d->nse_local_base[i];
    if (f < 3) sys_u[f * NU + b] = i; else sys_p[b] = i;
  }
  // first chunk that touches each node
  std::vector<int32_t> first_u((size_t)n_u, INT_MAX), first_p((size_t)n_p, INT_MAX);
  for (int64_t w = 0; w < n; ++w) {
    const int32_t* idx = d->nse_l2g + (int64_t)cells[w] * ND;
    const int32_t ch = (int32_t)(w / chunk);
    for (int a = 0; a < NU; ++a) {
      int32_t& f = first_u[idx[sys_u[a]]];
      if (ch < f) f = ch;
    }
    for (int a = 0; a < NP; ++a) {
      int32_t& f = first_p[idx[sys_p[a]] - n_u];
      if (ch < f) f = ch;
    }
  }
  struct ChunkItems {
    std::vector<int32_t> g0;
    std::vector<uint32_t> cnt, inc;
    std::vector<uint8_t> flag;
  };
  std::vector<ChunkItems> V((size_t)n_chunks), P((size_t)n_chunks);
  bool too_many = false;   // the gather walks at most 8 incidences per item (hexahedral meshes without extraordinary edges)
#pragma omp parallel
  {
    std::vector<uint64_t> keys;
#pragma omp for schedule(dynamic, 1)
    for (int64_t ch = 0; ch < n_chunks; ++ch) {
      const int64_t w0 = ch * chunk, w1 = std::min(n, w0 + chunk);
      for (int pass = 0; pass < 2; ++pass) {
        const int nl = pass == 0 ? NU : NP;
        keys.clear();
        keys.reserve((size_t)(w1 - w0) * nl);
        for (int64_t w = w0; w < w1; ++w) {
          const int32_t* idx = d->nse_l2g + (int64_t)cells[w] * ND;
          for (int a = 0; a < nl; ++a) {
            const uint64_t g0 = pass == 0 ? (uint64_t)idx[sys_u[a]] : (uint64_t)(idx[sys_p[a]] - n_u);
            keys.push_back((g0 << 32) | (cell_has_constraints[w] ? (uint64_t)INC_CS : 0) | ((uint64_t)(w - w0) << 5) | (uint64_t)a);
          }
        }
        std::sort(keys.begin(), keys.end());
        ChunkItems& I = pass == 0 ? V[ch] : P[ch];
        const std::vector<int32_t>& fst = pass == 0 ? first_u : first_p;
        for (size_t k = 0; k < keys.size();) {
          const uint32_t g0 = (uint32_t)(keys[k] >> 32);
          size_t e = k;
          while (e < keys.size() && (uint32_t)(keys[e] >> 32) == g0) {
            I.inc.push_back((uint32_t)(keys[e] & 0xffffffffu));
            ++e;
          }
          I.g0.push_back((int32_t)g0);
          I.cnt.push_back((uint32_t)(e - k));
          if (e - k > 8) too_many = true;
          I.flag.push_back(fst[g0] == (int32_t)ch ? 1 : 0);
          k = e;
        }
      }
    }
  }
  if (too_many) return DCP_OK;
  GatherPlan* G = new GatherPlan;
  G->chunk = chunk;
  G->n_chunks = n_chunks;
  G->n_cells = n;
  G->fused = fused;
  G->n_stage = n_stage;
  G->n_gather_ctas = n_gather_ctas;
  G->ring_chunks = ring_chunks;
  if (std::getenv("DCP_VERBOSE"))
    std::fprintf(stderr, "[dcp] staged assembly: %s, %lld cells in %lld chunks of %lld, %d staging + %d gathering CTAs, ring of %d chunks\n",
                 fused ? "persistent kernel" : "one launch pair per chunk", (long long)n, (long long)n_chunks, (long long)chunk, n_stage,
                 n_gather_ctas, ring_chunks);
  dcp_ctx* ctx = m->ctx;
  int rc = DCP_OK;
  for (int pass = 0; pass < 2 && rc == DCP_OK; ++pass) {
    std::vector<ChunkItems>& L = pass == 0 ? V : P;
    std::vector<int64_t>& cptr = pass == 0 ? G->v_chunk_ptr : G->p_chunk_ptr;
    cptr.assign((size_t)n_chunks + 1, 0);
    int64_t n_inc = 0;
    for (int64_t ch = 0; ch < n_chunks; ++ch) {
      cptr[ch + 1] = cptr[ch] + (int64_t)L[ch].g0.size();
      n_inc += (int64_t)L[ch].inc.size();
    }
    if (n_inc >= (int64_t(1) << 32)) {
      dcp_gather_plan_free(G);
      return DCP_OK;
    }
    const int64_t n_items = cptr[n_chunks];
    std::vector<int32_t> g0((size_t)n_items);
    std::vector<uint32_t> incptr((size_t)n_items + 1), inc((size_t)n_inc);
    std::vector<uint8_t> flag((size_t)n_items);
    int64_t ip = 0;
    for (int64_t ch = 0; ch < n_chunks; ++ch) {
      ChunkItems& I = L[ch];
      const int64_t b = cptr[ch];
      std::copy(I.g0.begin(), I.g0.end(), g0.begin() + b);
      std::copy(I.flag.begin(), I.flag.end(), flag.begin() + b);
      for (size_t k = 0; k < I.cnt.size(); ++k) {
        incptr[b + k] = (uint32_t)ip;
        ip += I.cnt[k];
      }
      std::copy(I.inc.begin(), I.inc.end(), inc.begin() + (ip - (int64_t)I.inc.size()));
      ChunkItems().g0.swap(I.g0);
      std::vector<uint32_t>().swap(I.inc);
      std::vector<uint32_t>().swap(I.cnt);
      std::vector<uint8_t>().swap(I.flag);
    }
    incptr[n_items] = (uint32_t)ip;
    rc = upg(ctx, pass == 0 ? &G->v_g0 : &G->p_g0, g0);
    if (rc == DCP_OK) rc = upg(ctx, pass == 0 ? &G->v_incptr : &G->p_incptr, incptr);
    if (rc == DCP_OK) rc = upg(ctx, pass == 0 ? &G->v_inc : &G->p_inc, inc);
    if (rc == DCP_OK) rc = upg(ctx, pass == 0 ? &G->v_flag : &G->p_flag, flag);
    cudaStreamSynchronize(ctx->stream);
  }
  if (rc == DCP_OK) {
    if (cudaMalloc((void**)&G->dphi_lane, sizeof(double) * 7 * 3 * 4 * NQ) != cudaSuccess) {
      cudaGetLastError();
      dcp_set_error("gather plan: table allocation failed");
      rc = DCP_ERR_CUDA;
    } else {
      dphi_lane_kernel<<<(7 * 3 * 4 * NQ + 255) / 256, 256, 0, ctx->stream>>>(m->dphi_u_qn, G->dphi_lane);
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = DCP_ERR_CUDA;
    }
  }
  if (rc == DCP_OK && fused) {
    rc = upg(ctx, reinterpret_cast<int64_t**>(&G->d_v_chunk_ptr), G->v_chunk_ptr);
    if (rc == DCP_OK) rc = upg(ctx, reinterpret_cast<int64_t**>(&G->d_p_chunk_ptr), G->p_chunk_ptr);
    if (rc == DCP_OK && cudaMalloc((void**)&G->sync, sizeof(unsigned) * 2 * (size_t)n_chunks) != cudaSuccess) rc = DCP_ERR_CUDA;
    cudaStreamSynchronize(ctx->stream);
  }
  // DCP_STAGED_OVERLAP=1: two streams (see dcp_launch_th_staged).  Measured slower at refine 5 for every split of the SMs
  // (19.2 ms with 2 + 1 CTAs per SM, 15.1 ms with full grids on both streams, against 14.1 ms on one stream): both sides
  // are occupancy-starved already, halving their resident warps costs more than the overlap returns.
  bool overlap = false;
  if (const char* e = std::getenv("DCP_STAGED_OVERLAP")) overlap = !fused && n_chunks > 1 && std::atoi(e) != 0;
  if (overlap) {
    if (const char* e = std::getenv("DCP_OVERLAP_STAGE_CTAS")) G->overlap_stage_ctas = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("DCP_OVERLAP_GATHER_CTAS")) G->overlap_gather_ctas = std::max(1, std::atoi(e));
    bool ok = cudaStreamCreateWithFlags(&G->stream2, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i)
      ok = cudaEventCreateWithFlags(&G->ev_staged[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&G->ev_gathered[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      cudaGetLastError();
      if (G->stream2) cudaStreamDestroy(G->stream2);
      G->stream2 = nullptr;
      overlap = false;
    }
  }
  const size_t staging_cells = fused ? (size_t)ring_chunks * (size_t)n_stage : (size_t)chunk * (overlap ? 2 : 1);
  if (rc == DCP_OK && cudaMalloc((void**)&G->staging, sizeof(double) * (size_t)REC * staging_cells) != cudaSuccess) {
    cudaGetLastError();
    dcp_set_error("gather plan: staging allocation failed");
    rc = DCP_ERR_CUDA;
  }
  if (rc != DCP_OK) {
    dcp_gather_plan_free(G);
    return rc;
  }
  *out = G;
  return DCP_OK;
}

// Fused preconditioner: map the system plan's cells to the preconditioner plan.  Not attached (the preconditioner keeps
// its own pass) when a row of that matrix is longer than the accumulators.
int dcp_gather_plan_attach_pre(dcp_model* m, const dcp_model_desc* d, GatherPlan* G, const MaskedPlan* nse_plan, const MaskedPlan* pre_plan) {
  if (!G || !nse_plan || !pre_plan || std::getenv("DCP_NO_FUSED_PRECONDITIONER")) return DCP_OK;
  const dcp_csr_desc(*pat)[DCP_MAX_BLOCKS] = d->pre_pattern;
  for (const dcp_csr_desc* P : {&pat[0][0], &pat[1][1]}) {
    int64_t mx = 0;
#pragma omp parallel for reduction(max : mx)
    for (int64_t r = 0; r < P->n_rows; ++r) mx = std::max(mx, P->rowptr[r + 1] - P->rowptr[r]);
    if (mx > LROW) return DCP_OK;
  }
  if (pre_plan->max_off_plain >= PSTRIDE) return DCP_OK;   // a gathered entry would not fit the compact accumulators
  {
    int64_t mx = 0;   // pressure mass rows go through the same accumulators
    const dcp_csr_desc& P11 = pat[1][1];
#pragma omp parallel for reduction(max : mx)
    for (int64_t r = 0; r < P11.n_rows; ++r) mx = std::max(mx, P11.rowptr[r + 1] - P11.rowptr[r]);
    if (mx > PSTRIDE) return DCP_OK;
  }
  std::vector<int32_t> of_cell((size_t)d->n_cells, -1);
  for (size_t i = 0; i < pre_plan->h_cells.size(); ++i) of_cell[pre_plan->h_cells[i]] = (int32_t)i;
  std::vector<long long> pre_w(nse_plan->h_cells.size(), -1ll);
  std::vector<int32_t> rest;
  for (size_t w = 0; w < nse_plan->h_cells.size(); ++w) {
    const int32_t wp = of_cell[nse_plan->h_cells[w]];
    if (wp >= 0 && pre_plan->h_nnf_idx[wp] < 0) {
      const long long hi = ((long long)(pre_plan->h_wide_idx[wp] + 1) << 1) | (pre_plan->h_cflag[wp] ? 1 : 0);
      pre_w[w] = (hi << 32) | (unsigned)wp;
    }
  }
  for (size_t i = 0; i < pre_plan->h_cells.size(); ++i)
    if (pre_plan->h_nnf_idx[i] >= 0) rest.push_back((int32_t)i);
  dcp_ctx* ctx = m->ctx;
  int rc = upg(ctx, &G->pre_w, pre_w);
  if (rc == DCP_OK) rc = upg(ctx, &G->pre_rest, rest);
  if (rc != DCP_OK) return rc;
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  G->n_pre_rest = (int64_t)rest.size();
  G->has_pre = true;
  return DCP_OK;
}

// NSE system, write-once: per chunk of plan cells, stage then gather (stream-ordered).
int dcp_launch_th_staged(dcp_model* m, const dcp_params& p, const MaskedPlan* plan, const double* old_nse, const double* old_temp) {
  dcp_ctx* ctx = m->ctx;
  const GatherPlan* G = plan->gather;
  MmaArgs a;
  a.n_fast = plan->n;
  a.cells = plan->cells;
  a.pos = plan->pos;
  a.nmask = plan->nmask;
  a.pos_wide = plan->pos_wide;
  a.pos9 = plan->pos9;
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
  a.n_u = m->nse.start[1];
  a.prm = p;
  const size_t smem_s = stage_smem_bytes(), smem_g = sizeof(double) * GACC * GWARPS, smem_gs = sizeof(double) * GSWARP * GSW,
               smem_p = sizeof(double) * PACC * GWARPS;
  DCP_CUDA(cudaFuncSetAttribute(th_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
  const bool bulk = std::getenv("DCP_GATHER_BULK") != nullptr;
  DCP_CUDA(cudaFuncSetAttribute(th_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
  DCP_CUDA(cudaFuncSetAttribute(th_gather_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_gs));
  int per_sm_s = 1, per_sm_g = 1, per_sm_p = 1;
  DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_s, th_stage_kernel, MTHREADS, smem_s));
  if (bulk)
    DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_g, th_gather_bulk_kernel, GSW * 32, smem_gs));
  else
    DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_g, th_gather_kernel, GWARPS * 32, smem_g));
  DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_p, th_pre_gather_kernel, GWARPS * 32, smem_p));
  per_sm_s = std::max(per_sm_s, 1);
  per_sm_g = std::max(per_sm_g, 1);
  per_sm_p = std::max(per_sm_p, 1);
  GatherArgs g;
  g.v_g0 = G->v_g0;
  g.v_incptr = G->v_incptr;
  g.v_flag = G->v_flag;
  g.v_inc = G->v_inc;
  g.p_g0 = G->p_g0;
  g.p_incptr = G->p_incptr;
  g.p_flag = G->p_flag;
  g.p_inc = G->p_inc;
  g.ring = (int)G->chunk;
  g.pos = plan->pos;
  g.nmask = plan->nmask;
  g.stage = G->staging;
  const BlockView A = make_view(m->nse);
  const CsView cs = make_view(m->nse_cs);
  const bool fuse_pre = G->has_pre && m->masked_pre;
  if (G->fused) {
    FusedArgs f{};
    f.a = a;
    f.cs = cs;
    f.dphi_lane = G->dphi_lane;
    f.stage = G->staging;
    f.g = g;
    f.g.w_base = 0;
    f.g.slot_base = 0;
    f.g.v_begin = f.g.v_end = f.g.p_begin = f.g.p_end = 0;
    if (fuse_pre) {
      f.pa.pre_w = G->pre_w;
      f.pa.pos = m->masked_pre->pos;
      f.pa.nmask = m->masked_pre->nmask;
      f.pa.pos_wide = m->masked_pre->pos_wide;
    }
    f.v_chunk_ptr = G->d_v_chunk_ptr;
    f.p_chunk_ptr = G->d_p_chunk_ptr;
    f.n_cells = plan->n;
    f.n_chunks = G->n_chunks;
    f.n_stage = G->n_stage;
    f.n_gather_ctas = G->n_gather_ctas;
    f.ring_chunks = G->ring_chunks;
    f.fuse_pre = fuse_pre ? 1 : 0;
    f.sy.staged = G->sync;
    f.sy.gdone = G->sync + G->n_chunks;
    f.sy.err = ctx->d_err;
    BlockView Av = A, Apv = make_view(m->pre);
    DCP_CUDA(cudaMemsetAsync(G->sync, 0, sizeof(unsigned) * 2 * (size_t)G->n_chunks, ctx->stream));
    const size_t smem_f = fused_smem_bytes();
    DCP_CUDA(cudaFuncSetAttribute(th_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
    void* params[] = {(void*)&f, (void*)&Av, (void*)&Apv};
    DCP_CUDA(cudaLaunchCooperativeKernel((const void*)th_fused_kernel, dim3((unsigned)(G->n_stage + G->n_gather_ctas)), dim3(MTHREADS), params,
                                         smem_f, ctx->stream));
    ctx->launches++;
    if (fuse_pre) {
      if (G->n_pre_rest > 0) DCP_TRY(dcp_launch_th_mma(m, p, false, m->masked_pre, nullptr, nullptr, G->pre_rest, G->n_pre_rest));
      if (m->masked_pre->n_other > 0)
        DCP_TRY(dcp_launch_th_cells(m, p, false, nullptr, nullptr, m->masked_pre->other_cells, m->masked_pre->n_other, false));
      m->pre_fused_valid = true;
      m->pre_fused_dt = p.dt;
      m->pre_fused_inv_re = p.inv_re;
    }
    DCP_CUDA(cudaGetLastError());
    return DCP_OK;
  }
  PreArgs pa{};
  BlockView Apre = make_view(m->pre);
  if (fuse_pre) {
    pa.pre_w = G->pre_w;
    pa.pos = m->masked_pre->pos;
    pa.nmask = m->masked_pre->nmask;
    pa.pos_wide = m->masked_pre->pos_wide;
    DCP_CUDA(cudaFuncSetAttribute(th_pre_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
  }
  // Two streams: the stage pass of chunk c + 1 (tensor / LSU bound, writes) runs next to the gather passes of chunk c
  // (DRAM bound, reads and writes).  Both kernels stride over their work with a fixed grid, so the grids set the split
  // of the SMs: 2 staging CTAs + 1 gathering CTA per SM fit together (registers, shared memory).  The staging buffer has
  // two halves; events order reuse.  Opt-in (DCP_STAGED_OVERLAP=1): measured slower than one stream with full grids.
  const bool overlap = G->stream2 != nullptr && G->n_chunks > 1;
  cudaStream_t s_stage = ctx->stream, s_gather = overlap ? G->stream2 : ctx->stream;
  const int stage_per_sm = overlap ? std::min(per_sm_s, G->overlap_stage_ctas) : per_sm_s;
  const int gather_per_sm = overlap ? std::min(per_sm_g, G->overlap_gather_ctas) : per_sm_g;
  const int pre_per_sm = overlap ? std::min(per_sm_p, G->overlap_gather_ctas) : per_sm_p;
  for (int64_t ch = 0; ch < G->n_chunks; ++ch) {
    const int half = overlap ? (int)(ch & 1) : 0;
    const long long w0 = ch * G->chunk, w1 = std::min<long long>(plan->n, w0 + G->chunk);
    long long grid = std::min<long long>((long long)ctx->sm_count * stage_per_sm, w1 - w0);
    if (overlap && ch >= 2) DCP_CUDA(cudaStreamWaitEvent(s_stage, G->ev_gathered[half], 0));   // the half is free again
    double* half_base = G->staging + (size_t)half * (size_t)G->chunk * REC;
    th_stage_kernel<<<(unsigned)grid, MTHREADS, smem_s, s_stage>>>(a, cs, G->dphi_lane, half_base, w0, w1, 0, (int)G->chunk);
    if (overlap) {
      DCP_CUDA(cudaEventRecord(G->ev_staged[half], s_stage));
      DCP_CUDA(cudaStreamWaitEvent(s_gather, G->ev_staged[half], 0));
    }
    g.stage = half_base;
    g.v_begin = G->v_chunk_ptr[ch];
    g.v_end = G->v_chunk_ptr[ch + 1];
    g.p_begin = G->p_chunk_ptr[ch];
    g.p_end = G->p_chunk_ptr[ch + 1];
    g.w_base = w0;
    g.slot_base = 0;
    const long long items = (g.v_end - g.v_begin) + (g.p_end - g.p_begin);
    const long long blocks = (items + WB - 1) / WB;   // a warp takes blocks of WB items
    const int gwarps = bulk ? GSW : GWARPS;
    grid = std::min<long long>((long long)ctx->sm_count * gather_per_sm, (blocks + gwarps - 1) / gwarps);
    if (grid > 0) {
      if (bulk)
        th_gather_bulk_kernel<<<(unsigned)grid, GSW * 32, smem_gs, s_gather>>>(g, A);
      else
        th_gather_kernel<<<(unsigned)grid, GWARPS * 32, smem_g, s_gather>>>(g, A);
    }
    ctx->launches += 2;
    grid = std::min<long long>((long long)ctx->sm_count * pre_per_sm, (blocks + GWARPS - 1) / GWARPS);
    if (fuse_pre && grid > 0) {
      th_pre_gather_kernel<<<(unsigned)grid, GWARPS * 32, smem_p, s_gather>>>(g, pa, Apre);
      ctx->launches++;
    }
    if (overlap) DCP_CUDA(cudaEventRecord(G->ev_gathered[half], s_gather));
  }
  if (overlap) {   // the context's stream continues after the last gathers
    DCP_CUDA(cudaStreamWaitEvent(s_stage, G->ev_gathered[0], 0));
    DCP_CUDA(cudaStreamWaitEvent(s_stage, G->ev_gathered[1], 0));
  }
  if (fuse_pre) {
    // the cells the second gather skipped: no-normal-flux cells through the reduction kernel, cells outside the
    // preconditioner's plan through the general kernel -- both add to rows the gather has already stored
    if (G->n_pre_rest > 0) DCP_TRY(dcp_launch_th_mma(m, p, false, m->masked_pre, nullptr, nullptr, G->pre_rest, G->n_pre_rest));
    if (m->masked_pre->n_other > 0)
      DCP_TRY(dcp_launch_th_cells(m, p, false, nullptr, nullptr, m->masked_pre->other_cells, m->masked_pre->n_other, false));
    m->pre_fused_valid = true;
    m->pre_fused_dt = p.dt;
    m->pre_fused_inv_re = p.inv_re;
  }
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}
