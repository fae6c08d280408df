// Classic 3-D Taylor-Hood NSE system and its preconditioner: write-once assembly (DCP_STRATEGY_STAGED).
//
// Same integrals and the same tensor-core contraction as assemble_th_mma.cu (reference:
// include/core/boussinesq_model.tpp:550-687 and :421-476), but the scatter of copy_local_to_global_nse_system
// (:677-687) is split in two so that no CSR value is ever reduced with red.global.add.f64 and the matrices need no
// zero-fill:
//
//   1. th_stage_kernel: per cell, the constraint-resolved 3x3 blocks L[(a,.),(b,.)], the velocity-pressure coupling and
//      (for the preconditioner) m + nu k and the pressure mass block leave the DMMA registers as 64-byte row segments.
//      The staging buffer is NODE-MAJOR: the row a cell contributes to a node is written to the slot of that
//      (node, cell) incidence, and the slots of one node are consecutive.  Velocity-node row (304 doubles):
//      [m + nu k: 28][component c: [d][28 column nodes] + 8 pressure columns] x 3; pressure-node row (92 doubles):
//      [c][28] + 8 pressure-mass columns.
//   2. th_gather_kernel: one warp per (chunk, node) item.  Everything the item needs lies in one contiguous piece of the
//      staging buffer and of a static record stream (header + row positions per incidence); the warp streams both with
//      TMA bulk copies (cp.async.bulk + mbarrier) through a ring of RD slots in shared memory, one component row at a
//      time (three sweeps over the item's incidences), adds the segments at the row positions into shared-memory
//      accumulators and writes every CSR row of nse_matrix and of nse_preconditioner_matrix once, as one bulk copy from
//      shared memory (first chunk that touches the node: store of the whole row, which doubles as the zero-fill; later
//      chunks: bulk reduce-add at the L2; chunks are stream-ordered and a row belongs to one warp).
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
constexpr int CSEG = 3 * BW + NP;           // one component of a velocity-node row: [d][28] then 8 pressure columns (92)
constexpr int VROW = BW + 3 * CSEG;         // staged velocity-node row: [m + nu k][c = 0][c = 1][c = 2]  (304)
constexpr int PROW = 3 * BW + NP;           // staged pressure-node row: [c][28] then the pressure mass columns (92)
constexpr int REC = NU * VROW + NP * PROW;  // staged doubles per cell (8 944)
constexpr int SLOTS = 36;                   // slot table row: 27 velocity-node slots, 8 pressure-node slots, pad
constexpr int LROW = 384;                   // longest block(0,0) / block(1,0) row the gather accumulators hold
constexpr int ACC0 = 388;                   // accumulator of one block(0,0) / block(1,0) row
constexpr int L01 = 32;                     // longest block(0,1) row
constexpr int RD = 3;                       // staged rows in flight per gather warp
constexpr int GW = 1;                       // warps per CTA of the gather (the warps are independent; shared memory sets how many fit an SM)
static_assert((VROW * 8) % 32 == 0 && (PROW * 8) % 32 == 0 && (BW * 8) % 32 == 0 && (CSEG * 8) % 32 == 0, "sector alignment of the staged segments");
static_assert(ACC0 % 2 == 0 && L01 % 2 == 0, "16-byte alignment of the ring slots");

// ---- stage: contraction + constraint epilogue, coalesced write of the staged rows --------------------------------
// The DMMA accumulator layout does the transposition: lane (frow = lane / 4, fk = lane % 4) holds the 3x3 blocks of
// (row node 8 ta + frow, column nodes 8 tb + 2 fk + {0,1}).  Direct orientation: for every (c, d) the warp writes
// 8 rows x 64 contiguous bytes with one 16-byte store per lane; transposed orientation (row node b, column node a,
// entry [d][c]): for every (c, d, jj) 4 rows x 64 contiguous bytes with one 8-byte store per lane.  All segments start
// on 32-byte sector boundaries (BW = 28).
__constant__ unsigned char c_task_ta[10] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3};
__constant__ unsigned char c_task_tb[10] = {0, 1, 2, 3, 1, 2, 3, 2, 3, 3};

struct StageOut {
  double* vstage;            // velocity-node rows of the chunk
  double* pstage;            // pressure-node rows of the chunk
  const unsigned* slots;     // [plan cell][SLOTS]: slot of every local node inside its chunk
};

__global__ void __launch_bounds__(MTHREADS, 4)
th_stage_kernel(MmaArgs a, CsView cs, const double* __restrict__ dphi_lane, StageOut so, long long w_begin, long long w_end) {
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
  unsigned* sslot2 = (unsigned*)(sidt2 + 2 * 28);              // 2 x SLOTS
  int* sys_u = (int*)(sslot2 + 2 * SLOTS);                     // 3*NU
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
  // mapping record, masks, dof indices and staging slots of a cell, copied asynchronously into buffer `b`
  auto issue_raw = [&](long long w, int b) {
    const long long cell = a.cells[w];
    const double* g = a.geom + cell * GS;
    for (int i = tid; i < GS; i += nt) cp_async8(sgeo2 + b * (GS + 1) + i, g + i);
    if (tid < MSTR / 8) cp_async8(snm2 + b * MSTR + 8 * tid, a.nmask + w * MSTR + 8 * tid);
    if (tid >= 32 && tid - 32 < SLOTS / 2) cp_async8(sslot2 + b * SLOTS + 2 * (tid - 32), so.slots + w * SLOTS + 2 * (tid - 32));
    for (int i = tid; i < ND; i += nt) cp_async4(sidx2 + b * IDS + i, a.l2g + cell * ND + i);
    if (do_rhs)
      for (int i = tid; i < a.ndt; i += nt) cp_async4(sidt2 + b * 28 + i, a.l2g_t + cell * a.ndt + i);
  };
  const int frow = lane >> 2, fk = lane & 3;

  int buf = 0;
  long long w = w_begin + blockIdx.x;
  if (w < w_end) issue_raw(w, 0);
  cp_async_commit();
  for (; w < w_end; w += gridDim.x, buf ^= 1) {
    const long long w_next = w + gridDim.x;
    const double* sgeo = sgeo2 + buf * (GS + 1);
    const unsigned char* snm = snm2 + buf * MSTR;
    const int* sidx = sidx2 + buf * IDS;
    const int* sidt = sidt2 + buf * 28;
    const unsigned* sslot = sslot2 + buf * SLOTS;
    auto node_cs = [&](int n) {
      NodeCs c;
      c.mask = snm[n];
      c.k = skc[n];
      c.w0 = swt[3 * n];
      c.w1 = swt[3 * n + 1];
      c.w2 = swt[3 * n + 2];
      return c;
    };
    auto vrow = [&](int n) { return so.vstage + (size_t)sslot[n] * VROW; };
    auto prow = [&](int pn) { return so.pstage + (size_t)sslot[NU + pn] * PROW; };
    cp_async_wait<0>();
    __syncthreads();   // this cell's raw inputs have landed; every warp is done with the previous cell
    if (do_rhs) {
      for (int i = tid; i < 3 * NU; i += nt) {
        const int c = i / NU, n = i - c * NU;
        cp_async8(sUc + c * BW + n, a.old_nse + sidx[sys_u[i]]);
      }
      for (int i = tid; i < a.ndt; i += nt) cp_async8(sTn + i, a.old_temp + sidt[i]);
    }
    cp_async_commit();
    if (w_next < w_end) issue_raw(w_next, buf ^ 1);   // the CTA's next cell, while this one is computed
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
            __stcg(reinterpret_cast<double2*>(prow(frow) + 3 * BW + 2 * fk), make_double2(c0, c1));
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
          double* row = vrow(na) + 8 * tb_ + 2 * fk;
          __stcg(reinterpret_cast<double2*>(row), make_double2(dgj[0], dgj[1]));
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int d = 0; d < 3; ++d)
              __stcg(reinterpret_cast<double2*>(row + BW + c * CSEG + d * BW), make_double2(Fj[0][c * 3 + d], Fj[1][c * 3 + d]));
        }
        if (ta != tb_) {
          // transposed orientation: row node b, column node a, entry [d][c] = F[c][d]
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int nb = 8 * tb_ + 2 * fk + jj;
            if (nb < NU) {   // column 8ta + frow <= 23 is always a real node here (ta < tb)
              double* row = vrow(nb) + 8 * ta + frow;
              __stcg(row, dgj[jj]);
#pragma unroll
              for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int d = 0; d < 3; ++d) __stcg(row + BW + d * CSEG + c * BW, Fj[jj][c * 3 + d]);
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
          double* vr = vrow(na) + BW + 3 * BW + 2 * fk;
          double* p0 = prow(2 * fk) + na;
          double* p1 = prow(2 * fk + 1) + na;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            // velocity-node row: component c, pressure columns; pressure-node rows: [c][a]
            __stcg(reinterpret_cast<double2*>(vr + c * CSEG), make_double2(sv[0][c], sv[1][c]));
            __stcg(p0 + c * BW, sv[0][c]);
            __stcg(p1 + c * BW, sv[1][c]);
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
  }
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
         sizeof(int) * (2 * IDS + 2 * 28 + 2 * SLOTS + 3 * NU + NP) + 28 + 36;
}

// ---- gather: one warp per (chunk, node) ---------------------------------------------------------------------------
// Everything an item needs arrives through one sequential stream per kind: the staged rows of its incidences (written
// by the stage pass) and, in the same order, a static record per incidence (built once with the plan): a 16-byte
// header, the positions of the incidence's column nodes inside the rows of nse_matrix and inside the rows of the
// preconditioner.  A warp owns a contiguous run of items; one lane feeds a ring of RD slots with two bulk copies per
// incidence (staged row, static record), all lanes add a landed slot into shared-memory accumulators and write the rows
// of a finished item.
struct IncHdr {
  unsigned e;    // (cell inside the chunk << 5) | local node, bit 31: the cell holds constrained velocity dofs
  int wp;        // index of the cell in the preconditioner's plan; -1: not gathered for that matrix
  unsigned hi;   // bit 0: that plan's cell holds constrained dofs, bits 1..27: index of its wide table + 1, bit 29: the item has
                 // such cells (general path for the preconditioner), bit 30: first chunk that touches the node, bit 31: last
                 // incidence of the item
  int g0n;       // first dof / row of the next item (-1: none): its row starts are fetched one item ahead
};
constexpr unsigned HI_NP = 1u << 29, HI_FIRST = 1u << 30, HI_LAST = 1u << 31, HI_WIDE = 0x0ffffffeu;
constexpr int META = 18;                    // doubles per static record: header (16 B), positions (72 B), preconditioner positions (56 B)
constexpr int MPOS = 8, MPPOS = 8 + 36;     // u16 offsets of the two position rows inside the record
constexpr int GSLOT = BW + CSEG + META;     // ring slot (138 doubles): [m + nu k][one component segment or a pressure-node row][record]

struct GatherArgs {
  const int* v_g0;
  const unsigned* v_incptr;
  const double* v_meta;                      // [incidence][META]
  const int* p_g0;
  const unsigned* p_incptr;
  const double* p_meta;
  long long v_begin, v_end, p_begin, p_end;  // item ranges of this chunk
  long long w_base;                          // first plan cell of the chunk
  const unsigned char* nmask;
  const double* vstage;                      // the chunk's staged velocity-node rows, in incidence order
  const double* pstage;                      // ... pressure-node rows
  // fused preconditioner: masks and per-component positions of the cells on the general path
  const unsigned char* pnmask;
  const unsigned short* ppos_wide;
  int pstr;                                  // length of the accumulator of one preconditioner row (even)
  int wb;                                    // items per block: a warp takes blocks round-robin
};

constexpr unsigned INC_CS = 0x80000000u;     // incidence word: the cell holds constrained velocity dofs
constexpr unsigned FULLM = 0xffffffffu;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
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

// ---- writing a finished row -----------------------------------------------------------------------------------------
// Small rows and the odd cases go through the lanes (flush_row); a long row leaves shared memory as ONE bulk copy
// (cp.async.bulk shared -> global; rows that earlier chunks already touched: cp.reduce.async.bulk .add.f64, the addition
// happens at the L2).  Bulk copies need 16-byte aligned ends on both sides: the accumulator row starts at index
// (row start & 1), so element k of the row is acc[par + k] and the 16-byte phase of the two sides agree; a leading /
// trailing single element is written by a lane.
template <bool ZERO>
__device__ __forceinline__ void flush_row(double* __restrict__ out, double* acc, int len, int cap, bool first, int lane) {
  const int n = len < cap ? len : cap;
  if (first) {
    for (int k = lane; k < n; k += 32) {
      __stcs(out + k, acc[k]);
      if (ZERO) acc[k] = 0.0;
    }
    for (int k = cap + lane; k < len; k += 32) __stcs(out + k, 0.0);
  } else {
    for (int k = lane; k < n; k += 32) {
      __stcs(out + k, __ldcs(out + k) + acc[k]);
      if (ZERO) acc[k] = 0.0;
    }
  }
}
__device__ __forceinline__ void bulk_store(double* dst, unsigned src_s, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_add(double* dst, unsigned src_s, unsigned bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(src_s), "r"(bytes) : "memory");
}
// `acc` (shared-memory address acc_s) holds the row from index par = (row start & 1) on; n = number of accumulated entries
// (<= len; the rest of the row, if any, is stored as zeros by the first chunk).  Issued by lane 0; the caller waits
// (bulk_wait_read) before the accumulators are cleared.
__device__ __forceinline__ void bulk_row(double* __restrict__ out, const double* acc, unsigned acc_s, int par, int n, bool first, int lane) {
  if (lane == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the lanes' accumulator updates come first
    const int nb = (n - par) & ~1;
    if (par && n > 0) out[0] = first ? acc[1] : out[0] + acc[1];
    if (nb > 0) {
      if (first)
        bulk_store(out + par, acc_s + 16 * par, 8 * nb);
      else
        bulk_add(out + par, acc_s + 16 * par, 8 * nb);
    }
    if (par + nb < n) out[n - 1] = first ? acc[par + n - 1] : out[n - 1] + acc[par + n - 1];
  }
}
__device__ __forceinline__ void bulk_wait_read(int lane) {
  if (lane == 0) {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  __syncwarp();
}
// clear acc[0 .. n) (n rounded up to 2), 16 bytes per lane
__device__ __forceinline__ void clear_acc(double* acc, int n, int lane) {
#pragma unroll 1
  for (int k = 2 * lane; k < n; k += 64) *reinterpret_cast<double2*>(acc + k) = make_double2(0.0, 0.0);
}

// Per-warp shared memory: acc [ACC0] (one row of block(0,0) / block(1,0); the last three entries absorb the idle lanes'
// updates), acc01 [L01] (likewise its last entry), accP [pstr] (one preconditioner row, likewise), accd [4] (diagonal of a
// constrained row), the ring and its mbarriers.
struct WarpMem {
  double *acc, *acc01, *accP, *accd, *ring;
  unsigned long long* bars;
};

// Walk this warp's blocks of items.  VEL: velocity nodes -- rows (g0 + c) of block(0,0), block(0,1) and of the
// preconditioner's block(0,0): an item is swept three times over its incidences, sweep c accumulates row component c
// from the component's segment of the staged rows (the accumulators hold one row: three times as many warps fit an SM
// as with all three rows at once, and the pass is bound by how many warps work on it).  The three preconditioner rows of
// a node are equal unless a cell has masks or per-component positions: they are accumulated in sweep 0 and stored
// three times.  Else pressure nodes: one sweep, a row of block(1,0) and of the preconditioner's block(1,1).
template <bool VEL, bool PRE>
__device__ __forceinline__ void gather_items(const GatherArgs& g, const BlockView& A, const BlockView& Ap, const WarpMem& m, unsigned& phases,
                                             long long gw, long long nw, int lane) {
  constexpr int ROWD = VEL ? VROW : PROW;
  constexpr int NSWEEP = VEL ? 3 : 1;
  const int* __restrict__ g0s = VEL ? g.v_g0 : g.p_g0;
  const unsigned* __restrict__ incptr = VEL ? g.v_incptr : g.p_incptr;
  const double* __restrict__ meta = VEL ? g.v_meta : g.p_meta;
  // the chunk's staged rows start at incidence i0: staging slot = incidence - i0
  const double* __restrict__ stage = (VEL ? g.vstage : g.pstage) - (size_t)((long long)(VEL ? NU : NP) * g.w_base) * ROWD;
  const long long begin = VEL ? g.v_begin : g.p_begin, end = VEL ? g.v_end : g.p_end;
  const long long* rpa = VEL ? A.rowptr[0][0] : A.rowptr[1][0];
  const long long* rpb = A.rowptr[0][1];
  const long long* rpq = VEL ? Ap.rowptr[0][0] : Ap.rowptr[1][1];
  double* va = VEL ? A.val[0][0] : A.val[1][0];
  double* vb = A.val[0][1];
  double* vq = VEL ? Ap.val[0][0] : Ap.val[1][1];
  double* acc = m.acc;
  double* acc01 = m.acc01;
  double* accP = m.accP;
  double* accd = m.accd;
  const unsigned ring_s = smem_u32(m.ring), bars_s = smem_u32(m.bars), acc_s = smem_u32(m.acc), accP_s = smem_u32(m.accP);
  const int ptrash = g.pstr - 1;
  // row starts of an item: lanes 0..3 block(0,0) (or 0..1 block(1,0)), 4..7 block(0,1), 8..11 the preconditioner's block
  auto rows = [&](int g0) {
    if (VEL) return lane < 4 ? rpa[g0 + lane] : (lane < 8 ? rpb[g0 + lane - 4] : (PRE && lane < 12 ? rpq[g0 + lane - 8] : 0ll));
    return lane < 2 ? rpa[g0 + lane] : (PRE && lane >= 8 && lane < 10 ? rpq[g0 + lane - 8] : 0ll);
  };
  // sweep c of incidence i into ring slot u: the component's segment (sweep 0 of a velocity node: with m + nu k in front
  // of it) and the static record
  auto issue = [&](unsigned i, int c, int u) {
    if (lane == 0) {
      const unsigned slot_s = ring_s + u * (GSLOT * 8), bar_s = bars_s + u * 8;
      const double* src = stage + (size_t)i * ROWD;
      const bool with_dg = VEL && PRE && c == 0;
      const unsigned bytes = with_dg ? (BW + CSEG) * 8 : CSEG * 8;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slot's previous readers (generic proxy) come first
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes + META * 8) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(with_dg ? slot_s : slot_s + BW * 8),
                   "l"(with_dg || !VEL ? src : src + BW + c * CSEG), "r"(bytes), "r"(bar_s)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(slot_s + (BW + CSEG) * 8),
                   "l"(meta + (size_t)i * META), "r"(META * 8), "r"(bar_s)
                   : "memory");
    }
  };
  for (long long j_lo = begin + gw * g.wb; j_lo < end; j_lo += nw * g.wb) {
    const int n_it = (int)((j_lo + g.wb < end ? j_lo + g.wb : end) - j_lo);   // <= 32
    const unsigned h_iend = lane < n_it ? incptr[j_lo + lane + 1] : 0u;
    const unsigned i_lo = incptr[j_lo];
    // producer cursor: item pt, sweep pc, incidence pi of [pb, pe)
    int pt = 0, pc = 0;
    unsigned pb = i_lo, pe = __shfl_sync(FULLM, h_iend, 0), pi = i_lo;
    auto produce = [&](int u) {
      issue(pi, pc, u);
      if (++pi == pe) {
        pi = pb;
        if (++pc == NSWEEP) {
          pc = 0;
          pb = pe;
          pi = pb;
          if (++pt < n_it) pe = __shfl_sync(FULLM, h_iend, pt);
        }
      }
    };
    int u = 0;
    for (; u < RD && pt < n_it; ++u) produce(u);
    u = 0;
    long long rcur = rows(g0s[j_lo]), rnext = 0;
    bool item_start = true;
    int maskA = 7, c = 0, ct = 0;
    int par = (int)__shfl_sync(FULLM, rcur, 0) & 1;                   // 16-byte phase of the row of this sweep
    int parP = PRE ? (int)__shfl_sync(FULLM, rcur, 8) & 1 : 0;        // ... of the first preconditioner row
    while (ct < n_it) {
      mbar_wait(m.bars + u, (phases >> u) & 1u);
      phases ^= 1u << u;
      const double* S = m.ring + u * GSLOT;
      const double* seg = S + BW;
      const uint4 hw = *reinterpret_cast<const uint4*>(S + BW + CSEG);
      const unsigned short* spos = reinterpret_cast<const unsigned short*>(S + BW + CSEG) + MPOS;
      const unsigned short* sppos = reinterpret_cast<const unsigned short*>(S + BW + CSEG) + MPPOS;
      const unsigned e = hw.x, hi = hw.z;
      const int wp = (int)hw.y;
      const bool first = (hi & HI_FIRST) != 0, np = PRE && (hi & HI_NP) != 0;
      if (item_start) {
        item_start = false;
        const int g0n = (int)hw.w;
        if (g0n >= 0) rnext = rows(g0n);
        if (VEL && np && first) {
          // general path of the preconditioner: its rows are updated in place, so the first chunk clears them first
          const long long ps = __shfl_sync(FULLM, rcur, 8), pe2 = __shfl_sync(FULLM, rcur, 11);
          for (long long k = ps + lane; k < pe2; k += 32) vq[k] = 0.0;
          __syncwarp();
        }
      }
      if (VEL) {
        if (!(e & INC_CS)) {
          // no constrained dof in the cell (all but the boundary cells): every lane updates -- the idle ones a spare entry
          double* t = acc + (lane < NU ? par + spos[lane] : ACC0 - 3);
          const double t0 = t[0], t1 = t[1], t2 = t[2];
          const double s0 = seg[lane], s1 = seg[BW + lane], s2 = seg[2 * BW + lane];
          double* t01 = acc01 + (lane < NP ? (int)spos[NU + (lane & 7)] : L01 - 1);
          const double t3 = *t01, s3 = seg[3 * BW + (lane & 7)];
          t[0] = t0 + s0;
          t[1] = t1 + s1;
          t[2] = t2 + s2;
          *t01 = t3 + s3;
        } else {
          const int a = e & 31;
          const unsigned char* mrow = g.nmask + ((size_t)g.w_base + ((e & ~INC_CS) >> 5)) * MSTR;
          const int mb = lane < NU ? mrow[lane] : 0, mA = mrow[a];
          maskA = mA;
          if ((mA >> c) & 1) {
            if (lane < NU) {
              double* t = acc + par + spos[lane];
              int idx[3];
              double tv[3];
#pragma unroll
              for (int d = 0; d < 3; ++d) idx[d] = ((mb >> d) & 1) ? __popc(mb & ((1 << d) - 1)) : -1;
#pragma unroll
              for (int d = 0; d < 3; ++d) tv[d] = idx[d] >= 0 ? t[idx[d]] : 0.0;
#pragma unroll
              for (int d = 0; d < 3; ++d)
                if (idx[d] >= 0) t[idx[d]] = tv[d] + seg[d * BW + lane];
            }
            if (lane < NP) acc01[spos[NU + lane]] += seg[3 * BW + lane];
          } else if (lane == a)
            accd[0] += seg[c * BW + a];   // constrained dof: |L_ii| staged in the slot [c][c] of the node's own column
        }
        if (PRE && c == 0 && wp >= 0) {
          const double dg = S[lane & 31];   // lanes >= 27: the first entries of the segment, added to the spare entry
          if (!np) {
            const unsigned o = lane < NU ? (unsigned)sppos[lane] : 0xffffu;
            double* t = accP + (o != 0xffffu ? parP + (int)o : ptrash);
            *t += dg;
          } else {
            // cells with masks or per-component positions (boundary): read-modify-write of the rows in global memory
            const int a = e & 31;
            const int wide = (int)((hi & HI_WIDE) >> 1) - 1;
            int pmb = 7, pa = 7;
            if (hi & 1u) {
              const unsigned char* nm = g.pnmask + (size_t)wp * MSTR;
              pmb = lane < NU ? nm[lane] : 0;
              pa = nm[a];
            }
            __syncwarp();
#pragma unroll 1
            for (int c2 = 0; c2 < 3; ++c2) {
              double* row = vq + __shfl_sync(FULLM, rcur, 8 + c2);
              if (lane < NU) {
                unsigned o = sppos[lane];
                if (wide >= 0) o = g.ppos_wide[(size_t)wide * (3 * NU * NU) + c2 * (NU * NU) + a * NU + lane];
                if ((((pa & pmb) >> c2) & 1) && o != 0xffff) row[o] += dg;
                if (lane == a && !((pa >> c2) & 1)) row[0] += fabs(dg);
              }
              __syncwarp();
            }
          }
        }
      } else {
        if (lane < NU) {
          int mb = 7;
          if (e & INC_CS) mb = g.nmask[((size_t)g.w_base + ((e & ~INC_CS) >> 5)) * MSTR + lane];
          double* t = acc + par + spos[lane];
          int idx[3];
          double tv[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) idx[d] = ((mb >> d) & 1) ? __popc(mb & ((1 << d) - 1)) : -1;
#pragma unroll
          for (int d = 0; d < 3; ++d) tv[d] = idx[d] >= 0 ? t[idx[d]] : 0.0;
#pragma unroll
          for (int d = 0; d < 3; ++d)
            if (idx[d] >= 0) t[idx[d]] = tv[d] + seg[d * BW + lane];
        }
        if (PRE && wp >= 0 && lane < NP) {
          const unsigned o = sppos[lane];
          if (o != 0xffff) accP[o] += seg[3 * BW + lane];
        }
      }
      __syncwarp();   // every lane has read the slot and updated the accumulators
      if (pt < n_it) produce(u);
      u = u + 1 == RD ? 0 : u + 1;
      if (hi & HI_LAST) {   // the sweep is complete: write its rows
        if (VEL) {
          const long long rs = __shfl_sync(FULLM, rcur, c), rs01 = __shfl_sync(FULLM, rcur, 4 + c);
          const int len = (int)(__shfl_sync(FULLM, rcur, c + 1) - rs), len01 = (int)(__shfl_sync(FULLM, rcur, 5 + c) - rs01);
          const bool plain_pre = PRE && c == 0 && !np;
          long long p0 = 0, p1 = 0, p2 = 0, p3 = 0;
          int nq = 0;
          if ((maskA >> c) & 1) {
            bulk_row(va + rs, acc, acc_s, par, len, first, lane);
            flush_row<true>(vb + rs01, acc01, len01, L01, first, lane);
          } else if (lane == 0) {
            // constrained dof: the row holds its diagonal only
            double* out = va + rs;
            if (first) {
              out[0] = accd[0];
              for (int k = 1; k < len; ++k) out[k] = 0.0;
              for (int k = 0; k < len01; ++k) vb[rs01 + k] = 0.0;
            } else
              out[0] = __ldcg(out) + accd[0];
            accd[0] = 0.0;
          }
          if (plain_pre) {
            // the three component rows of the node are equal: accumulated once, stored three times -- as bulk copies
            // where the row's 16-byte phase is the accumulator's, through the lanes otherwise
            p0 = __shfl_sync(FULLM, rcur, 8), p1 = __shfl_sync(FULLM, rcur, 9), p2 = __shfl_sync(FULLM, rcur, 10), p3 = __shfl_sync(FULLM, rcur, 11);
            const int cap = g.pstr - 2;
            const long long ps[4] = {p0, p1, p2, p3};
#pragma unroll
            for (int c2 = 0; c2 < 3; ++c2) {
              const int plen = (int)(ps[c2 + 1] - ps[c2]), n = plen < cap ? plen : cap;
              if (n > nq) nq = n;
              if (((int)ps[c2] & 1) == parP)
                bulk_row(vq + ps[c2], accP, accP_s, parP, n, first, lane);
              else
                flush_row<false>(vq + ps[c2], accP + parP, n, n, first, lane);
              if (first)
                for (int k = cap + lane; k < plen; k += 32) __stcs(vq + ps[c2] + k, 0.0);
            }
          }
          bulk_wait_read(lane);   // the bulk copies have read the accumulators (also orders the lanes' own reads)
          if ((maskA >> c) & 1) clear_acc(acc, len + par, lane);
          if (plain_pre) clear_acc(accP, nq + parP, lane);
        } else {
          const long long rs = __shfl_sync(FULLM, rcur, 0);
          const int len = (int)(__shfl_sync(FULLM, rcur, 1) - rs);
          bulk_row(va + rs, acc, acc_s, par, len, first, lane);
          if (PRE) {
            const long long ps = __shfl_sync(FULLM, rcur, 8);
            flush_row<true>(vq + ps, accP, (int)(__shfl_sync(FULLM, rcur, 9) - ps), g.pstr - 2, first, lane);
          }
          bulk_wait_read(lane);
          clear_acc(acc, len + par, lane);
        }
        __syncwarp();
        maskA = 7;
        if (++c == NSWEEP) {   // next item
          c = 0;
          ++ct;
          rcur = rnext;
          item_start = true;
          parP = PRE ? (int)__shfl_sync(FULLM, rcur, 8) & 1 : 0;
        }
        par = (int)__shfl_sync(FULLM, rcur, c) & 1;
      }
    }
  }
}

__host__ __device__ constexpr int gather_warp_doubles(int pstr) { return ACC0 + L01 + pstr + 4 + RD * GSLOT + RD + (RD & 1); }

template <bool PRE>
__global__ void __launch_bounds__(GW * 32) th_gather_kernel(GatherArgs g, BlockView A, BlockView Ap) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nacc = ACC0 + L01 + g.pstr + 4;
  WarpMem m;
  m.acc = smem + warp * gather_warp_doubles(g.pstr);
  m.acc01 = m.acc + ACC0;
  m.accP = m.acc01 + L01;
  m.accd = m.accP + g.pstr;
  m.ring = m.acc + nacc;
  m.bars = reinterpret_cast<unsigned long long*>(m.ring + RD * GSLOT);
  for (int k = lane; k < nacc; k += 32) m.acc[k] = 0.0;   // invariant: all accumulators are zero between sweeps
  if (lane == 0) {
    for (int u = 0; u < RD; ++u) mbar_init(m.bars + u, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  unsigned phases = 0;   // parity of every slot's next completion
  const long long gw = (long long)blockIdx.x * GW + warp, nw = (long long)gridDim.x * GW;
  gather_items<true, PRE>(g, A, Ap, m, phases, gw, nw, lane);
  gather_items<false, PRE>(g, A, Ap, m, phases, gw, nw, lane);
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the rows are in global memory
}

// The static records of all incidences, in incidence order: one warp per item.
struct MetaArgs {
  long long n_items;
  const int* g0s;
  const unsigned* incptr;
  const unsigned* incs;
  const unsigned char* flags;
  const long long* pre_w;          // nullptr: no fused preconditioner
  const unsigned short* pos;
  const unsigned short* ppos;
  long long chunk;                 // cells per chunk
  double* meta;
};
template <bool VEL>
__global__ void gather_meta_kernel(MetaArgs a) {
  const int lane = threadIdx.x & 31;
  const long long j = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= a.n_items) return;
  constexpr long long NL = VEL ? NU : NP;
  const unsigned ib = a.incptr[j], ie = a.incptr[j + 1];
  const long long ch = (long long)ib / (NL * a.chunk), w_base = ch * a.chunk;
  int g0n = -1;
  if (j + 1 < a.n_items && (long long)a.incptr[j + 1] / (NL * a.chunk) == ch) g0n = a.g0s[j + 1];
  unsigned e = 0;
  long long ax = -1;
  if (ib + lane < ie) {
    e = a.incs[ib + lane];
    if (a.pre_w) ax = a.pre_w[w_base + ((e & ~INC_CS) >> 5)];
  }
  const bool np_item = __ballot_sync(FULLM, (int)ax >= 0 && (int)(ax >> 32) != 0) != 0;
  const unsigned item_bits = (VEL && np_item ? HI_NP : 0u) | ((a.flags[j] & 1) ? HI_FIRST : 0u);
  for (unsigned i = ib; i < ie; ++i) {
    const unsigned ei = __shfl_sync(FULLM, e, (int)(i - ib));
    const long long axi = __shfl_sync(FULLM, ax, (int)(i - ib));
    const int wp = (int)axi, hi = (int)(axi >> 32);
    const size_t w = (size_t)w_base + ((ei & ~INC_CS) >> 5);
    const int ln = ei & 31;
    double* rec = a.meta + (size_t)i * META;
    if (lane == 0) {
      IncHdr h;
      h.e = ei;
      h.wp = wp;
      h.hi = (wp >= 0 ? ((unsigned)hi & (HI_WIDE | 1u)) : 0u) | item_bits | (i + 1 == ie ? HI_LAST : 0u);
      h.g0n = g0n;
      *reinterpret_cast<IncHdr*>(rec) = h;
    }
    unsigned short* r16 = reinterpret_cast<unsigned short*>(rec);
    const unsigned short* prow = a.pos + w * PSTR + (VEL ? ln : NU + ln) * NE;
    for (int t = lane; t < 36; t += 32) r16[MPOS + t] = t < (VEL ? NE : NU) ? prow[t] : (unsigned short)0xffff;
    if (lane < 28) {
      unsigned short v = 0xffff;
      if (wp >= 0) {
        const unsigned short* pprow = a.ppos + (size_t)wp * PSTR + (VEL ? ln : NU + ln) * NE;
        if (VEL ? lane < NU : lane < NP) v = VEL ? pprow[lane] : pprow[NU + lane];
      }
      r16[MPPOS + lane] = v;
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
  cudaFree(p->slots);
  cudaFree(p->v_meta);
  cudaFree(p->p_meta);
  cudaFree(p->staging);
  cudaFree(p->dphi_lane);
  cudaFree(p->pre_w);
  cudaFree(p->pre_rest);
  delete p;
}

// Items (chunk, node) with their incidences (cell of the chunk, local node) for the gather pass, and the staging slot
// of every (cell, local node).  `cells` is the plan order of the masked plan.  Returns DCP_OK with *out == nullptr when
// the model does not qualify (row longer than the accumulators, more than 8 cells of a chunk at one node): the caller
// keeps the reduction path.
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
  if (max_len(pat[0][0]) > LROW || max_len(pat[1][0]) > LROW || max_len(pat[0][1]) > L01 - 1) return DCP_OK;   // (the last entry of acc01 is the spare one)
  // one stage + gather launch pair per chunk of DCP_GATHER_CHUNK cells (default 65 536; 131 072 was measured equal at refine 6 and costs 4.7 GB more)
  int64_t chunk = 65536;
  if (const char* e = std::getenv("DCP_GATHER_CHUNK")) chunk = std::max<int64_t>(1, std::atoll(e));
  chunk = std::min<int64_t>(chunk, n);
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
  std::vector<uint32_t> slots((size_t)n * SLOTS, 0u);
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
            const uint32_t word = (uint32_t)(keys[e] & 0xffffffffu);
            I.inc.push_back(word);
            // the staged row of this (cell, node) goes to the slot of its incidence: the rows of one node are consecutive
            slots[(size_t)(w0 + ((word & ~INC_CS) >> 5)) * SLOTS + (pass == 0 ? 0 : NU) + (word & 31)] = (uint32_t)e;
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
  if (std::getenv("DCP_VERBOSE"))
    std::fprintf(stderr, "[dcp] staged assembly: %lld cells in %lld chunks of %lld\n", (long long)n, (long long)n_chunks, (long long)chunk);
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
  if (rc == DCP_OK) rc = upg(ctx, &G->slots, slots);
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
  if (rc == DCP_OK && cudaMalloc((void**)&G->staging, sizeof(double) * (size_t)REC * (size_t)chunk) != cudaSuccess) {
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

// The static record stream of the gather (header + positions per incidence, in incidence order), built on the device
// from the item lists and the two plans; the item lists' incidence words and flags are not needed afterwards.
static int gather_meta_build(dcp_model* m, GatherPlan* G, const MaskedPlan* nse_plan, const MaskedPlan* pre_plan) {
  dcp_ctx* ctx = m->ctx;
  for (int pass = 0; pass < 2; ++pass) {
    const int64_t n_items = pass == 0 ? G->v_chunk_ptr.back() : G->p_chunk_ptr.back();
    const int64_t n_inc = (int64_t)(pass == 0 ? NU : NP) * G->n_cells;
    double** dst = pass == 0 ? &G->v_meta : &G->p_meta;
    if (cudaMalloc((void**)dst, sizeof(double) * META * (size_t)n_inc) != cudaSuccess) {
      cudaGetLastError();
      dcp_set_error("gather plan: allocation of the record stream failed");
      return DCP_ERR_CUDA;
    }
    MetaArgs a;
    a.n_items = n_items;
    a.g0s = pass == 0 ? G->v_g0 : G->p_g0;
    a.incptr = pass == 0 ? G->v_incptr : G->p_incptr;
    a.incs = pass == 0 ? G->v_inc : G->p_inc;
    a.flags = pass == 0 ? G->v_flag : G->p_flag;
    a.pre_w = G->has_pre ? G->pre_w : nullptr;
    a.pos = nse_plan->pos;
    a.ppos = G->has_pre ? pre_plan->pos : nullptr;
    a.chunk = G->chunk;
    a.meta = *dst;
    const unsigned grid = (unsigned)((n_items + 7) / 8);
    if (pass == 0)
      gather_meta_kernel<true><<<grid, 256, 0, ctx->stream>>>(a);
    else
      gather_meta_kernel<false><<<grid, 256, 0, ctx->stream>>>(a);
  }
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  DCP_CUDA(cudaGetLastError());
  cudaFree(G->v_inc);
  cudaFree(G->p_inc);
  cudaFree(G->v_flag);
  cudaFree(G->p_flag);
  cudaFree(G->pre_w);
  G->v_inc = G->p_inc = nullptr;
  G->v_flag = G->p_flag = nullptr;
  G->pre_w = nullptr;
  return DCP_OK;
}

// Finish the gather plan: map the system plan's cells to the preconditioner plan (fused preconditioner; not attached --
// the preconditioner keeps its own pass -- when a row of that matrix that the gather has to accumulate is longer than the
// accumulators), then build the record stream.
int dcp_gather_plan_attach_pre(dcp_model* m, const dcp_model_desc* d, GatherPlan* G, const MaskedPlan* nse_plan, const MaskedPlan* pre_plan) {
  if (!G || !nse_plan) return DCP_OK;
  bool fuse = pre_plan != nullptr && !std::getenv("DCP_NO_FUSED_PRECONDITIONER");
  if (fuse) {
    // accumulator of one preconditioner row: the largest offset a gathered cell writes to (rows of nodes with
    // no-normal-flux lines are longer, but their cells are not gathered), and the longest pressure-mass row
    int64_t need = (int64_t)pre_plan->max_off_plain + 1;
    const dcp_csr_desc& P11 = d->pre_pattern[1][1];
    int64_t mx = 0;
#pragma omp parallel for reduction(max : mx)
    for (int64_t r = 0; r < P11.n_rows; ++r) mx = std::max(mx, P11.rowptr[r + 1] - P11.rowptr[r]);
    need = std::max(need, mx);
    if (need > LROW) fuse = false;
    else G->pstr = (int)((need + 2 + 3) / 4 * 4);   // + the parity shift and the spare entry
  }
  if (fuse) {
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
  }
  return gather_meta_build(m, G, nse_plan, pre_plan);
}

// NSE system (and, fused, its preconditioner), write-once: per chunk of plan cells, stage then gather (stream-ordered).
int dcp_launch_th_staged(dcp_model* m, const dcp_params& p, const MaskedPlan* plan, const double* old_nse, const double* old_temp) {
  dcp_ctx* ctx = m->ctx;
  const GatherPlan* G = plan->gather;
  MmaArgs a;
  a.n_fast = plan->n;
  a.wlist = nullptr;
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
  const bool fuse_pre = G->has_pre && m->masked_pre;
  const int pstr = fuse_pre ? G->pstr : 4;
  const size_t smem_s = stage_smem_bytes(), smem_g = sizeof(double) * (size_t)gather_warp_doubles(pstr) * GW;
  auto gather_fn = fuse_pre ? th_gather_kernel<true> : th_gather_kernel<false>;
  DCP_CUDA(cudaFuncSetAttribute(th_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
  DCP_CUDA(cudaFuncSetAttribute(gather_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
  int per_sm_s = 1, per_sm_g = 1;
  DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_s, th_stage_kernel, MTHREADS, smem_s));
  DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_g, gather_fn, GW * 32, smem_g));
  per_sm_s = std::max(per_sm_s, 1);
  per_sm_g = std::max(per_sm_g, 1);
  if (const char* e = std::getenv("DCP_GATHER_CTAS")) per_sm_g = std::max(1, std::min(per_sm_g, std::atoi(e)));
  GatherArgs g{};
  g.v_g0 = G->v_g0;
  g.v_incptr = G->v_incptr;
  g.v_meta = G->v_meta;
  g.p_g0 = G->p_g0;
  g.p_incptr = G->p_incptr;
  g.p_meta = G->p_meta;
  g.nmask = plan->nmask;
  g.wb = 8;
  if (const char* e = std::getenv("DCP_GATHER_BLOCK")) g.wb = std::min(32, std::max(1, std::atoi(e)));   // <= 32: one item per lane in a block
  g.vstage = G->staging;
  g.pstage = G->staging + (size_t)NU * VROW * (size_t)G->chunk;
  g.pstr = pstr;
  if (fuse_pre) {
    g.pnmask = m->masked_pre->nmask;
    g.ppos_wide = m->masked_pre->pos_wide;
  }
  StageOut so;
  so.vstage = G->staging;
  so.pstage = G->staging + (size_t)NU * VROW * (size_t)G->chunk;
  so.slots = G->slots;
  const BlockView A = make_view(m->nse);
  const BlockView Apre = make_view(m->pre);
  const CsView cs = make_view(m->nse_cs);
  const bool debug_sync = std::getenv("DCP_DEBUG_SYNC") != nullptr;   // name the kernel that faults
  auto checkpoint = [&](const char* what, long long ch) {
    if (!debug_sync) return DCP_OK;
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) return DCP_OK;
    dcp_set_error(std::string("staged assembly, ") + what + " of chunk " + std::to_string(ch) + ": " + cudaGetErrorString(e));
    return DCP_ERR_CUDA;
  };
  for (int64_t ch = 0; ch < G->n_chunks; ++ch) {
    const long long w0 = ch * G->chunk, w1 = std::min<long long>(plan->n, w0 + G->chunk);
    long long grid = std::min<long long>((long long)ctx->sm_count * per_sm_s, w1 - w0);
    th_stage_kernel<<<(unsigned)grid, MTHREADS, smem_s, ctx->stream>>>(a, cs, G->dphi_lane, so, w0, w1);
    DCP_TRY(checkpoint("stage kernel", ch));
    g.v_begin = G->v_chunk_ptr[ch];
    g.v_end = G->v_chunk_ptr[ch + 1];
    g.p_begin = G->p_chunk_ptr[ch];
    g.p_end = G->p_chunk_ptr[ch + 1];
    g.w_base = w0;
    const long long items = std::max(g.v_end - g.v_begin, g.p_end - g.p_begin);
    const long long blocks = (items + g.wb - 1) / g.wb;   // a warp takes blocks of g.wb items
    grid = std::min<long long>((long long)ctx->sm_count * per_sm_g, (blocks + GW - 1) / GW);
    if (grid > 0) gather_fn<<<(unsigned)grid, GW * 32, smem_g, ctx->stream>>>(g, A, Apre);
    DCP_TRY(checkpoint("gather kernel", ch));
    ctx->launches += 2;
  }
  if (fuse_pre) {
    // the cells the gather skipped: no-normal-flux cells through the reduction kernel, cells outside the
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
