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

constexpr int VROW = 268;                   // staged velocity-node row: [9][27] + pad + [3][8]
constexpr int VPRS = 244;                   // start of the pressure columns inside a velocity-node row
constexpr int PROW = 82;                    // staged pressure-node row: [3][27] + pad
constexpr int REC = NU * VROW + NP * PROW;  // doubles per cell record
constexpr int TBUF = 8 * 74;                // per-warp transposition buffer (direct: stride 72, transposed: stride 74)
constexpr int LROW = 384;                   // longest block(0,0) / block(1,0) row the gather accumulators hold
constexpr int ASTR = 387;                   // accumulator row stride (387 mod 16 == 3: the three components of one column land in different banks)
constexpr int L01 = 32;                     // longest block(0,1) row
constexpr int GACC = 3 * ASTR + 3 * L01 + 7;  // per-warp accumulators (+3 diagonal slots of constrained components), 1264
constexpr int GWARPS = 8;
static_assert(GACC % 2 == 0, "accumulator alignment");
static_assert((REC * 8) % 16 == 0, "record alignment");

// ---- stage: contraction + constraint epilogue, coalesced write of the cell record ---------------------------------
__global__ void __launch_bounds__(MTHREADS, 3)
th_stage_kernel(MmaArgs a, CsView cs, double* __restrict__ stage, long long w_begin, long long w_end, long long ring) {
  extern __shared__ __align__(16) double smem[];
  double* X = smem;                     // KQ * LDB
  double* wq = X + KQ * LDB;            // KQ (+4)
  double* sgeo2 = wq + 32;              // GS
  double* swt = sgeo2 + GS + 1;         // 3*NU
  double* sF = swt + 3 * NU + 1;        // NQ*3
  double* sU = sF + NQ * 3;             // ND
  double* sT = sU + ND + 1;             // 32
  double* sTn = sT + 32;                // 28
  double* sGU = sTn + 28;               // NQ*12
  double* tbuf_all = sGU + NQ * 12 + 1; // 4 * TBUF   (offset 5166: 16-byte aligned)
  unsigned char* snm2 = (unsigned char*)(tbuf_all + 4 * TBUF);  // MSTR
  int* sidx2 = (int*)(snm2 + MSTR);                            // IDS
  int* sidt2 = sidx2 + IDS;                                    // 28
  int* sys_u = sidt2 + 28;                                     // 3*NU
  int* sys_p = sys_u + 3 * NU;                                 // NP
  unsigned char* skc = (unsigned char*)(sys_p + NP);           // 28
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  double* tb = tbuf_all + warp * TBUF;

  for (int i = tid; i < ND; i += nt) {
    const int f = a.local_field[i], bs = a.local_base[i];
    if (f < 3) sys_u[f * NU + bs] = i; else sys_p[bs] = i;
  }
  for (int i = tid; i < KQ * LDB; i += nt) X[i] = 0.0;
  if (tid < 4) wq[NQ + tid] = 0.0;
  const double nu = a.prm.dt * a.prm.inv_re;
  const bool do_rhs = a.rhs != nullptr;
  const double* sgeo = sgeo2;
  const unsigned char* snm = snm2;
  const int* sidx = sidx2;
  const int* sidt = sidt2;
  auto node_cs = [&](int n) {
    NodeCs c;
    c.mask = snm[n];
    c.k = skc[n];
    c.w0 = swt[3 * n];
    c.w1 = swt[3 * n + 1];
    c.w2 = swt[3 * n + 2];
    return c;
  };

  for (long long w = w_begin + blockIdx.x; w < w_end; w += gridDim.x) {
    const long long cell = a.cells[w];
    double* S = stage + (size_t)(w % ring) * REC;
    __syncthreads();
    {
      const double* g = a.geom + cell * GS;
      for (int i = tid; i < GS; i += nt) cp_async8(sgeo2 + i, g + i);
      if (tid < MSTR / 8) cp_async8(snm2 + 8 * tid, a.nmask + w * MSTR + 8 * tid);
      for (int i = tid; i < ND; i += nt) cp_async4(sidx2 + i, a.l2g + cell * ND + i);
      if (do_rhs)
        for (int i = tid; i < a.ndt; i += nt) cp_async4(sidt2 + i, a.l2g_t + cell * a.ndt + i);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (do_rhs) {
      for (int i = tid; i < ND; i += nt) cp_async8(sU + i, a.old_nse + sidx[i]);
      for (int i = tid; i < a.ndt; i += nt) cp_async8(sTn + i, a.old_temp + sidt[i]);
    }
    cp_async_commit();
    const int cflag = snm[35];
    if (tid < NU) {
      int kc = 3;
      double w0 = 0.0, w1 = 0.0, w2 = 0.0;
      if (cflag && snm[tid] != 7) {
        const int g0 = sidx[sys_u[tid]];
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
      skc[tid] = (unsigned char)kc;
      swt[tid * 3] = w0;
      swt[tid * 3 + 1] = w1;
      swt[tid * 3 + 2] = w2;
    }
    if (tid < NQ) wq[tid] = sgeo[tid];
    for (int i = tid; i < NQ * NU; i += nt) {
      const int q = i / NU, b = i - q * NU;
      const double r0 = __ldg(a.dphi_u + i * 3), r1 = __ldg(a.dphi_u + i * 3 + 1), r2 = __ldg(a.dphi_u + i * 3 + 2);
      double* x = X + q * LDB + b;
#pragma unroll
      for (int d = 0; d < 3; ++d)
        x[32 * d] = sgeo[NQ * (1 + d) + q] * r0 + sgeo[NQ * (4 + d) + q] * r1 + sgeo[NQ * (7 + d) + q] * r2;
      x[96] = __ldg(a.phi_u + i);
    }
    for (int i = tid; i < NQ * NP; i += nt) X[(i / NP) * LDB + PSI0 + (i % NP)] = __ldg(a.phi_p + i);
    cp_async_wait<0>();
    __syncthreads();

    // ---- tasks: 10 node x node blocks (ta <= tb), 4 node x psi blocks; weights 10 : 3, so the four warps take
    // {0,1,2}, {3,4,5}, {6,7,10,11}, {8,9,12,13}
    const int frow = lane >> 2, fk = lane & 3;
    for (int s = 0; s < 4; ++s) {
      int t;
      if (warp < 2) {
        if (s == 3) break;
        t = 3 * warp + s;
      } else
        t = s < 2 ? (warp == 2 ? 6 : 8) + s : (warp == 2 ? 10 : 12) + (s - 2);
      if (t < 10) {
        int ta = 0, r = t;
        while (r >= 4 - ta) { r -= 4 - ta; ++ta; }
        const int tb_ = ta + r;
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
        double Fj[2][9];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int nb = 8 * tb_ + 2 * fk + jj;
#pragma unroll
          for (int r9 = 0; r9 < 9; ++r9) Fj[jj][r9] = 0.0;
          if (na >= NU || nb >= NU) continue;
          const NodeCs ca = node_cs(na), cb = node_cs(nb);
          const double dg = acc[3][3][jj] + nu * (acc[0][0][jj] + acc[1][1][jj] + acc[2][2][jj]);
          double F[3][3];
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int d = 0; d < 3; ++d) F[c][d] = nu * acc[d][c][jj] + (c == d ? dg : 0.0);
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
        // direct orientation: rows 8ta.., columns 8tb..; buffer [a][r][b], conflict-free 16-byte stores
#pragma unroll
        for (int r9 = 0; r9 < 9; ++r9)
          *reinterpret_cast<double2*>(tb + (frow * 9 + r9) * 8 + 2 * fk) = make_double2(Fj[0][r9], Fj[1][r9]);
        __syncwarp();
#pragma unroll 6
        for (int k = 0; k < 18; ++k) {
          const int seg = 4 * k + (lane >> 3), e = lane & 7;
          const int al = seg / 9, r9 = seg - 9 * al;
          const int na2 = 8 * ta + al, nb2 = 8 * tb_ + e;
          if (na2 < NU && nb2 < NU) __stcg(S + na2 * VROW + r9 * NU + nb2, tb[seg * 8 + e]);
        }
        __syncwarp();
        if (ta != tb_) {
          // transposed orientation: row node b, column node a, entry [d][c] = F[c][d]; buffer [b] stride 74, [r][a]
#pragma unroll
          for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int d = 0; d < 3; ++d) tb[(2 * fk + jj) * 74 + (d * 3 + c) * 8 + frow] = Fj[jj][c * 3 + d];
          __syncwarp();
#pragma unroll 6
          for (int k = 0; k < 18; ++k) {
            const int seg = 4 * k + (lane >> 3), e = lane & 7;
            const int bl = seg / 9, r9 = seg - 9 * bl;
            const int nb2 = 8 * tb_ + bl, na2 = 8 * ta + e;
            if (na2 < NU && nb2 < NU) __stcg(S + nb2 * VROW + r9 * NU + na2, tb[bl * 74 + r9 * 8 + e]);
          }
          __syncwarp();
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
          const NodeCs ca = node_cs(na);
          double sv[2][3];
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            sv[jj][0] = -acc[0][jj];
            sv[jj][1] = -acc[1][jj];
            sv[jj][2] = -acc[2][jj];
            if (ca.k != 3) {
              const double sk = ca.k == 0 ? sv[jj][0] : (ca.k == 1 ? sv[jj][1] : sv[jj][2]);
              sv[jj][0] += ca.w0 * sk;
              sv[jj][1] += ca.w1 * sk;
              sv[jj][2] += ca.w2 * sk;
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            // velocity-node row: [c][p]; pressure-node rows: [c][a]
            *reinterpret_cast<double2*>(S + na * VROW + VPRS + c * 8 + 2 * fk) = make_double2(sv[0][c], sv[1][c]);
            __stcg(S + NU * VROW + (2 * fk) * PROW + c * NU + na, sv[0][c]);
            __stcg(S + NU * VROW + (2 * fk + 1) * PROW + c * NU + na, sv[1][c]);
          }
        }
      }
    }
    if (do_rhs) {
      for (int q = tid; q < NQ; q += nt) {
        double tq = 0.0;
        for (int k = 0; k < a.ndt; ++k) tq += sTn[k] * __ldg(a.phi_t + q * a.ndt + k);
        sT[q] = tq;
      }
      __syncthreads();
      for (int i = tid; i < NQ * 12; i += nt) {
        const int q = i / 12, r = i - q * 12, c = r >> 2, e = r & 3;
        const double* x = X + q * LDB + 32 * e;
        double sacc = 0.0;
        for (int n = 0; n < NU; ++n) sacc += sU[sys_u[c * NU + n]] * x[n];
        sGU[i] = sacc;
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

constexpr size_t stage_smem_bytes() {
  return sizeof(double) * (KQ * LDB + 32 + GS + 1 + 3 * NU + 1 + NQ * 3 + ND + 1 + 32 + 28 + NQ * 12 + 1 + 4 * TBUF) + MSTR +
         sizeof(int) * (IDS + 28 + 3 * NU + NP) + 28 + 36;
}
static_assert((KQ * LDB + 32 + GS + 1 + 3 * NU + 1 + NQ * 3 + ND + 1 + 32 + 28 + NQ * 12 + 1) % 2 == 0, "tbuf alignment");

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
  long long ring;
  const unsigned short* pos;
  const unsigned char* nmask;
  const double* stage;
};

__global__ void __launch_bounds__(GWARPS * 32, 2) th_gather_kernel(GatherArgs g, BlockView A) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* acc = smem + warp * GACC;         // [3][ASTR]
  double* acc01 = acc + 3 * ASTR;           // [3][L01]
  double* accd = acc01 + 3 * L01;           // [3] constrained diagonals
  const long long gw = (long long)blockIdx.x * GWARPS + warp, nw = (long long)gridDim.x * GWARPS;
  const long long* rp00 = A.rowptr[0][0];
  const long long* rp01 = A.rowptr[0][1];
  const long long* rp10 = A.rowptr[1][0];
  double* v00 = A.val[0][0];
  double* v01 = A.val[0][1];
  double* v10 = A.val[1][0];

  // velocity nodes: rows (g0 + c) of block(0,0) and block(0,1)
  for (long long j = g.v_begin + gw; j < g.v_end; j += nw) {
    const int g0 = g.v_g0[j];
    const unsigned i0 = g.v_incptr[j], i1 = g.v_incptr[j + 1];
    const bool first = g.v_flag[j] & 1;
    long long ra = 0, rb = 0;
    if (lane < 4) {
      ra = rp00[g0 + lane];
      rb = rp01[g0 + lane];
    }
    long long rs[3], rs01[3];
    int len[3], len01[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      rs[c] = __shfl_sync(0xffffffffu, ra, c);
      len[c] = (int)(__shfl_sync(0xffffffffu, ra, c + 1) - rs[c]);
      rs01[c] = __shfl_sync(0xffffffffu, rb, c);
      len01[c] = (int)(__shfl_sync(0xffffffffu, rb, c + 1) - rs01[c]);
    }
    for (int k = lane; k < GACC; k += 32) acc[k] = 0.0;
    __syncwarp();
    int maskA = 7;
    for (unsigned i = i0; i < i1; ++i) {
      const unsigned e = g.v_inc[i];
      const int a = e & 31;
      const long long w = g.w_base + (e >> 5);
      const unsigned short* prow = g.pos + w * PSTR + a * NE;
      const unsigned char* mrow = g.nmask + w * MSTR;
      const int ob = lane < NU ? prow[lane] : 0;
      const int op = lane < NP ? prow[NU + lane] : 0;
      const int mb = lane < NU ? mrow[lane] : 0;
      maskA = mrow[a];
      const double* S = g.stage + (size_t)(w % g.ring) * REC + a * VROW;
      if (lane < NU) {
        double v[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) v[r] = __ldcg(S + r * NU + lane);
#pragma unroll
        for (int r = 0; r < 9; ++r) {
          const int c = r / 3, d = r - 3 * c;
          if (((maskA >> c) & 1) && ((mb >> d) & 1)) acc[c * ASTR + ob + __popc(mb & ((1 << d) - 1))] += v[r];
        }
        if (lane == a && maskA != 7) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (!((maskA >> c) & 1)) accd[c] += v[4 * c];
        }
      }
      {
        const int c = lane >> 3, pb = lane & 7;
        const int o = __shfl_sync(0xffffffffu, op, pb);
        if (lane < 24 && ((maskA >> c) & 1)) acc01[c * L01 + o] += __ldcg(S + VPRS + lane);
      }
      __syncwarp();
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if ((maskA >> c) & 1) {
        double* out = v00 + rs[c];
        for (int k = lane; k < len[c]; k += 32) {
          const double v = acc[c * ASTR + k];
          if (first) out[k] = v; else if (v != 0.0) out[k] = __ldcg(out + k) + v;
        }
        double* out1 = v01 + rs01[c];
        for (int k = lane; k < len01[c]; k += 32) {
          const double v = acc01[c * L01 + k];
          if (first) out1[k] = v; else if (v != 0.0) out1[k] = __ldcg(out1 + k) + v;
        }
      } else if (lane == 0) {
        // constrained dof: the row holds its diagonal only
        double* out = v00 + rs[c];
        if (first) {
          out[0] = accd[c];
          for (int k = 1; k < len[c]; ++k) out[k] = 0.0;
          for (int k = 0; k < len01[c]; ++k) v01[rs01[c] + k] = 0.0;
        } else
          out[0] = __ldcg(out) + accd[c];
      }
    }
    __syncwarp();
  }

  // pressure nodes: row of block(1,0)
  for (long long j = g.p_begin + gw; j < g.p_end; j += nw) {
    const int pr = g.p_g0[j];
    const unsigned i0 = g.p_incptr[j], i1 = g.p_incptr[j + 1];
    const bool first = g.p_flag[j] & 1;
    const long long rs = rp10[pr];
    const int len = (int)(rp10[pr + 1] - rs);
    for (int k = lane; k < ASTR; k += 32) acc[k] = 0.0;
    __syncwarp();
    for (unsigned i = i0; i < i1; ++i) {
      const unsigned e = g.p_inc[i];
      const int pa = e & 31;
      const long long w = g.w_base + (e >> 5);
      const unsigned short* prow = g.pos + w * PSTR + (NU + pa) * NE;
      if (lane < NU) {
        const int ob = prow[lane];
        const int mb = g.nmask[w * MSTR + lane];
        const double* S = g.stage + (size_t)(w % g.ring) * REC + NU * VROW + pa * PROW;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double v = __ldcg(S + c * NU + lane);
          if ((mb >> c) & 1) acc[ob + __popc(mb & ((1 << c) - 1))] += v;
        }
      }
      __syncwarp();
    }
    double* out = v10 + rs;
    for (int k = lane; k < len; k += 32) {
      const double v = acc[k];
      if (first) out[k] = v; else if (v != 0.0) out[k] = __ldcg(out + k) + v;
    }
    __syncwarp();
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
  delete p;
}

// Items (chunk, node) with their incidences (cell of the chunk, local node) for the gather pass.  `cells` is the plan
// order of the masked plan.  Returns DCP_OK with *out == nullptr when the model does not qualify (row longer than the
// accumulators): the caller keeps the reduction path.
int dcp_gather_plan_build(dcp_model* m, const dcp_model_desc* d, const std::vector<int32_t>& cells, GatherPlan** out) {
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
            keys.push_back((g0 << 32) | ((uint64_t)(w - w0) << 5) | (uint64_t)a);
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
          I.flag.push_back(fst[g0] == (int32_t)ch ? 1 : 0);
          k = e;
        }
      }
    }
  }
  GatherPlan* G = new GatherPlan;
  G->chunk = chunk;
  G->n_chunks = n_chunks;
  G->n_cells = n;
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
  const size_t smem_s = stage_smem_bytes(), smem_g = sizeof(double) * GACC * GWARPS;
  DCP_CUDA(cudaFuncSetAttribute(th_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
  DCP_CUDA(cudaFuncSetAttribute(th_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
  int per_sm_s = 1, per_sm_g = 1;
  DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_s, th_stage_kernel, MTHREADS, smem_s));
  DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_g, th_gather_kernel, GWARPS * 32, smem_g));
  per_sm_s = std::max(per_sm_s, 1);
  per_sm_g = std::max(per_sm_g, 1);
  GatherArgs g;
  g.v_g0 = G->v_g0;
  g.v_incptr = G->v_incptr;
  g.v_flag = G->v_flag;
  g.v_inc = G->v_inc;
  g.p_g0 = G->p_g0;
  g.p_incptr = G->p_incptr;
  g.p_flag = G->p_flag;
  g.p_inc = G->p_inc;
  g.ring = G->chunk;
  g.pos = plan->pos;
  g.nmask = plan->nmask;
  g.stage = G->staging;
  const BlockView A = make_view(m->nse);
  const CsView cs = make_view(m->nse_cs);
  for (int64_t ch = 0; ch < G->n_chunks; ++ch) {
    const long long w0 = ch * G->chunk, w1 = std::min<long long>(plan->n, w0 + G->chunk);
    long long grid = std::min<long long>((long long)ctx->sm_count * per_sm_s, w1 - w0);
    th_stage_kernel<<<(unsigned)grid, MTHREADS, smem_s, ctx->stream>>>(a, cs, G->staging, w0, w1, G->chunk);
    g.v_begin = G->v_chunk_ptr[ch];
    g.v_end = G->v_chunk_ptr[ch + 1];
    g.p_begin = G->p_chunk_ptr[ch];
    g.p_end = G->p_chunk_ptr[ch + 1];
    g.w_base = w0;
    const long long items = (g.v_end - g.v_begin) + (g.p_end - g.p_begin);
    grid = std::min<long long>((long long)ctx->sm_count * per_sm_g, (items + GWARPS - 1) / GWARPS);
    if (grid > 0) th_gather_kernel<<<(unsigned)grid, GWARPS * 32, smem_g, ctx->stream>>>(g, A);
    ctx->launches += 2;
  }
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}
