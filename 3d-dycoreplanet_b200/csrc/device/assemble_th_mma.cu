// Classic 3-D Taylor-Hood NSE system / preconditioner: the per-cell contraction on the FP64 tensor cores
// (mma.sync.m8n8k4.f64, SASS DMMA), constraints resolved in the epilogue, scatter through a per-cell plan.
//
// Same integrals as assemble_th.cu (reference: include/core/boussinesq_model.tpp:421-464, 550-687).  ncu of the
// DFMA position-table kernel (assemble_th_fast.cu) showed it bound by shared-memory operand traffic (MIO throttle:
// 8 LDS per 10 DFMA); B200's DMMA runs at the full FP64 rate (37.0 vs 33.9 TFLOP/s DFMA,
// profiles/r01_fp64_peaks.json) and needs one operand double per lane per 256 FMA, so the contraction moves to the
// tensor pipe:
//
//   X[q][32*alpha + b] = (d_0 phi_b, d_1 phi_b, d_2 phi_b, phi_b)(x_q), alpha = 0..3, for the 27 Q2 nodes b (padded
//   to 32), then the 8 psi_b'(x_q) at column PSI0;  K = 27 quadrature points padded to 28
//   D[(a,alpha),(b,beta)] = sum_q w_q X[q][32 alpha + a] X[q][32 beta + b]
//
// A warp takes an 8x8 block of node pairs (row block ta <= column block tb: the matrix is symmetric, the transposed
// block is added from the same registers) and keeps all 16 (alpha,beta) tiles, so every lane owns the complete 4x4
// block of its two node pairs and the epilogue is lane-local:
//   diag_ab = D[(a,3),(b,3)] + nu * sum_e D[(a,e),(b,e)]
//   L[(a,c),(b,d)] = nu * D[(a,d),(b,c)] + delta_cd diag_ab                      (:626-632)
//   L[(a,c),p_b'] = L[p_b',(a,c)] = -D[(a,c),psi_b']                             (:633-635)
//   preconditioner: L[(a,c),(b,c)] = diag_ab, L[p_a,p_b] = D[psi_a,psi_b]        (:455-462)
// Homogeneous Dirichlet lines mask rows/columns and keep |L_ii| on the diagonal; no-normal-flux lines (one
// component of a node expressed by the other two) are applied as C^T F C on the 3x3 block -- exactly what
// AffineConstraints::distribute_local_to_global does entry by entry (:677-687).  The 6 of 16 cross terms
// (phi_a d phi_b) of each block are computed and dropped: the tensor pipe has the headroom, the L1TEX pipe (global
// reductions, one lane per clock when the addresses are scattered, plus the shared-memory operand loads) is what
// limits the kernel (ncu: L1TEX 74 % busy, DMMA pipe 23 %).
#include <omp.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "th_mma_common.cuh"

namespace {

using namespace dcpdev;
using namespace thmma;

// Per-cell inputs (mapping record, dof indices, plan row, masks) are copied with cp.async; the gathers that depend
// on the dof indices (row starts, old solution values) are cp.async gathers issued before the operand table is built
// and awaited after it.  A cross-cell double buffer (prefetching cell n+1 while cell n is processed) was measured
// and gave nothing at refine 5 and 6: the kernel is bound by the L1TEX pipe (reductions + shared-memory operands),
// not by the latency of these loads.
template <bool SYSTEM>
__global__ void __launch_bounds__(MTHREADS, 4) th_mma_kernel(MmaArgs a, BlockView A, CsView cs) {
  extern __shared__ __align__(16) double smem[];
  double* X = smem;                     // KQ * LDB, column = 32*alpha + node (alpha: d_0, d_1, d_2, value), psi at PSI0
  double* wq = X + KQ * LDB;            // KQ (+4)
  double* sgeo2 = wq + 32;              // GS
  double* swt = sgeo2 + GS + 1;         // 3*NU master weights of the constrained component
  double* sF = swt + 3 * NU + 1;        // NQ*3   (sF .. sGU: right-hand side only, i.e. unused by the preconditioner)
  double* sU = sF + NQ * 3;             // ND
  double* sT = sU + ND + 1;             // 32: old temperature at the quadrature points
  double* sTn = sT + 32;                // 28: nodal old temperature
  double* sGU = sTn + 28;               // NQ*12 old velocity / gradient at the quadrature points
  long long* rb00 = (long long*)(sGU + NQ * 12);  // 3*NU
  long long* rb01 = rb00 + 3 * NU;          // 3*NU
  long long* rb10 = rb01 + 3 * NU;          // NP
  unsigned short* spos2 = (unsigned short*)(rb10 + NP);       // PSTR
  unsigned char* snm2 = (unsigned char*)(spos2 + PSTR);       // MSTR
  int* sidx2 = (int*)(snm2 + MSTR);                           // IDS
  int* sidt2 = sidx2 + IDS;                                   // 28
  int* sys_u = sidt2 + 28;                                    // 3*NU
  int* sys_p = sys_u + 3 * NU;                                // NP
  unsigned char* skc = (unsigned char*)(sys_p + NP);          // 28
  // preconditioner, cells whose offsets differ between the components: 3*NU*NU offsets in the right-hand-side scratch
  unsigned short* swide = (unsigned short*)sF;
  static_assert(sizeof(double) * (NQ * 3 + ND + 1 + 32 + 28 + NQ * 12) >= sizeof(unsigned short) * 3 * NU * NU, "swide alias");
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;

  for (int i = tid; i < ND; i += nt) {
    const int f = a.local_field[i], bs = a.local_base[i];
    if (f < 3) sys_u[f * NU + bs] = i; else sys_p[bs] = i;
  }
  for (int i = tid; i < KQ * LDB; i += nt) X[i] = 0.0;   // padding nodes 27..31, padded point 27 stay zero
  if (tid < 4) wq[NQ + tid] = 0.0;
  const double nu = a.prm.dt * a.prm.inv_re;
  const bool do_rhs = SYSTEM && a.rhs != nullptr;
  const long long* rp00 = A.rowptr[0][0];
  const long long* rp01 = A.rowptr[0][1];
  const long long* rp10 = SYSTEM ? A.rowptr[1][0] : A.rowptr[1][1];
  double* v00 = A.val[0][0];
  double* v01 = A.val[0][1];
  double* v10 = SYSTEM ? A.val[1][0] : A.val[1][1];

  const double* sgeo = sgeo2;
  const unsigned short* spos = spos2;
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
  auto issue_raw = [&](long long w, long long cell) {
    const double* g = a.geom + cell * GS;
    for (int i = tid; i < GS; i += nt) cp_async8(sgeo2 + i, g + i);
    const unsigned short* p = a.pos + w * PSTR;
    for (int i = tid; i < PSTR / 4; i += nt) cp_async8(spos2 + 4 * i, p + 4 * i);
    if (tid < MSTR / 8) cp_async8(snm2 + 8 * tid, a.nmask + w * MSTR + 8 * tid);
    for (int i = tid; i < ND; i += nt) cp_async4(sidx2 + i, a.l2g + cell * ND + i);
    if (do_rhs)
      for (int i = tid; i < a.ndt; i += nt) cp_async4(sidt2 + i, a.l2g_t + cell * a.ndt + i);
  };

  for (long long wi = blockIdx.x; wi < a.n_fast; wi += gridDim.x) {
    const long long w = a.wlist ? a.wlist[wi] : wi;
    __syncthreads();  // every warp is done with the previous cell
    issue_raw(w, a.cells[w]);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    // gathers through the dof indices: row starts, old solution
    for (int i = tid; i < 3 * NU; i += nt) {
      const int gi = sidx[sys_u[i]];
      cp_async8(rb00 + i, rp00 + gi);
      if (SYSTEM) cp_async8(rb01 + i, rp01 + gi);
    }
    for (int i = tid; i < NP; i += nt) cp_async8(rb10 + i, rp10 + (sidx[sys_p[i]] - a.n_u));
    if (do_rhs) {
      for (int i = tid; i < ND; i += nt) cp_async8(sU + i, a.old_nse + sidx[i]);
      for (int i = tid; i < a.ndt; i += nt) cp_async8(sTn + i, a.old_temp + sidt[i]);
    }
    cp_async_commit();
    const int wide = SYSTEM ? -1 : *reinterpret_cast<const int*>(snm + 36);
    const int nnf = SYSTEM ? -1 : *reinterpret_cast<const int*>(snm + 40);
    if (!SYSTEM && wide >= 0) {
      const unsigned short* p = a.pos_wide + (size_t)wide * (3 * NU * NU);
      for (int i = tid; i < 3 * NU * NU; i += nt) swide[i] = p[i];
    }
    const int cflag = snm[35];  // 1: the cell holds constrained velocity dofs
    if (tid < NU) {
      // no-normal-flux lines: u_k = sum_{c != k} w_c u_c on the same node (verified when the plan was built)
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
    // operand table: physical gradients and values of the Q2 functions (component-major), then the Q1 functions
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
    cp_async_wait<0>();  // the gathers
    __syncthreads();

    // ---- Gram blocks on the tensor cores.  Task (ta <= tb): 8 row nodes x 8 column nodes, all 16 (alpha,beta)
    // tiles in one warp, so every lane ends up with the complete 4x4 block of two node pairs (symmetric half only:
    // the transposed block is added from the same registers).  The preconditioner needs the 4 diagonal tiles only.
    const int frow = lane >> 2, fk = lane & 3;
    for (int t = warp; t < 10 + (SYSTEM ? 4 : 1); t += nwarps) {
      if (t < 10) {
        int ta = 0, r = t;
        while (r >= 4 - ta) { r -= 4 - ta; ++ta; }
        const int tb = ta + r;
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
            bf[i] = xr[32 * i + 8 * tb + frow];
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (SYSTEM || i == k) dmma(acc[i][k][0], acc[i][k][1], af[i], bf[k]);
        }
        const int na = 8 * ta + frow;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int nb = 8 * tb + 2 * fk + jj;
          if (na >= NU || nb >= NU) continue;
          const NodeCs ca = node_cs(na), cb = node_cs(nb);
          const double dg = acc[3][3][jj] + nu * (acc[0][0][jj] + acc[1][1][jj] + acc[2][2][jj]);
          if (SYSTEM) {
            // F[c][d] = L[(a,c),(b,d)] = nu * D[(a,d),(b,c)] + delta_cd diag     (acc[alpha][beta] = D[(a,alpha),(b,beta)])
            double F[3][3];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int d = 0; d < 3; ++d) F[c][d] = nu * acc[d][c][jj] + (c == d ? dg : 0.0);
            const double d00 = fabs(F[0][0]), d11 = fabs(F[1][1]), d22 = fabs(F[2][2]);
            if (ca.k != 3 || cb.k != 3) {
              // C^T F C on the 3x3 block: R[c][d] = F[c][d] + wa_c F[ka][d] + wb_d F[c][kb] + wa_c wb_d F[ka][kb]
              const double wa[3] = {ca.w0, ca.w1, ca.w2}, wb[3] = {cb.w0, cb.w1, cb.w2};
              double Fa[3], Fb[3], Fab;  // F[ka][d], F[c][kb], F[ka][kb]
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
            const long long off_ab = spos[na * NE + nb], off_ba = spos[nb * NE + na];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              if (!(ca.mask & (1 << c))) continue;
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                if (!(cb.mask & (1 << d))) continue;
                red_add_f64(v00 + rb00[c * NU + na] + off_ab + __popc(cb.mask & ((1 << d) - 1)), F[c][d]);
                if (ta != tb)  // transposed block L[(b,d),(a,c)]
                  red_add_f64(v00 + rb00[d * NU + nb] + off_ba + __popc(ca.mask & ((1 << c) - 1)), F[c][d]);
              }
            }
            if (na == nb) {  // constrained dofs keep |L_ii| on their own diagonal (the row holds nothing else)
              if (!(ca.mask & 1)) red_add_f64(v00 + rb00[na], d00);
              if (!(ca.mask & 2)) red_add_f64(v00 + rb00[NU + na], d11);
              if (!(ca.mask & 4)) red_add_f64(v00 + rb00[2 * NU + na], d22);
            }
          } else if (nnf >= 0) {
            // preconditioner cell with no-normal-flux lines: C^T (dg I) C spreads dg over the component pairs that
            // involve a master; positions of all nine pairs come from the cell's own table (global memory, few cells)
            const unsigned short* p9 = a.pos9 + (size_t)nnf * (9 * NU * NU);
            const double wa[3] = {ca.w0, ca.w1, ca.w2}, wb[3] = {cb.w0, cb.w1, cb.w2};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              if (!(ca.mask & (1 << c))) {
                if (na == nb) red_add_f64(v00 + rb00[c * NU + na], fabs(dg));
                continue;
              }
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                if (!(cb.mask & (1 << d))) continue;
                double r = c == d ? 1.0 : 0.0;
                if (ca.k == d) r += wa[c];
                if (cb.k == c) r += wb[d];
                if (ca.k != 3 && ca.k == cb.k) r += wa[c] * wb[d];
                if (r == 0.0) continue;
                r *= dg;
                const unsigned o1 = p9[((c * 3 + d) * NU + na) * NU + nb];
                if (o1 != 0xffffu) red_add_f64(v00 + rb00[c * NU + na] + o1, r);
                if (ta != tb) {
                  const unsigned o2 = p9[((d * 3 + c) * NU + nb) * NU + na];
                  if (o2 != 0xffffu) red_add_f64(v00 + rb00[d * NU + nb] + o2, r);
                }
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              if (ca.mask & cb.mask & (1 << c)) {
                const unsigned o1 = wide >= 0 ? swide[(c * NU + na) * NU + nb] : spos[na * NE + nb];
                if (o1 != 0xffffu) red_add_f64(v00 + rb00[c * NU + na] + o1, dg);
                if (ta != tb) {
                  const unsigned o2 = wide >= 0 ? swide[(c * NU + nb) * NU + na] : spos[nb * NE + na];
                  if (o2 != 0xffffu) red_add_f64(v00 + rb00[c * NU + nb] + o2, dg);
                }
              } else if (na == nb && !(ca.mask & (1 << c)))
                red_add_f64(v00 + rb00[c * NU + na], fabs(dg));
            }
          }
        }
      } else if (SYSTEM) {
        // velocity-pressure coupling: rows (a, alpha < 3) of row block ta against the 8 psi columns
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
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int pb = 2 * fk + jj;
            if (!snm[NU + pb]) continue;
            double sv[3] = {-acc[0][jj], -acc[1][jj], -acc[2][jj]};
            if (ca.k != 3) {
              const double sk = ca.k == 0 ? sv[0] : (ca.k == 1 ? sv[1] : sv[2]);
              sv[0] += ca.w0 * sk;
              sv[1] += ca.w1 * sk;
              sv[2] += ca.w2 * sk;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
              if (ca.mask & (1 << c)) {
                red_add_f64(v01 + rb01[c * NU + na] + spos[na * NE + NU + pb], sv[c]);
                red_add_f64(v10 + rb10[pb] + spos[(NU + pb) * NE + na] + __popc(ca.mask & ((1 << c) - 1)), sv[c]);
              }
          }
        }
      } else {
        // preconditioner: pressure mass matrix, psi x psi
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int ks = 0; ks < KQ / 4; ++ks) {
          const int q = 4 * ks + fk;
          const double* xr = X + q * LDB;
          dmma(c0, c1, wq[q] * xr[PSI0 + frow], xr[PSI0 + frow]);
        }
        const int pa = frow;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int pb = 2 * fk + jj;
          if (snm[NU + pa] && snm[NU + pb]) red_add_f64(v10 + rb10[pa] + spos[(NU + pa) * NE + NU + pb], jj ? c1 : c0);
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
      // old velocity and its gradient at the quadrature points: 27 x 3 x (3 gradients + value) dot products
      for (int i = tid; i < NQ * 12; i += nt) {
        const int q = i / 12, r = i - q * 12, c = r >> 2, e = r & 3;
        const double* x = X + q * LDB + 32 * e;   // e < 3: d_e phi_n, e == 3: phi_n
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
        else {  // constrained dof: its share goes to the masters (none for Dirichlet lines)
          const int li = cs.line_of_dof[gi];
          for (int k = cs.line_ptr[li]; k < cs.line_ptr[li + 1]; ++k) red_add_f64(a.rhs + cs.entry_dof[k], cs.entry_w[k] * s);
        }
      }
    }
  }
}

constexpr size_t mma_smem_bytes(bool system) {
  (void)system;
  return sizeof(double) * (KQ * LDB + 32 + GS + 1 + NQ * 3 + ND + 1 + 32 + 28 + 3 * NU + 1 + NQ * 12) + sizeof(long long) * (6 * NU + NP) +
         sizeof(unsigned short) * PSTR + MSTR + sizeof(int) * (IDS + 28 + 3 * NU + NP) + 28 + 36;
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
  cudaFree(p->pos_wide);
  cudaFree(p->pos9);
  dcp_gather_plan_free(p->gather);
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
  std::vector<uint8_t> ok((size_t)nc, 0), is_wide((size_t)nc, 0), is_nnf((size_t)nc, 0);
  std::vector<std::vector<uint16_t>> nnf_rows((size_t)nc);
  std::vector<uint16_t> pos_all((size_t)nc * NE * NE, 0xFFFF);
  std::vector<uint8_t> mask_all((size_t)nc * 36, 0);
  std::vector<std::vector<uint16_t>> wide_rows((size_t)nc);
  std::vector<uint16_t> cell_max_off((size_t)nc, 0);   // preconditioner, cells without no-normal-flux lines: largest row offset
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
    bool any_cs = false, any_master = false;
    int kcomp[NU];
    for (int a = 0; a < NU; ++a) kcomp[a] = 3;
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
          any_master = true;   // preconditioner: redistributed entries leave the same-component pattern (table of 9 below)
          kcomp[a] = k;
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
    } else if (any_master) {
      // preconditioner, no-normal-flux lines: C^T (dg I) C fills the component pairs (c,d) with c == d, d == k_a,
      // c == k_b, or k_a == k_b; every such entry must exist in the pattern, else the cell goes to the general kernel
      std::vector<uint16_t> W9(9 * NU * NU, 0xFFFF);
      for (int a = 0; a < NU && good; ++a)
        for (int b = 0; b < NU && good; ++b)
          for (int cc = 0; cc < 3 && good; ++cc) {
            if (!(mk[a] & (1 << cc))) continue;
            for (int dd = 0; dd < 3 && good; ++dd) {
              if (!(mk[b] & (1 << dd))) continue;
              const int64_t off = A00.find((int64_t)idx[sys_u[a]] + cc, idx[sys_u[b]] + dd);
              const bool needed = cc == dd || kcomp[a] == dd || kcomp[b] == cc || (kcomp[a] != 3 && kcomp[a] == kcomp[b]);
              if (off >= 0 && off < 65535)
                W9[((cc * 3 + dd) * NU + a) * NU + b] = (uint16_t)off;
              else if (needed)
                good = false;
            }
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
      if (good) {
        is_nnf[c] = 1;
        nnf_rows[c].swap(W9);
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
            if (off > cell_max_off[c]) cell_max_off[c] = (uint16_t)off;
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
  std::vector<int32_t> cells, other, wide_idx, nnf_idx;
  std::vector<uint16_t> pos, pos_wide, pos9;
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
    if (is_nnf[c]) {
      nnf_idx.push_back((int32_t)(pos9.size() / (9 * NU * NU)));
      pos9.insert(pos9.end(), nnf_rows[c].begin(), nnf_rows[c].end());
      std::vector<uint16_t>().swap(nnf_rows[c]);
    } else
      nnf_idx.push_back(-1);
  }
  pos.resize(cells.size() * (size_t)PSTR, 0xFFFF);
  nmask.resize(cells.size() * (size_t)MSTR, 0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)cells.size(); ++i) {
    std::copy(&pos_all[(size_t)cells[i] * NE * NE], &pos_all[(size_t)cells[i] * NE * NE] + NE * NE, &pos[(size_t)i * PSTR]);
    std::copy(&mask_all[(size_t)cells[i] * 36], &mask_all[(size_t)cells[i] * 36] + 36, &nmask[(size_t)i * MSTR]);
    std::memcpy(&nmask[(size_t)i * MSTR + 36], &wide_idx[i], sizeof(int32_t));
    std::memcpy(&nmask[(size_t)i * MSTR + 40], &nnf_idx[i], sizeof(int32_t));
  }
  MaskedPlan* P = new MaskedPlan;
  P->h_cells = cells;
  P->h_nnf_idx = nnf_idx;
  P->h_wide_idx = wide_idx;
  P->max_off_plain = 0;
  for (size_t i = 0; i < cells.size(); ++i)
    if (nnf_idx[i] < 0) P->max_off_plain = std::max<int>(P->max_off_plain, cell_max_off[cells[i]]);
  P->h_cflag.resize(cells.size());
  for (size_t i = 0; i < cells.size(); ++i) P->h_cflag[i] = mask_all[(size_t)cells[i] * 36 + 35];
  P->n = (int64_t)cells.size();
  P->n_other = (int64_t)other.size();
  P->n_wide = (int64_t)(pos_wide.size() / (3 * NU * NU));
  dcp_ctx* ctx = m->ctx;
  int rc = upm(ctx, &P->cells, cells);
  if (rc == DCP_OK) rc = upm(ctx, &P->other_cells, other);
  if (rc == DCP_OK) rc = upm(ctx, &P->pos, pos);
  if (rc == DCP_OK) rc = upm(ctx, &P->nmask, nmask);
  if (rc == DCP_OK) rc = upm(ctx, &P->pos_wide, pos_wide);
  if (rc == DCP_OK) rc = upm(ctx, &P->pos9, pos9);
  cudaStreamSynchronize(ctx->stream);
  if (rc == DCP_OK && system && other.empty() && !std::getenv("DCP_NO_GATHER_PLAN")) {
    std::vector<uint8_t> has_cs(cells.size());
    for (size_t i = 0; i < cells.size(); ++i) has_cs[i] = mask_all[(size_t)cells[i] * 36 + 35];
    rc = dcp_gather_plan_build(m, d, cells, has_cs, &P->gather);
  }
  if (rc != DCP_OK) {
    dcp_masked_plan_free(P);
    return rc;
  }
  *out = P;
  return DCP_OK;
}

int dcp_launch_th_mma(dcp_model* m, const dcp_params& p, bool system, const MaskedPlan* plan, const double* old_nse,
                      const double* old_temp, const int32_t* wlist, int64_t n_list) {
  dcp_ctx* ctx = m->ctx;
  MmaArgs a;
  a.n_fast = wlist ? n_list : plan->n;
  a.wlist = wlist;
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
  a.rhs = system ? m->nse_rhs : nullptr;
  a.n_u = m->nse.start[1];
  a.prm = p;
  if (a.n_fast == 0) return DCP_OK;
  const size_t smem = mma_smem_bytes(system);
  auto launch = [&](auto kernel) -> int {
    DCP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    DCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, MTHREADS, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)ctx->sm_count * per_sm;
    if (grid > a.n_fast) grid = a.n_fast;
    const BlockMat& mat = system ? m->nse : m->pre;
    kernel<<<(unsigned)grid, MTHREADS, smem, ctx->stream>>>(a, make_view(mat), make_view(m->nse_cs));
    return DCP_OK;
  };
  if (system)
    DCP_TRY(launch(th_mma_kernel<true>));
  else
    DCP_TRY(launch(th_mma_kernel<false>));
  ctx->launches++;
  DCP_CUDA(cudaGetLastError());
  return DCP_OK;
}
