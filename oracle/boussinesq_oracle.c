/* boussinesq_oracle.c -- CPU restatement of the reference's hot path (classic Taylor-Hood family).
 *
 * TEST INFRASTRUCTURE ONLY.  Imported/linked/executed only by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  The product (libdcp.so) never calls into this file.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference ships no golden vectors, known-answer tests or
 * fixtures for this path (its only test prints a string, /root/reference/test/test_dummy.cc:19-42) and it
 * cannot be compiled here (needs deal.II>=9.2 + Trilinos + p4est + MPI + TBB).  The oracle is pinned
 * instead by (1) analytic identities (tests/test_oracle_identities.py), (2) an independent numpy
 * re-derivation (tests/test_oracle_independent.py) and (3) frozen golden fixtures produced by this file
 * (tests/golden/, generator committed).
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).  The loops are
 * deliberately the reference's dense i,j,q loops over FEValues-style per-dof views, not the
 * structure-exploiting form the CUDA kernels use.
 *
 * Conventions restated from deal.II 9.2 (un-vendored dependency; "from memory"):
 *   geometry record per cell: [JxW(nq) | Kinv[e][d](nq each) | xq[d](nq each)], Kinv[e][d] = d xi_e/d x_d,
 *   mapped gradient  d_d phi = sum_e Kinv[e][d] * dhat_e phi   (covariant transform of FE_Q).
 *   AffineConstraints::distribute_local_to_global: exact-zero local entries are skipped; constrained
 *   rows/columns are redistributed to their masters with weights; inhomogeneities go to the rhs;
 *   a constrained dof gets |L_ii| (or the mean |diag| if L_ii == 0) on its own diagonal.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_common.h"

static int g_orc_missing = 0; /* counts scatter targets that are not in the pattern */
int orc_missing_entries(void) { return g_orc_missing; }
void orc_reset_missing(void) { g_orc_missing = 0; }

void orc_csr_add(const orc_csr* A, int64_t r, int32_t c, double v, int atomic) {
  const int32_t* b = A->col + A->rowptr[r];
  int64_t lo = 0, hi = A->rowptr[r + 1] - A->rowptr[r];
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (b[mid] < c) lo = mid + 1; else hi = mid;
  }
  if (lo == A->rowptr[r + 1] - A->rowptr[r] || b[lo] != c) {
#pragma omp atomic
    g_orc_missing++;
    return;
  }
  double* p = A->val + A->rowptr[r] + lo;
  if (atomic) {
#pragma omp atomic
    *p += v;
  } else
    *p += v;
}

void orc_vec_add(double* b, int64_t i, double v, int atomic) {
  if (atomic) {
#pragma omp atomic
    b[i] += v;
  } else
    b[i] += v;
}

/* AffineConstraints::distribute_local_to_global(local_matrix[, local_vector], indices, A[, b]).
 * Call sites: boussinesq_model.tpp:473-475 (matrix only), :682-686 (matrix+rhs), :809-816. */
void orc_distribute_matrix(const orc_constraints* cs, int n, const double* L, const double* l, const int32_t* idx,
                              const orc_csr* A, double* b, int atomic) {
  int any_constrained = 0;
  for (int i = 0; i < n; ++i) {
    int32_t gi = idx[i], li = cs->line_of_dof[gi];
    int32_t one_r = gi;
    double one_w = 1.0;
    const int32_t* rd = &one_r;
    const double* rw = &one_w;
    int nr = 1;
    if (li >= 0) {
      any_constrained = 1;
      nr = cs->line_ptr[li + 1] - cs->line_ptr[li];
      rd = cs->entry_dof + cs->line_ptr[li];
      rw = cs->entry_w + cs->line_ptr[li];
    }
    double rhs_i = l ? l[i] : 0.0;
    for (int j = 0; j < n; ++j) {
      double v = L[i * n + j];
      if (v == 0.0) continue;
      int32_t gj = idx[j], lj = cs->line_of_dof[gj];
      if (lj < 0) {
        for (int a = 0; a < nr; ++a) orc_csr_add(A, rd[a], gj, rw[a] * v, atomic);
      } else {
        int nc = cs->line_ptr[lj + 1] - cs->line_ptr[lj];
        const int32_t* cd = cs->entry_dof + cs->line_ptr[lj];
        const double* cw = cs->entry_w + cs->line_ptr[lj];
        for (int a = 0; a < nr; ++a)
          for (int e = 0; e < nc; ++e) orc_csr_add(A, rd[a], cd[e], rw[a] * cw[e] * v, atomic);
        if (l) rhs_i -= v * cs->inhom[lj];
      }
    }
    if (b)
      for (int a = 0; a < nr; ++a) orc_vec_add(b, rd[a], rw[a] * rhs_i, atomic);
  }
  if (any_constrained) {
    double avg = 0;
    for (int i = 0; i < n; ++i) avg += fabs(L[i * n + i]);
    avg /= n;
    for (int i = 0; i < n; ++i) {
      int32_t gi = idx[i];
      if (cs->line_of_dof[gi] < 0) continue;
      double d = fabs(L[i * n + i]);
      orc_csr_add(A, gi, gi, d != 0.0 ? d : avg, atomic);
    }
  }
}

/* AffineConstraints::distribute_local_to_global(local_vector, indices, global_vector, local_matrix)
 * -- the "matrix_for_bc" overload used at boussinesq_model.tpp:960-963. */
void orc_distribute_vector_bc(const orc_constraints* cs, int n, const double* l, const double* Lbc,
                                 const int32_t* idx, double* b, int atomic) {
  for (int i = 0; i < n; ++i) {
    int32_t gi = idx[i], li = cs->line_of_dof[gi];
    if (li < 0) {
      orc_vec_add(b, gi, l[i], atomic);
      continue;
    }
    double val = cs->inhom[li];
    if (val != 0.0)
      for (int j = 0; j < n; ++j) {
        int32_t gj = idx[j], lj = cs->line_of_dof[gj];
        if (lj < 0) {
          orc_vec_add(b, gj, -val * Lbc[j * n + i], atomic);
          continue;
        }
        double m = Lbc[j * n + i];
        if (m == 0.0) continue;
        for (int32_t e = cs->line_ptr[lj]; e < cs->line_ptr[lj + 1]; ++e)
          orc_vec_add(b, cs->entry_dof[e], -val * cs->entry_w[e] * m, atomic);
      }
    for (int32_t e = cs->line_ptr[li]; e < cs->line_ptr[li + 1]; ++e)
      orc_vec_add(b, cs->entry_dof[e], l[i] * cs->entry_w[e], atomic);
  }
}

/* CoreModelData::gravity_vector / vertical_gravity_vector (include/model_data/core_model_data.tpp:86-106) */
void orc_gravity(const orc_params* P, const double* x, double* g) {
  int dim = P->dim;
  if (P->cuboid) {
    for (int d = 0; d < dim; ++d) g[d] = 0;
    g[dim - 1] = -P->g_const;
    return;
  }
  double r = 0;
  for (int d = 0; d < dim; ++d) r += x[d] * x[d];
  r = sqrt(r);
  double s = r > 1 ? r : sqrt(r);
  for (int d = 0; d < dim; ++d) g[d] = -P->g_const * x[d] / s;
}

/* per-cell FEValues-equivalent of the Taylor-Hood FESystem: one primitive shape function per local dof.
 * field[k] in [0,dim) = velocity component, field[k]==dim = pressure; base[k] = index in the scalar
 * base element (hierarchical numbering of the Q2 nodes; vertex number for Q1). */
typedef struct {
  int dim, nd, nq, ndu, ndp;
  const int32_t* field;
  const int32_t* base;
  const double* phi_u;  /* [nq][ndu] */
  const double* dphi_u; /* [nq][ndu][dim] */
  const double* phi_p;  /* [nq][ndp] */
} th_fe;

static inline void mapped_grad(int dim, int nq, const double* geom, int q, const double* dref, double* g) {
  for (int d = 0; d < dim; ++d) {
    double s = 0;
    for (int e = 0; e < dim; ++e) s += geom[nq * (1 + e * dim + d) + q] * dref[e];
    g[d] = s;
  }
}

/* Standard::BoussinesqModel::local_assemble_nse_system (include/core/boussinesq_model.tpp:550-673)
 * + copy_local_to_global_nse_system (:677-687), driven like assemble_nse_system (:691-740).
 * A is the whole (block-concatenated) nse_matrix pattern; rhs has n_u+n_p entries. Both are zeroed here
 * like `nse_matrix = 0; nse_rhs = 0` (:700-706). */
void orc_assemble_nse_system(const orc_params* P, int64_t n_cells, int nd, int nq, int ndu, int ndp, int ndt,
                             const int32_t* field, const int32_t* base, const double* phi_u, const double* dphi_u,
                             const double* phi_p, const double* phi_t, const double* geom, const int32_t* l2g,
                             const int32_t* l2g_t, const double* old_nse, const double* old_temp,
                             const orc_constraints* cs, orc_csr* A, double* rhs, int64_t n_rhs, int use_omp) {
  const int dim = P->dim;
  const int gs = nq * (1 + dim * dim + dim);
  memset(A->val, 0, sizeof(double) * (size_t)A->rowptr[A->n_rows]);
  memset(rhs, 0, sizeof(double) * (size_t)n_rhs);
  (void)ndu;
  (void)ndp;
#pragma omp parallel if (use_omp)
  {
    double* L = (double*)malloc(sizeof(double) * nd * nd);
    double l[ORC_MAXD];
    double phiu[ORC_MAXD][3], sym[ORC_MAXD][3][3], divu[ORC_MAXD], phip[ORC_MAXD], gradu[ORC_MAXD][3][3];
#pragma omp for schedule(dynamic, 16)
    for (int64_t c = 0; c < n_cells; ++c) {
      const double* g = geom + (size_t)c * gs;
      const int32_t* idx = l2g + (size_t)c * nd;
      const int32_t* idt = l2g_t + (size_t)c * ndt;
      memset(L, 0, sizeof(double) * nd * nd);
      memset(l, 0, sizeof(l));
      for (int q = 0; q < nq; ++q) {
        /* fe_values views (:602-613) */
        for (int k = 0; k < nd; ++k) {
          for (int a = 0; a < 3; ++a) {
            phiu[k][a] = 0;
            for (int b2 = 0; b2 < 3; ++b2) gradu[k][a][b2] = 0;
          }
          divu[k] = 0;
          phip[k] = 0;
          int f = field[k], a = base[k];
          if (f < dim) {
            double gr[3] = {0, 0, 0};
            mapped_grad(dim, nq, g, q, dphi_u + ((size_t)q * ndu + a) * dim, gr);
            phiu[k][f] = phi_u[(size_t)q * ndu + a];
            for (int d = 0; d < dim; ++d) gradu[k][f][d] = gr[d];
            divu[k] = gr[f];
          } else
            phip[k] = phi_p[(size_t)q * ndp + a];
          for (int a2 = 0; a2 < dim; ++a2)
            for (int b2 = 0; b2 < dim; ++b2) sym[k][a2][b2] = 0.5 * (gradu[k][a2][b2] + gradu[k][b2][a2]);
        }
        /* get_function_values / gradients of the old solution (:583-589) */
        double oldT = 0, oldu[3] = {0, 0, 0}, oldgu[3][3] = {{0}};
        for (int k = 0; k < ndt; ++k) oldT += old_temp[idt[k]] * phi_t[(size_t)q * ndt + k];
        for (int k = 0; k < nd; ++k) {
          double U = old_nse[idx[k]];
          for (int a = 0; a < dim; ++a) {
            oldu[a] += U * phiu[k][a];
            for (int b2 = 0; b2 < dim; ++b2) oldgu[a][b2] += U * gradu[k][a][b2];
          }
        }
        const double density_scaling = 1 - P->beta * (oldT - P->T_ref); /* core_model_data.cc:88-94 */
        const double JxW = g[q];
        /* matrix (:626-637) */
        for (int i = 0; i < nd; ++i)
          for (int j = 0; j < nd; ++j) {
            double uu = 0, ee = 0;
            for (int a = 0; a < dim; ++a) {
              uu += phiu[i][a] * phiu[j][a];
              for (int b2 = 0; b2 < dim; ++b2) ee += sym[i][a][b2] * sym[j][a][b2];
            }
            L[i * nd + j] += (uu + P->dt * (P->inv_re * 2 * ee) - divu[i] * phip[j] - phip[i] * divu[j]) * JxW;
          }
        /* gravity, coriolis (:615-621, 640-650) */
        double xq[3] = {0, 0, 0}, grav[3] = {0, 0, 0}, cor[3] = {0, 0, 0};
        for (int d = 0; d < dim; ++d) xq[d] = g[nq * (1 + dim * dim + d) + q];
        orc_gravity(P, xq, grav);
        for (int d = 0; d < dim; ++d) grav[d] *= P->g_scale;
        if (P->cuboid) cor[dim - 1] = P->cor_scale * P->omega; /* L * coriolis_vector / U */
        /* advection: old_velocity * transpose(grad u)  ->  (u . grad) u  (:599-600, 660) */
        double adv[3] = {0, 0, 0};
        for (int a = 0; a < dim; ++a)
          for (int d = 0; d < dim; ++d) adv[a] += oldu[d] * oldgu[a][d];
        double cterm[3] = {0, 0, 0};
        if (dim == 2) { /* -2 * phi . cross_product_2d(u), cross_product_2d(u) = (u_y, -u_x) */
          cterm[0] = -2 * oldu[1];
          cterm[1] = 2 * oldu[0];
        } else {
          cterm[0] = 2 * (cor[1] * oldu[2] - cor[2] * oldu[1]);
          cterm[1] = 2 * (cor[2] * oldu[0] - cor[0] * oldu[2]);
          cterm[2] = 2 * (cor[0] * oldu[1] - cor[1] * oldu[0]);
        }
        /* rhs (:655-669) */
        for (int i = 0; i < nd; ++i) {
          double a1 = 0, a2 = 0, a3 = 0, a4 = 0;
          for (int a = 0; a < dim; ++a) {
            a1 += phiu[i][a] * oldu[a];
            a2 += grav[a] * phiu[i][a];
            a3 += phiu[i][a] * adv[a];
            a4 += phiu[i][a] * cterm[a];
          }
          l[i] += (a1 + P->dt * density_scaling * a2 - P->dt * a3 - P->dt * a4) * JxW;
        }
      }
      orc_distribute_matrix(cs, nd, L, l, idx, A, rhs, use_omp);
    }
    free(L);
  }
}

/* Standard::BoussinesqModel::local_assemble_nse_preconditioner (boussinesq_model.tpp:421-464) + copier
 * (:468-476), driver :479-514 */
void orc_assemble_nse_preconditioner(const orc_params* P, int64_t n_cells, int nd, int nq, int ndu, int ndp,
                                     const int32_t* field, const int32_t* base, const double* phi_u,
                                     const double* dphi_u, const double* phi_p, const double* geom,
                                     const int32_t* l2g, const orc_constraints* cs, orc_csr* A, int use_omp) {
  const int dim = P->dim;
  const int gs = nq * (1 + dim * dim + dim);
  memset(A->val, 0, sizeof(double) * (size_t)A->rowptr[A->n_rows]);
#pragma omp parallel if (use_omp)
  {
    double* L = (double*)malloc(sizeof(double) * nd * nd);
    double phiu[ORC_MAXD][3], gradu[ORC_MAXD][3][3], phip[ORC_MAXD];
#pragma omp for schedule(dynamic, 16)
    for (int64_t c = 0; c < n_cells; ++c) {
      const double* g = geom + (size_t)c * gs;
      const int32_t* idx = l2g + (size_t)c * nd;
      memset(L, 0, sizeof(double) * nd * nd);
      for (int q = 0; q < nq; ++q) {
        for (int k = 0; k < nd; ++k) {
          for (int a = 0; a < 3; ++a) {
            phiu[k][a] = 0;
            for (int b2 = 0; b2 < 3; ++b2) gradu[k][a][b2] = 0;
          }
          phip[k] = 0;
          int f = field[k], a = base[k];
          if (f < dim) {
            double gr[3] = {0, 0, 0};
            mapped_grad(dim, nq, g, q, dphi_u + ((size_t)q * ndu + a) * dim, gr);
            phiu[k][f] = phi_u[(size_t)q * ndu + a];
            for (int d = 0; d < dim; ++d) gradu[k][f][d] = gr[d];
          } else
            phip[k] = phi_p[(size_t)q * ndp + a];
        }
        const double JxW = g[q];
        for (int i = 0; i < nd; ++i)
          for (int j = 0; j < nd; ++j) {
            double uu = 0, gg = 0;
            for (int a = 0; a < dim; ++a) {
              uu += phiu[i][a] * phiu[j][a];
              for (int b2 = 0; b2 < dim; ++b2) gg += gradu[i][a][b2] * gradu[j][a][b2];
            }
            L[i * nd + j] += (uu + P->dt * P->inv_re * gg + phip[i] * phip[j]) * JxW;
          }
      }
      orc_distribute_matrix(cs, nd, L, NULL, idx, A, NULL, use_omp);
    }
    free(L);
  }
}

/* local_assemble_temperature_matrix (boussinesq_model.tpp:748-800) + copier (:804-817), driver :821-864 */
void orc_assemble_temperature_matrix(const orc_params* P, int64_t n_cells, int nd, int nq, const double* phi,
                                     const double* dphi, const double* geom, const int32_t* l2g,
                                     const orc_constraints* cs, orc_csr* Mass, orc_csr* Stiff, int use_omp) {
  const int dim = P->dim;
  const int gs = nq * (1 + dim * dim + dim);
  memset(Mass->val, 0, sizeof(double) * (size_t)Mass->rowptr[Mass->n_rows]);
  memset(Stiff->val, 0, sizeof(double) * (size_t)Stiff->rowptr[Stiff->n_rows]);
#pragma omp parallel if (use_omp)
  {
    double* LM = (double*)malloc(sizeof(double) * nd * nd);
    double* LK = (double*)malloc(sizeof(double) * nd * nd);
    double gr[ORC_MAXD][3], ph[ORC_MAXD];
#pragma omp for schedule(dynamic, 64)
    for (int64_t c = 0; c < n_cells; ++c) {
      const double* g = geom + (size_t)c * gs;
      memset(LM, 0, sizeof(double) * nd * nd);
      memset(LK, 0, sizeof(double) * nd * nd);
      for (int q = 0; q < nq; ++q) {
        for (int k = 0; k < nd; ++k) {
          gr[k][0] = gr[k][1] = gr[k][2] = 0;
          mapped_grad(dim, nq, g, q, dphi + ((size_t)q * nd + k) * dim, gr[k]);
          ph[k] = phi[(size_t)q * nd + k];
        }
        for (int i = 0; i < nd; ++i)
          for (int j = 0; j < nd; ++j) {
            double gg = 0;
            for (int d = 0; d < dim; ++d) gg += gr[i][d] * gr[j][d];
            LM[i * nd + j] += ph[i] * ph[j] * g[q];
            LK[i * nd + j] += P->inv_pe * gg * g[q];
          }
      }
      orc_distribute_matrix(cs, nd, LM, NULL, l2g + (size_t)c * nd, Mass, NULL, use_omp);
      orc_distribute_matrix(cs, nd, LK, NULL, l2g + (size_t)c * nd, Stiff, NULL, use_omp);
    }
    free(LM);
    free(LK);
  }
}

/* temperature_matrix.copy_from(mass); temperature_matrix.add(dt/n, stiffness)  (boussinesq_model.tpp:975-978) */
void orc_temperature_matrix_combine(int64_t nnz, const double* mass, const double* stiff, double factor, double* out) {
  for (int64_t i = 0; i < nnz; ++i) out[i] = mass[i] + factor * stiff[i];
}

/* local_assemble_temperature_rhs (boussinesq_model.tpp:873-952) + copier (:955-964), driver :988-1017.
 * nse_solution is the NEW velocity (quirk Q14), old_temp the old temperature. */
void orc_assemble_temperature_rhs(const orc_params* P, int64_t n_cells, int nd, int nq, int nd_nse, int ndu,
                                  const double* phi, const double* dphi, const int32_t* field_nse,
                                  const int32_t* base_nse, const double* phi_u, const double* geom,
                                  const int32_t* l2g, const int32_t* l2g_nse, const double* old_temp,
                                  const double* nse_solution, const orc_constraints* cs, double* rhs, int64_t n_rhs,
                                  int use_omp) {
  const int dim = P->dim;
  const int gs = nq * (1 + dim * dim + dim);
  const double tau = P->dt / P->nse_interval;
  memset(rhs, 0, sizeof(double) * (size_t)n_rhs);
#pragma omp parallel if (use_omp)
  {
    double* Lbc = (double*)malloc(sizeof(double) * nd * nd);
    double l[ORC_MAXD], gr[ORC_MAXD][3], ph[ORC_MAXD];
#pragma omp for schedule(dynamic, 64)
    for (int64_t c = 0; c < n_cells; ++c) {
      const double* g = geom + (size_t)c * gs;
      const int32_t* idx = l2g + (size_t)c * nd;
      const int32_t* idn = l2g_nse + (size_t)c * nd_nse;
      memset(Lbc, 0, sizeof(double) * nd * nd);
      memset(l, 0, sizeof(l));
      for (int q = 0; q < nq; ++q) {
        double oldT = 0, gT[3] = {0, 0, 0}, u[3] = {0, 0, 0};
        for (int k = 0; k < nd; ++k) {
          gr[k][0] = gr[k][1] = gr[k][2] = 0;
          mapped_grad(dim, nq, g, q, dphi + ((size_t)q * nd + k) * dim, gr[k]);
          ph[k] = phi[(size_t)q * nd + k];
          double T = old_temp[idx[k]];
          oldT += T * ph[k];
          for (int d = 0; d < dim; ++d) gT[d] += T * gr[k][d];
        }
        for (int k = 0; k < nd_nse; ++k) {
          int f = field_nse[k];
          if (f < dim) u[f] += nse_solution[idn[k]] * phi_u[(size_t)q * ndu + base_nse[k]];
        }
        const double gamma = 0; /* (L/(U*T_ref)) * 0, :922-926 (quirk Q3) */
        double ugT = 0;
        for (int d = 0; d < dim; ++d) ugT += u[d] * gT[d];
        for (int i = 0; i < nd; ++i) {
          l[i] += (ph[i] * oldT - tau * ph[i] * ugT - tau * gamma * ph[i]) * g[q];
          int32_t li = cs->line_of_dof[idx[i]];
          if (li >= 0 && cs->inhom[li] != 0.0) /* is_inhomogeneously_constrained (:939) */
            for (int j = 0; j < nd; ++j) {
              double gg = 0;
              for (int d = 0; d < dim; ++d) gg += gr[i][d] * gr[j][d];
              Lbc[j * nd + i] += (ph[i] * ph[j] + tau * P->inv_pe * gg) * g[q];
            }
        }
      }
      orc_distribute_vector_bc(cs, nd, l, Lbc, idx, rhs, use_omp);
    }
    free(Lbc);
  }
}

/* Epetra CrsMatrix::Multiply as used by LA::SparseMatrix::vmult / vmult_add (call sites:
 * include/linear_algebra/schur_complement.hpp:147-149, block_schur_preconditioner.hpp:55, ...) */
void orc_spmv(int64_t n_rows, const int64_t* rowptr, const int32_t* col, const double* val, const double* x,
              double* y, int add, int use_omp) {
#pragma omp parallel for schedule(static) if (use_omp)
  for (int64_t r = 0; r < n_rows; ++r) {
    double s = 0;
    for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) s += val[p] * x[col[p]];
    y[r] = add ? y[r] + s : s;
  }
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
