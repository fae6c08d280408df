/* feec_oracle.c -- CPU restatement of the reference's FEEC hot path
 * (ExteriorCalculus::BoussinesqModel<3>, /root/reference/include/core/boussineq_model_FEEC.tpp).
 *
 * TEST INFRASTRUCTURE ONLY (see boussinesq_oracle.c).  PARITY UNPINNED BY THE REFERENCE; pinned by
 * tests/test_feec_oracle.py (identities + independent numpy re-derivation) and golden fixtures.
 *
 * FESystem(FE_Nedelec(0), FE_RaviartThomas(0), FE_DGQ(0)): 12 line dofs (vorticity w), 6 face dofs (velocity u),
 * 1 cell dof (pressure p).  Restated deal.II conventions (un-vendored, from memory): Nedelec is mapped
 * covariantly, phi = J^{-T} phi_hat, curl phi = J curl_hat(phi_hat) / det J; Raviart-Thomas by the contravariant
 * Piola transform, phi = J phi_hat / det J, div phi = div_hat(phi_hat) / det J; FE_DGQ unmapped.
 * Extended geometry record per cell: [JxW | Kinv[e][d] | xq[d] | J[i][j] | detJ], nq entries each.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_common.h"

#define NW 12
#define NU 6
#define ND 19

typedef struct {
  double phi_w[ND][3], curl_w[ND][3], phi_u[ND][3], div_u[ND], phi_p[ND];
  double raw_u[ND][3]; /* RT values without the face sign: what get_function_values sees */
} feec_views;

/* fe_values[vorticity].value/curl, sign_change * fe_values[velocities].value/divergence, fe_values[pressure].value
 * at quadrature point q (boussineq_model_FEEC.tpp:540-549, 721-737) */
static void feec_point(int nq, const double* g, int q, const double* tw, const double* tc, const double* tu,
                       const double* td, const double* sign, feec_views* v) {
  double J[3][3], K[3][3], det = g[nq * 22 + q];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      K[i][j] = g[nq * (1 + i * 3 + j) + q];
      J[i][j] = g[nq * (13 + i * 3 + j) + q];
    }
  memset(v, 0, sizeof(*v));
  for (int k = 0; k < NW; ++k) {
    const double* ph = tw + ((size_t)q * NW + k) * 3;
    const double* ch = tc + ((size_t)q * NW + k) * 3;
    for (int d = 0; d < 3; ++d) {
      double s = 0, c = 0;
      for (int e = 0; e < 3; ++e) {
        s += K[e][d] * ph[e];      /* J^{-T} phi_hat */
        c += J[d][e] * ch[e];      /* J curl_hat / det */
      }
      v->phi_w[k][d] = s;
      v->curl_w[k][d] = c / det;
    }
  }
  for (int k = 0; k < NU; ++k) {
    const double* ph = tu + ((size_t)q * NU + k) * 3;
    for (int d = 0; d < 3; ++d) {
      double s = 0;
      for (int e = 0; e < 3; ++e) s += J[d][e] * ph[e];
      v->raw_u[NW + k][d] = s / det;
      v->phi_u[NW + k][d] = sign[NW + k] * s / det;
    }
    v->div_u[NW + k] = sign[NW + k] * td[k] / det;
  }
  v->phi_p[NW + NU] = 1.0;
}

static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* ExteriorCalculus::BoussinesqModel::local_assemble_nse_system (boussineq_model_FEEC.tpp:669-808) + copier
 * (:812-822), driver :826-875 */
void orc_feec_assemble_nse_system(const orc_params* P, int64_t n_cells, int nq, int ndt, const double* tw,
                                  const double* tc, const double* tu, const double* td, const double* phi_t,
                                  const double* geom, const double* sign, const int32_t* l2g, const int32_t* l2g_t,
                                  const double* old_nse, const double* old_temp, const orc_constraints* cs,
                                  orc_csr* A, double* rhs, int64_t n_rhs, int use_omp) {
  const int gs = nq * 23;
  memset(A->val, 0, sizeof(double) * (size_t)A->rowptr[A->n_rows]);
  memset(rhs, 0, sizeof(double) * (size_t)n_rhs);
#pragma omp parallel if (use_omp)
  {
    double L[ND * ND], l[ND];
    feec_views v;
#pragma omp for schedule(dynamic, 32)
    for (int64_t c = 0; c < n_cells; ++c) {
      const double* g = geom + (size_t)c * gs;
      const int32_t* idx = l2g + (size_t)c * ND;
      const int32_t* idt = l2g_t + (size_t)c * ndt;
      const double* sg = sign + (size_t)c * ND;
      memset(L, 0, sizeof(L));
      memset(l, 0, sizeof(l));
      for (int q = 0; q < nq; ++q) {
        feec_point(nq, g, q, tw, tc, tu, td, sg, &v);
        double oldT = 0, oldw[3] = {0, 0, 0}, oldu[3] = {0, 0, 0};
        for (int k = 0; k < ndt; ++k) oldT += old_temp[idt[k]] * phi_t[(size_t)q * ndt + k];
        for (int k = 0; k < ND; ++k) { /* get_function_values: unsigned shape functions (:705-708) */
          double U = old_nse[idx[k]];
          for (int d = 0; d < 3; ++d) {
            oldw[d] += U * v.phi_w[k][d];
            oldu[d] += U * v.raw_u[k][d];
          }
        }
        const double density_scaling = 1 - P->beta * (oldT - P->T_ref);
        const double JxW = g[q];
        for (int i = 0; i < ND; ++i)
          for (int j = 0; j < ND; ++j)
            L[i * ND + j] += (dot3(v.phi_w[i], v.phi_w[j]) - dot3(v.curl_w[i], v.phi_u[j]) +
                              dot3(v.phi_u[i], v.phi_u[j]) + P->dt * P->inv_re * dot3(v.phi_u[i], v.curl_w[j]) -
                              v.div_u[i] * v.phi_p[j] - v.phi_p[i] * v.div_u[j]) * JxW; /* :753-769 */
        double xq[3], grav[3], cor[3] = {0, 0, 0};
        for (int d = 0; d < 3; ++d) xq[d] = g[nq * (10 + d) + q];
        orc_gravity(P, xq, grav);
        for (int d = 0; d < 3; ++d) grav[d] *= P->g_scale;
        if (P->cuboid) cor[2] = P->cor_scale * P->omega;
        double wxu[3] = {oldw[1] * oldu[2] - oldw[2] * oldu[1], oldw[2] * oldu[0] - oldw[0] * oldu[2],
                         oldw[0] * oldu[1] - oldw[1] * oldu[0]};
        double cxu[3] = {cor[1] * oldu[2] - cor[2] * oldu[1], cor[2] * oldu[0] - cor[0] * oldu[2],
                         cor[0] * oldu[1] - cor[1] * oldu[0]};
        const double uu = dot3(oldu, oldu);
        for (int i = 0; i < ND; ++i)
          l[i] += (dot3(v.phi_u[i], oldu) + P->dt * density_scaling * dot3(grav, v.phi_u[i]) -
                   P->dt * (v.div_u[i] * 0.5 * uu + dot3(v.phi_u[i], wxu)) - P->dt * 2 * dot3(v.phi_u[i], cxu)) *
                  JxW; /* :786-804 */
      }
      orc_distribute_matrix(cs, ND, L, l, idx, A, rhs, use_omp);
    }
  }
}

/* local_assemble_nse_preconditioner (boussineq_model_FEEC.tpp:509-572): JxW multiplies ONLY the phi_p phi_p term
 * (operator precedence, quirk Q5); the curl-curl and sign terms are summed unweighted over the 8 points. */
void orc_feec_assemble_nse_preconditioner(const orc_params* P, int64_t n_cells, int nq, const double* tw,
                                          const double* tc, const double* tu, const double* td, const double* geom,
                                          const double* sign, const int32_t* l2g, const orc_constraints* cs,
                                          orc_csr* A, int use_omp) {
  const int gs = nq * 23;
  memset(A->val, 0, sizeof(double) * (size_t)A->rowptr[A->n_rows]);
#pragma omp parallel if (use_omp)
  {
    double L[ND * ND];
    feec_views v;
#pragma omp for schedule(dynamic, 32)
    for (int64_t c = 0; c < n_cells; ++c) {
      const double* g = geom + (size_t)c * gs;
      memset(L, 0, sizeof(L));
      for (int q = 0; q < nq; ++q) {
        feec_point(nq, g, q, tw, tc, tu, td, sign + (size_t)c * ND, &v);
        for (int i = 0; i < ND; ++i)
          for (int j = 0; j < ND; ++j) {
            const double uw = dot3(v.phi_u[i], v.phi_w[j]);
            const double wu = dot3(v.phi_w[i], v.phi_u[j]);
            L[i * ND + j] += P->dt * P->inv_re * dot3(v.curl_w[i], v.curl_w[j]) +
                             +(fabs(uw) > 1.0e-9 ? -2 * (signbit(uw) ? 1 - 0.5 : 0 - 0.5) : 0.0) +
                             (fabs(wu) > 1.0e-9 ? -2 * (signbit(wu) ? 1 - 0.5 : 0 - 0.5) : 0.0) +
                             v.phi_p[i] * v.phi_p[j] * g[q];
          }
      }
      orc_distribute_matrix(cs, ND, L, NULL, l2g + (size_t)c * ND, A, NULL, use_omp);
    }
  }
}

/* FEEC local_assemble_temperature_rhs (boussineq_model_FEEC.tpp:1008-1087): identical to the classic one except
 * that the advecting velocity is the Raviart-Thomas field (extractor at component dim, :1021), evaluated with
 * get_function_values, i.e. WITHOUT the face sign. */
void orc_feec_assemble_temperature_rhs(const orc_params* P, int64_t n_cells, int nd, int nq, const double* phi,
                                       const double* dphi, const double* tu, const double* geom, const int32_t* l2g,
                                       const int32_t* l2g_nse, const double* old_temp, const double* nse_solution,
                                       const orc_constraints* cs, double* rhs, int64_t n_rhs, int use_omp) {
  const int gs = nq * 23;
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
      const int32_t* idn = l2g_nse + (size_t)c * ND;
      memset(Lbc, 0, sizeof(double) * nd * nd);
      memset(l, 0, sizeof(l));
      for (int q = 0; q < nq; ++q) {
        double oldT = 0, gT[3] = {0, 0, 0}, u[3] = {0, 0, 0};
        for (int k = 0; k < nd; ++k) {
          for (int d = 0; d < 3; ++d) {
            double s = 0;
            for (int e = 0; e < 3; ++e) s += g[nq * (1 + e * 3 + d) + q] * dphi[((size_t)q * nd + k) * 3 + e];
            gr[k][d] = s;
          }
          ph[k] = phi[(size_t)q * nd + k];
          double T = old_temp[idx[k]];
          oldT += T * ph[k];
          for (int d = 0; d < 3; ++d) gT[d] += T * gr[k][d];
        }
        const double det = g[nq * 22 + q];
        for (int k = 0; k < NU; ++k) {
          const double U = nse_solution[idn[NW + k]];
          const double* phh = tu + ((size_t)q * NU + k) * 3;
          for (int d = 0; d < 3; ++d) {
            double s = 0;
            for (int e = 0; e < 3; ++e) s += g[nq * (13 + d * 3 + e) + q] * phh[e];
            u[d] += U * s / det;
          }
        }
        double ugT = dot3(u, gT);
        for (int i = 0; i < nd; ++i) {
          l[i] += (ph[i] * oldT - tau * ph[i] * ugT - tau * 0.0 * ph[i]) * g[q];
          int32_t li = cs->line_of_dof[idx[i]];
          if (li >= 0 && cs->inhom[li] != 0.0)
            for (int j = 0; j < nd; ++j)
              Lbc[j * nd + i] += (ph[i] * ph[j] + tau * P->inv_pe * dot3(gr[i], gr[j])) * g[q];
        }
      }
      orc_distribute_vector_bc(cs, nd, l, Lbc, idx, rhs, use_omp);
    }
    free(Lbc);
  }
}
