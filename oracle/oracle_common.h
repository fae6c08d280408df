/* oracle_common.h -- shared declarations of the CPU oracle (TEST INFRASTRUCTURE ONLY; see boussinesq_oracle.c). */
#ifndef ORACLE_COMMON_H
#define ORACLE_COMMON_H
#include <stdint.h>

#define ORC_MAXD 96 /* max dofs per cell handled by the stack buffers (classic 3D: 89) */

typedef struct {
  int32_t dim;
  int32_t cuboid;        /* parameters.cuboid_geometry */
  int32_t nse_interval;  /* parameters.NSE_solver_interval */
  int32_t pad;
  double dt;             /* parameters.time_step */
  double inv_re;         /* 1/Re, boussinesq_model.tpp:564-568 */
  double inv_pe;         /* 1/Pe, :760-764 */
  double beta;           /* expansion_coefficient */
  double T_ref;          /* reference_quantities.temperature_ref */
  double g_scale;        /* L/U^2, :640-643 */
  double g_const;        /* physical_constants.gravity_constant */
  double cor_scale;      /* L/U, :615-621 */
  double omega;          /* physical_constants.omega */
} orc_params;

typedef struct {
  int64_t n_dofs;
  const int32_t* line_of_dof; /* [n_dofs] -> line or -1 */
  const int32_t* line_ptr;
  const int32_t* entry_dof;
  const double* entry_w;
  const double* inhom;
} orc_constraints;

typedef struct {
  int64_t n_rows;
  const int64_t* rowptr;
  const int32_t* col;
  double* val;
} orc_csr;


void orc_csr_add(const orc_csr* A, int64_t r, int32_t c, double v, int atomic);
void orc_vec_add(double* b, int64_t i, double v, int atomic);
void orc_distribute_matrix(const orc_constraints* cs, int n, const double* L, const double* l, const int32_t* idx,
                           const orc_csr* A, double* b, int atomic);
void orc_distribute_vector_bc(const orc_constraints* cs, int n, const double* l, const double* Lbc, const int32_t* idx,
                              double* b, int atomic);
void orc_gravity(const orc_params* P, const double* x, double* g);
#endif
