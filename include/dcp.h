/* dcp.h -- C ABI of the B200 device library (libdcp.so) for 3D-DyCorePlanet's data-parallel hot path:
 * per-cell finite-element assembly of the buoyancy-Boussinesq system and the FP64 CSR SpMVs of its
 * Krylov solves.  Plain pointers and sizes only; no C++ / torch types; every call returns an int status
 * (0 = DCP_OK) and leaves a message retrievable with dcp_last_error().
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *
 *   dcp_model_create            the "once per mesh" part of Standard::BoussinesqModel::setup_dofs
 *                               (include/core/boussinesq_model.tpp:184-412): DoF maps, constraint lines
 *                               and sparsity patterns are taken from the caller (deal.II keeps ownership
 *                               of mesh / DoFHandler / AffineConstraints) and uploaded into SoA device
 *                               buffers together with what FEValues::reinit would deliver per cell
 *                               (Scratch objects, include/core/boussineq_model_assembly.tpp:18-171).
 *   dcp_assemble_nse_system     the statement group `nse_matrix = 0; nse_rhs = 0; WorkStream::run(
 *                               local_assemble_nse_system, copy_local_to_global_nse_system);
 *                               compress(add)` of assemble_nse_system (boussinesq_model.tpp:691-740,
 *                               worker :550-673, copier :677-687).  FEEC: boussineq_model_FEEC.tpp:826-875.
 *   dcp_assemble_nse_preconditioner   assemble_nse_preconditioner (:479-514, worker :421-464, copier
 *                               :468-476) + the Jacobi set-up of build_nse_preconditioner (:520-542).
 *   dcp_assemble_temperature_matrix   assemble_temperature_matrix (:821-864, worker :748-800, copier :804-817).
 *   dcp_assemble_temperature_rhs      assemble_temperature_rhs (:966-1020): temperature_matrix = M + dt/n K,
 *                               Jacobi set-up, worker :873-952, copier :955-964 (matrix_for_bc overload).
 *   dcp_vmult / dcp_vmult_add   LA::SparseMatrix::vmult / vmult_add on one block (Epetra Multiply), call
 *                               sites include/linear_algebra/schur_complement.hpp:147-149,266-274,
 *                               approximate_schur_complement.hpp:139-141, shifted_schur_complement.hpp:
 *                               159-170, nested_schur_complement.hpp:177-179,
 *                               block_schur_preconditioner.hpp:55,131,144, boussinesq_model.tpp:1317,1392.
 *   dcp_block_vmult             LA::BlockSparseMatrix::vmult on nse_matrix as called by SolverFGMRES /
 *                               SolverGMRES (boussinesq_model.tpp:1196,1225; boussineq_model_FEEC.tpp:1372-1401).
 *   dcp_jacobi_vmult            LA::PreconditionJacobi::vmult (one sweep, omega = 1) on
 *                               Mu_plus_A_preconditioner / Mp_preconditioner / T_preconditioner
 *                               (boussinesq_model.tpp:531-539, 982-983).
 *
 * Threading: one context per GPU, calls from one host thread (the rank's main thread, where the
 * reference calls WorkStream::run).  Not re-entrant per context.
 * There is NO CPU fallback: every entry point fails with DCP_ERR_CUDA if no device is present.
 *
 * Deferred error reports: with DCP_DEVICE operands the assemblers do not synchronise, so a scatter target that is missing
 * from the sparsity pattern (DCP_ERR_PATTERN) is counted on the device and reported by the next synchronising entry point
 * -- dcp_ctx_synchronize, dcp_vec_dot, dcp_matrix_download / dcp_vector_download -- instead of by the assembly call itself
 * (with DCP_HOST operands the call synchronises and reports directly).
 */
#ifndef DCP_H
#define DCP_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
  DCP_OK = 0,
  DCP_ERR_CUDA = 1,     /* CUDA runtime error or no device */
  DCP_ERR_ARG = 2,      /* invalid argument */
  DCP_ERR_PATTERN = 3,  /* a scatter target is not in the sparsity pattern */
  DCP_ERR_STATE = 4,    /* call order (e.g. rhs before matrices) */
  DCP_ERR_NO_CONVERGENCE = 5  /* dcp_cg_solve: SolverControl::NoConvergence (last step / residual are reported) */
};

enum { DCP_HOST = 0, DCP_DEVICE = 1 }; /* where a caller-provided vector lives */
enum { DCP_FAMILY_CLASSIC = 0, DCP_FAMILY_FEEC = 1 };

/* which matrix of the model */
enum {
  DCP_MAT_NSE = 0,        /* nse_matrix                      (blocks) */
  DCP_MAT_NSE_PRECOND = 1,/* nse_preconditioner_matrix       (blocks) */
  DCP_MAT_TEMP_MASS = 2,  /* temperature_mass_matrix         (single block 0,0) */
  DCP_MAT_TEMP_STIFF = 3, /* temperature_stiffness_matrix */
  DCP_MAT_TEMP = 4        /* temperature_matrix = M + dt/n K */
};
/* row classes of a row-distributed operator (dcp_vmult_rows): all owned rows, the rows that read owned columns only,
 * the rows that read at least one ghost column */
enum { DCP_ROWS_ALL = 0, DCP_ROWS_INTERIOR = 1, DCP_ROWS_GHOSTED = 2 };
/* which vector of the model */
enum { DCP_VEC_NSE_RHS = 0, DCP_VEC_TEMP_RHS = 1 };
/* assembly strategy (dcp_model_set_strategy): all give the same matrix up to summation order */
enum {
  DCP_STRATEGY_SEARCH = 0,    /* every cell: general AffineConstraints scatter, column positions by binary search */
  DCP_STRATEGY_POSITIONS = 1, /* unconstrained cells scatter from registers through a precomputed position table
                                 (no search, no local matrix in shared memory); constrained cells use SEARCH */
  DCP_STRATEGY_OWNER = 2,     /* row-owner tiles: a CTA owns a contiguous row range, accumulates in shared memory and
                                 writes every CSR value of an unconstrained row exactly once (no atomics);
                                 contributions of constrained dofs are added afterwards by SEARCH on the constrained
                                 cells */
  DCP_STRATEGY_STAGED = 3     /* write-once (classic 3-D family, every cell in the position plan): the cell blocks go
                                 through a cell-major staging buffer, a row-owner gather pass writes every CSR value
                                 once -- no atomics on the matrix, no zero-fill.  Default when the model qualifies;
                                 matrices without a staged path (e.g. the preconditioner until it has one) use
                                 POSITIONS */
};

typedef struct dcp_ctx dcp_ctx;
typedef struct dcp_model dcp_model;

/* The pointwise coefficients the quadrature loops consume (boussinesq_model.tpp:564-568, 615-621, 640-650,
 * 760-764; core_model_data.cc:88-94; core_model_data.tpp:86-118). */
typedef struct {
  int32_t dim;
  int32_t cuboid;       /* parameters.cuboid_geometry: vertical gravity + Coriolis */
  int32_t nse_interval; /* parameters.NSE_solver_interval */
  int32_t pad;
  double dt;            /* parameters.time_step */
  double inv_re;        /* 1 / Reynolds */
  double inv_pe;        /* 1 / Peclet */
  double beta;          /* expansion coefficient */
  double T_ref;         /* reference temperature */
  double g_scale;       /* L / U^2 */
  double g_const;       /* gravity constant */
  double cor_scale;     /* L / U */
  double omega;         /* rotation rate */
} dcp_params;

/* AffineConstraints as a CSR of lines: dof -> sum_k entry_w[k] * dof(entry_dof[k]) + inhom */
typedef struct {
  int64_t n_dofs;
  int64_t n_lines;
  const int32_t* line_dof;  /* [n_lines], ascending */
  const int32_t* line_ptr;  /* [n_lines+1] */
  const int32_t* entry_dof;
  const double* entry_w;
  const double* inhom;      /* [n_lines] */
} dcp_constraints_desc;

/* one CSR block; rowptr == NULL means "empty block" */
typedef struct {
  int64_t n_rows, n_cols;
  const int64_t* rowptr; /* [n_rows+1] */
  const int32_t* col;    /* block-local, ascending within a row */
} dcp_csr_desc;

#define DCP_MAX_BLOCKS 3

/* Everything that is uploaded once per mesh.  All pointers are HOST pointers; the library copies. */
typedef struct {
  int32_t dim;
  int32_t family;
  int64_t n_cells;

  /* Navier-Stokes FE space: cell -> global dof map in the block-concatenated numbering
   * (after DoFRenumbering::component_wise, boussinesq_model.tpp:204) */
  int32_t nse_n_local;               /* dofs per cell (classic 3-D: 89) */
  int32_t nse_n_blocks;              /* classic 2 (u,p), FEEC 3 (w,u,p) */
  int64_t nse_block_size[DCP_MAX_BLOCKS];
  const int32_t* nse_l2g;            /* [n_cells][nse_n_local] */
  const int32_t* nse_local_field;    /* [nse_n_local] vector component of each cell dof */
  const int32_t* nse_local_base;     /* [nse_n_local] index within its scalar base element */
  dcp_constraints_desc nse_cs;

  /* temperature FE space */
  int32_t temp_n_local;
  int32_t pad0;
  const int32_t* temp_l2g;
  dcp_constraints_desc temp_cs;

  /* reference-cell tables, [nq][nd] values and [nq][nd][dim] reference gradients.
   * *_qn on the NSE rule QGauss(deg+1) (boussinesq_model.tpp:487,708), *_qt on the temperature rule
   * QGauss(Tdeg+2) (:834,990). */
  int32_t nq_nse, nq_temp;
  int32_t ndu, ndp, ndt;
  int32_t build_owner_plan; /* != 0: also build the row-owner tile plan (DCP_STRATEGY_OWNER) at create time */
  const double *phi_u_qn, *dphi_u_qn, *phi_p_qn, *phi_t_qn;
  const double *phi_u_qt, *phi_t_qt, *dphi_t_qt;

  /* mapping data per cell and rule: [n_cells][ JxW(nq) | Kinv[e][d](nq each) | xq[d](nq each) ],
   * Kinv[e][d] = d xi_e / d x_d (inverse Jacobian of the cell mapping at the quadrature point).
   * With geom_on_device != 0 (last field) these are DEVICE buffers made by dcp_geometry_create, which the model
   * adopts (it frees them in dcp_model_destroy; pass the same pointer twice when two rules coincide). */
  const double* geom_qn;
  const double* geom_qt; /* may be == geom_qn when both rules coincide */

  /* sparsity patterns as deal.II/Trilinos built them (setup_nse_matrices :79-112, setup_nse_preconditioner
   * :116-150, setup_temperature_matrices :154-180) */
  dcp_csr_desc nse_pattern[DCP_MAX_BLOCKS][DCP_MAX_BLOCKS];
  dcp_csr_desc pre_pattern[DCP_MAX_BLOCKS][DCP_MAX_BLOCKS];
  dcp_csr_desc temp_pattern;

  /* ---- FEEC family only (family == DCP_FAMILY_FEEC; ExteriorCalculus::BoussinesqModel<3>,
   * include/core/boussineq_model_FEEC.tpp).  FESystem(FE_Nedelec(0), FE_RaviartThomas(0), FE_DGQ(0)): cell dofs
   * 0..11 vorticity (lines), 12..17 velocity (faces), 18 pressure; three blocks (w,u,p).  The mapping records of
   * this family carry the Jacobian for the Piola transforms: [JxW | Kinv | xq | J[i][j] | detJ], nq entries each.
   * The Lagrange velocity/pressure tables above are unused; the temperature tables are used as in the classic family. */
  int32_t nq_pre;            /* preconditioner rule QGauss(deg+1) (:595) */
  int32_t pad2;
  const double* nse_sign;    /* [n_cells][19] face sign of Tools::get_face_sign_change_raviart_thomas (utilities.cc:20-46) */
  const double *feec_phi_w_qn, *feec_curl_w_qn, *feec_phi_u_qn; /* reference values [nq_nse][12|12|6][3] */
  const double *feec_phi_w_qp, *feec_curl_w_qp, *feec_phi_u_qp; /* ... on the preconditioner rule */
  const double* feec_phi_u_qt;                                   /* RT values on the temperature rule */
  const double* feec_div_u;                                      /* [6] reference divergence of the RT functions */
  const double* geom_qp;                                         /* mapping records on the preconditioner rule */

  /* ---- optional, for the time-step diagnostics (dcp_velocity_extrema) */
  const double* cell_vertices; /* [n_cells][2^dim][dim] vertex coordinates, deal.II (lexicographic) vertex order; may be NULL */
  int64_t n_owned_cells;       /* cells [0, n_owned_cells) are locally owned (cell->is_locally_owned()); 0 = all */
  int32_t geom_on_device;      /* != 0: geom_qn / geom_qt / geom_qp are device buffers from dcp_geometry_create */
  int32_t pad3;
} dcp_model_desc;

/* Input of the device-side mapping evaluation (what FEValues::reinit computes with update_JxW_values |
 * update_inverse_jacobians | update_quadrature_points (| update_jacobians for the Piola transforms),
 * boussineq_model_assembly.tpp:25,69-73).  MappingQ(p) as the reference constructs it (boussinesq_model.tpp:20):
 * cells with boundary lines carry n_high = (p+1)^dim support points, all others their n_low = 2^dim vertices. */
typedef struct {
  int32_t dim, nq;
  int32_t extended;              /* != 0: FEEC records [.. | J[i][j] | detJ] */
  int32_t n_low, n_high;
  int32_t pad;
  int64_t n_cells;
  const int64_t* support_ptr;    /* [n_cells+1] offsets, in points, into support_points */
  const double* support_points;  /* [..][dim], lexicographic (x fastest) within a cell */
  const double *N_low, *dN_low;  /* mapping basis at the quadrature points: [nq][n_low], [nq][n_low][dim] */
  const double *N_high, *dN_high;/* same for the high-order cells; may be NULL when no cell uses it */
  const double* weights;         /* [nq] quadrature weights */
} dcp_mapping_desc;

/* ---- context --------------------------------------------------------------------------------- */
int dcp_ctx_create(int device, dcp_ctx** out);
int dcp_ctx_destroy(dcp_ctx* ctx);
/* run all work of this context on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = the context's
 * own non-blocking stream; the legacy default stream is cudaStreamLegacy, i.e. (void*)1 */
int dcp_ctx_set_stream(dcp_ctx* ctx, void* cuda_stream);
int dcp_ctx_synchronize(dcp_ctx* ctx);
/* message of the last failing call on this thread */
const char* dcp_last_error(void);
/* number of kernel launches issued by this context since creation (for bench accounting) */
int64_t dcp_ctx_launch_count(const dcp_ctx* ctx);
/* device memory helpers for callers without a CUDA runtime of their own */
int dcp_malloc(dcp_ctx* ctx, int64_t bytes, void** out);
int dcp_free(dcp_ctx* ctx, void* p);
int dcp_memcpy_h2d(dcp_ctx* ctx, void* dst_dev, const void* src_host, int64_t bytes);
int dcp_memcpy_d2h(dcp_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes);
/* The same copies on the context's own copy stream, so that vectors travel while kernels run (host memory should be
 * pinned).  Ordering against the work of the context's stream is explicit: dcp_copy_fence(DCP_COMPUTE_WAITS_FOR_COPIES)
 * makes everything enqueued on the context's stream afterwards wait for the copies enqueued so far (an uploaded vector is
 * used), dcp_copy_fence(DCP_COPIES_WAIT_FOR_COMPUTE) makes later copies wait for the kernels enqueued so far (a result
 * is downloaded); dcp_copy_synchronize blocks the host until the copy stream is idle.  A deal.II host uses these to move
 * LA::MPI::Vector data of the next / previous operator call behind the current one. */
enum { DCP_COMPUTE_WAITS_FOR_COPIES = 0, DCP_COPIES_WAIT_FOR_COMPUTE = 1 };
int dcp_memcpy_h2d_async(dcp_ctx* ctx, void* dst_dev, const void* src_host, int64_t bytes);
int dcp_memcpy_d2h_async(dcp_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes);
int dcp_copy_fence(dcp_ctx* ctx, int direction);
int dcp_copy_synchronize(dcp_ctx* ctx);

/* ---- mapping data on the device ------------------------------------------------------- */
/* Evaluates the records of one quadrature rule for all cells into a new device buffer (*geom_dev).  Hand it to
 * dcp_model_create through geom_q* with geom_on_device = 1 (the model then owns it) or release it with dcp_free. */
int dcp_geometry_create(dcp_ctx* ctx, const dcp_mapping_desc* desc, double** geom_dev);

/* ---- model (one per mesh) -------------------------------------------------------------------- */
int dcp_model_create(dcp_ctx* ctx, const dcp_model_desc* desc, dcp_model** out);
int dcp_model_destroy(dcp_model* m);
int dcp_model_set_strategy(dcp_model* m, int strategy);
/* the strategy in use (DCP_STRATEGY_STAGED by default when the model qualifies, else DCP_STRATEGY_POSITIONS); < 0: bad handle */
int dcp_model_get_strategy(const dcp_model* m);
/* Multi-GPU (one process per GPU, cells partitioned along the p4est curve as the reference does over MPI,
 * include/core/boussinesq_model.tpp:241-252, 491-494): the rank's local numbering puts the dofs it owns first
 * inside every block, ghost dofs after them.  After this call the operators (vmult, vmult_add, block_vmult,
 * jacobi_vmult) compute only the owned rows of each block -- the ghost entries of `src` must have been refreshed
 * by the caller's halo exchange (Epetra_Import inside LA::SparseMatrix::vmult), the ghost rows of `dst` are left
 * untouched.  Assembly still integrates every local cell (owned + one ghost layer), which makes each owned row
 * complete without the reference's compress(VectorOperation::add) (:513, 736-737, 859-861, 1017). */
int dcp_model_set_owned(dcp_model* m, const int64_t* nse_owned_per_block, int64_t temp_owned);
/* halo pack / unpack: dst[i] = src[idx[i]]  and  dst[idx[i]] = src[i]  (device pointers) */
int dcp_gather_f64(dcp_ctx* ctx, int64_t n, const int32_t* idx_dev, const double* src_dev, double* dst_dev);
int dcp_scatter_f64(dcp_ctx* ctx, int64_t n, const int32_t* idx_dev, const double* src_dev, double* dst_dev);

/* ---- assembly (one call per reference assemble_* member) ------------------------------------- */
/* old_nse: [n_u+n_p(+..)] ghosted old_nse_solution, old_temp: [n_T] old_temperature_solution */
int dcp_assemble_nse_system(dcp_model* m, const dcp_params* p, const double* old_nse, const double* old_temp, int mem);
int dcp_assemble_nse_preconditioner(dcp_model* m, const dcp_params* p);
int dcp_assemble_temperature_matrix(dcp_model* m, const dcp_params* p);
/* nse_solution: the NEW velocity/pressure vector (reference advects with nse_solution, :904-905) */
int dcp_assemble_temperature_rhs(dcp_model* m, const dcp_params* p, const double* old_temp, const double* nse_solution, int mem);

/* ---- results --------------------------------------------------------------------------------- */
/* nnz / shape of block (bi,bj) of matrix `which`; nnz = 0 for an empty block */
int dcp_matrix_info(const dcp_model* m, int which, int bi, int bj, int64_t* n_rows, int64_t* n_cols, int64_t* nnz);
/* device pointer to the values of that block (owned by the model) */
int dcp_matrix_values_device(dcp_model* m, int which, int bi, int bj, double** out);
int dcp_matrix_download(dcp_model* m, int which, int bi, int bj, double* host_values);
int dcp_matrix_upload(dcp_model* m, int which, int bi, int bj, const double* host_values);
int dcp_vector_device(dcp_model* m, int which, double** out, int64_t* n);
int dcp_vector_download(dcp_model* m, int which, double* host);

/* ---- operators (deal.II vmult concept) ------------------------------------------------------- */
/* dst = A(bi,bj) * src ;  dst += A(bi,bj) * src */
int dcp_vmult(dcp_model* m, int which, int bi, int bj, double* dst, const double* src, int mem);
int dcp_vmult_add(dcp_model* m, int which, int bi, int bj, double* dst, const double* src, int mem);
/* dst = A * src over all blocks; vectors are block-concatenated */
int dcp_block_vmult(dcp_model* m, int which, double* dst, const double* src, int mem);
/* The same products restricted to a row class (device pointers only; needs dcp_model_set_owned).  Trilinos overlaps
 * the Epetra_Import of the ghost entries with the local part of Epetra_CrsMatrix::Multiply; here the caller runs
 * DCP_ROWS_INTERIOR while its halo exchange is in flight on another stream and DCP_ROWS_GHOSTED after it (a row of a
 * block row is GHOSTED when any of its blocks has a ghost column in that row, so the two classes partition the rows). */
int dcp_vmult_rows(dcp_model* m, int which, int bi, int bj, double* dst, const double* src, int rows);
int dcp_block_vmult_rows(dcp_model* m, int which, double* dst, const double* src, int rows);
/* dst = diag(A(bi,bi))^-1 * src  (Jacobi, one sweep, omega 1).  Diagonals are refreshed by the assemble calls. */
int dcp_jacobi_vmult(dcp_model* m, int which, int bi, double* dst, const double* src, int mem);

/* ---- passes next to the solves ---------------------------------------------------------
 * get_maximal_velocity and get_cfl_number in one sweep over the locally owned cells (boussinesq_model.tpp:1023-1061
 * and :1064-1098; FEEC boussineq_model_FEEC.tpp:1158-1240): result_host[0] = max_q |u(x_q)| over the
 * QIterated(QTrapez, degree) points, result_host[1] = max_cell max(1e-10, max_q |u|) / cell->diameter().
 * The values are rank-local; with several ranks the caller takes the maximum (Utilities::MPI::max, :1048, 1093).
 * Without desc.cell_vertices result_host[1] is set to -1 (classic) and the FEEC family fails with DCP_ERR_STATE. */
int dcp_velocity_extrema(dcp_model* m, const double* nse_solution, int mem, double* result_host);
/* AffineConstraints::distribute on a solution vector (nse_constraints.distribute :1233, temperature_constraints
 * .distribute :1442): x[line] = sum_k w_k x[master_k] + inhomogeneity.  space: 0 = NSE, 1 = temperature. */
int dcp_constraints_distribute(dcp_model* m, int space, double* x, int mem);

/* ---- ILU(0) -------------------------------------------------------------------------------
 * LA::PreconditionILU = Ifpack ILU with deal.II's defaults (ilu_fill 0, ilu_atol 0, ilu_rtol 1, overlap 0) of the
 * diagonal block (bi,bi) of matrix `which`: the inner preconditioner of the classic Schur-complement solve
 * (`inner_schur_preconditioner->initialize(nse_matrix.block(0,0), data)`, boussinesq_model.tpp:1265-1275;
 * preconditioner.h:36-42) and the PreconditionILU inside ApproximateSchurComplement
 * (approximate_schur_complement.hpp:118-141).  dcp_ilu_create analyses the pattern (dependency levels) and
 * factorises the current values; dcp_ilu_refactor repeats the numeric part after a re-assembly; dcp_ilu_vmult is
 * PreconditionILU::vmult (forward + backward substitution).  After dcp_model_set_owned only the rank-local square
 * block is factorised, as Ifpack does with overlap 0.  dcp_model_destroy also destroys the handles still alive
 * (do not use or destroy them afterwards). */
typedef struct dcp_ilu dcp_ilu;
int dcp_ilu_create(dcp_model* m, int which, int bi, dcp_ilu** out);
int dcp_ilu_refactor(dcp_ilu* p);
int dcp_ilu_vmult(dcp_ilu* p, double* dst, const double* src, int mem);
/* number of dependency levels of the two substitutions (= kernel launches per vmult) */
int dcp_ilu_levels(const dcp_ilu* p, int64_t* n_lower, int64_t* n_upper);
int dcp_ilu_destroy(dcp_ilu* p);

/* ---- device-resident vector algebra for the Krylov solvers around the SpMVs (all pointers DEVICE) -----------
 * Trilinos vector ops inside deal.II SolverCG / SolverGMRES / SolverFGMRES (boussinesq_model.tpp:1165,1191-1199,
 * 1426-1440; inverse_matrix.hpp:99-113): operator* / l2_norm (dot, bit-reproducible two-stage tree; the scalar
 * is returned to the host like the reference's MPI_Allreduce), add (axpy), sadd (y = s*y + a*x), scale, equ. */
int dcp_vec_dot(dcp_ctx* ctx, int64_t n, const double* x_dev, const double* y_dev, double* result_host);
int dcp_vec_axpy(dcp_ctx* ctx, int64_t n, double a, const double* x_dev, double* y_dev);
int dcp_vec_sadd(dcp_ctx* ctx, int64_t n, double s, double a, const double* x_dev, double* y_dev);
int dcp_vec_scale(dcp_ctx* ctx, int64_t n, double a, double* y_dev);
int dcp_vec_copy(dcp_ctx* ctx, int64_t n, const double* x_dev, double* y_dev);
/* Arnoldi orthogonalisation step of SolverGMRES / SolverFGMRES (modified Gram-Schmidt, deal.II's default; e.g. the outer
 * solve boussinesq_model.tpp:1191-1199): for i < k: h[i] = w . v_i, w -= h[i] v_i; then h[k] = w . w.  The inner products
 * stay in device memory between the updates; the k + 1 scalars come back with ONE synchronisation instead of k + 1.
 * v_dev: host array of k device pointers; k <= DCP_MGS_MAX.  Same kernels as dcp_vec_dot / dcp_vec_axpy: bit-identical
 * to the loop that calls them. */
enum { DCP_MGS_MAX = 256 };
int dcp_vec_mgs(dcp_ctx* ctx, int64_t n, int k, const double* const* v_dev, double* w_dev, double* h_host);
/* SolverCG<Vector>::solve(A, x, b, P) with SolverControl(max_steps, tol), resident on the device: A = block (bi, bj) of
 * matrix `which`, P = identity, the Jacobi diagonal of block (bp, bp) of matrix `which_p`, or an ILU(0) handle.  alpha, beta
 * and the residual norm stay in device memory; the host reads the convergence flag once per `check_every` iterations
 * (<= 0: 8) and iterations enqueued behind the converged one are no-ops, so x, *last_step and *last_residual are those
 * of the reference's loop (inner products: the tree of dcp_vec_dot).  x_dev holds the start vector.  Returns
 * DCP_ERR_NO_CONVERGENCE after max_steps steps (x_dev = last iterate, like the exception the reference catches).
 * Replaces the SolverCG loops of LinearAlgebra::InverseMatrix::vmult (include/linear_algebra/inverse_matrix.hpp:90-121),
 * ApproximateInverseMatrix::vmult (approximate_inverse.hpp:97-128) and solve_temperature
 * (include/core/boussinesq_model.tpp:1426-1440).  Single-rank models (row-distributed: DCP_ERR_STATE). */
enum { DCP_PRECOND_IDENTITY = 0, DCP_PRECOND_JACOBI = 1, DCP_PRECOND_ILU = 2 };
int dcp_cg_solve(dcp_model* m, int which, int bi, int bj, int precond, int which_p, int bp, dcp_ilu* ilu, double* x_dev,
                 const double* b_dev, double tol, int64_t max_steps, int check_every, int64_t* last_step, double* last_residual);
/* y = value (deal.II `dst = 0`: an assignment, so NaN / Inf in an uninitialised destination do not survive) */
int dcp_vec_fill(dcp_ctx* ctx, int64_t n, double value, double* y_dev);
/* y += a in every entry (Vector::add(a): the zero-mean correction of the FEEC pressure, nested_schur_complement.hpp:180-182) */
int dcp_vec_shift(dcp_ctx* ctx, int64_t n, double a, double* y_dev);

/* ---- multi-GPU data plane (one process per GPU, NCCL over NVLink; csrc/device/halo.cu) --------------------------
 * Replaces what Trilinos / deal.II hide behind MPI in the reference: the Epetra_Import of off-rank source entries
 * inside every LA::SparseMatrix::vmult (include/linear_algebra/schur_complement.hpp:143-150,
 * block_schur_preconditioner.hpp:55), the ghost refresh `nse_solution = distributed_nse_solution`
 * (include/core/boussinesq_model.tpp:1241, 1444), the MPI_Allreduce of l2_norm / operator* (:1165, 1427) and
 * Utilities::MPI::max (:1050, 1094, 1467).  NCCL is bound at run time (libnccl.so.2); single-GPU use never needs it.
 *
 * dcp_comm_unique_id: rank 0 fills 128 bytes (ncclUniqueId) and broadcasts them by whatever means the host has
 * (MPI_Bcast in a deal.II build); every rank then calls dcp_comm_create.  dcp_comm_adopt wraps an ncclComm_t the host
 * already owns (not destroyed with the handle). */
#define DCP_UNIQUE_ID_BYTES 128
typedef struct dcp_comm dcp_comm;
typedef struct dcp_halo dcp_halo;
int dcp_comm_unique_id(void* id_out);
int dcp_comm_create(dcp_ctx* ctx, const void* id, int rank, int n_ranks, dcp_comm** out);
int dcp_comm_adopt(dcp_ctx* ctx, void* nccl_comm, int rank, int n_ranks, dcp_comm** out);
int dcp_comm_info(const dcp_comm* c, int* rank, int* n_ranks);
int dcp_comm_destroy(dcp_comm* c);
/* Ghost plan of one local vector layout (n_local entries, owned entries first inside each block, ghosts after them):
 * send_idx = local indices of the owned entries other ranks read, grouped by destination rank (send_counts[n_ranks]);
 * recv_idx = local ghost slots grouped by owning rank (recv_counts[n_ranks]), in the order the owner sends them.
 * All arrays are host arrays and are copied. */
int dcp_halo_create(dcp_comm* c, int64_t n_local, const int32_t* send_idx, const int64_t* send_counts,
                    const int32_t* recv_idx, const int64_t* recv_counts, dcp_halo** out);
int dcp_halo_destroy(dcp_halo* h);
/* refresh the ghost entries of the device vector x in place (pack kernel -> grouped ncclSend/ncclRecv -> unpack
 * kernel on the context's stream; no host synchronisation) */
int dcp_halo_exchange(dcp_halo* h, double* x_dev);
/* dst(owned rows) = A * src for the row-distributed matrix `which` (all blocks; a single-block matrix is its own block
 * matrix): ghost refresh of src, then the products.  overlap != 0: the exchange runs on the communicator's stream while
 * the rows that read owned columns only are computed; the rows with ghost columns follow (needs dcp_model_set_owned).
 * Device pointers; src's ghost slots are overwritten. */
int dcp_halo_block_vmult(dcp_model* m, int which, dcp_halo* h, double* dst_dev, double* src_dev, int overlap);
/* sum over the given index ranges (the owned entries of each block) of x[i] * y[i], all-reduced over the ranks: the
 * Krylov solvers' inner product.  At most 16 ranges.  The scalar is returned to the host. */
int dcp_vec_dot_allreduce(dcp_comm* c, int n_ranges, const int64_t* range_begin, const int64_t* range_end,
                          const double* x_dev, const double* y_dev, double* result_host);
/* Utilities::MPI::max over the ranks of up to 8 host scalars, in place */
int dcp_allreduce_max(dcp_comm* c, int n, double* values_host);

#ifdef __cplusplus
}
#endif
#endif
