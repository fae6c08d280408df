// Internal declarations of the device library (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../../include/dcp.h"

#define DCP_MAXB DCP_MAX_BLOCKS

void dcp_set_error(const std::string& s);

#define DCP_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      dcp_set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                    std::to_string(__LINE__) + ")");                                           \
      return DCP_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define DCP_TRY(call)            \
  do {                           \
    int rc__ = (call);           \
    if (rc__ != DCP_OK) return rc__; \
  } while (0)

struct dcp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int64_t launches = 0;
  int sm_count = 148;
  int* d_err = nullptr;  // [4] device-side error counters (0: missing pattern entries)
  int* h_err = nullptr;  // pinned mirror
  // staging buffers for DCP_HOST vectors
  double* stage[3] = {nullptr, nullptr, nullptr};
  int64_t stage_cap[3] = {0, 0, 0};
  double* dot_scratch = nullptr;  // partial sums of dcp_vec_dot
  double* dot_host = nullptr;     // pinned result
  double* mgs_scalars = nullptr;  // dcp_vec_mgs: Hessenberg column + partial sums (device), pinned mirror
  double* mgs_host = nullptr;
  cudaStream_t copy_stream = nullptr;   // dcp_memcpy_*_async (created on first use)
  cudaEvent_t copy_event = nullptr;
};

// constraint lines on the device
struct DevCs {
  int64_t n_dofs = 0, n_lines = 0;
  int32_t* line_of_dof = nullptr;
  int32_t* line_ptr = nullptr;
  int32_t* entry_dof = nullptr;
  double* entry_w = nullptr;
  double* inhom = nullptr;
};
struct CsView {
  const int32_t* line_of_dof;
  const int32_t* line_ptr;
  const int32_t* entry_dof;
  const double* entry_w;
  const double* inhom;
};

struct DevCsr {
  int64_t n_rows = 0, n_cols = 0, nnz = 0;
  int64_t* rowptr = nullptr;
  int32_t* col = nullptr;
  double* val = nullptr;
  bool owns_pattern = true;  // mass/stiffness/temperature matrices share one pattern
  int lanes = 32;            // lanes per row chosen for the SpMV kernel
  mutable int triple = -1;   // SpMV: most groups of three consecutive rows share one pattern (-1: not tested yet)
  mutable unsigned char* triple_same = nullptr;   // ... per group: 1 = shared pattern
};

struct BlockMat {
  int nb = 1;
  int64_t start[DCP_MAXB + 1] = {0, 0, 0, 0};
  DevCsr blk[DCP_MAXB][DCP_MAXB];
  double* diag_inv[DCP_MAXB] = {nullptr, nullptr, nullptr};  // Jacobi: 1/diag of the diagonal blocks
  int32_t* diag_off[DCP_MAXB] = {nullptr, nullptr, nullptr};  // offset of the diagonal entry inside its row (cached)
  int64_t owned[DCP_MAXB] = {-1, -1, -1};  // rows of each block owned by this rank (-1: all)
  // multi-GPU overlap: per block row, the owned rows that read ghost columns (flag per row, compact list)
  uint8_t* ghost_flag[DCP_MAXB] = {nullptr, nullptr, nullptr};
  int32_t* ghost_list[DCP_MAXB] = {nullptr, nullptr, nullptr};
  int64_t n_ghost_rows[DCP_MAXB] = {0, 0, 0};
  bool owns_ghost_rows = true;
};

// by-value kernel argument describing a block matrix for scatter
struct BlockView {
  int nb;
  long long start[DCP_MAXB + 1];
  const long long* rowptr[DCP_MAXB][DCP_MAXB];
  const int* col[DCP_MAXB][DCP_MAXB];
  double* val[DCP_MAXB][DCP_MAXB];
};

inline BlockView make_view(const BlockMat& M) {
  BlockView v;
  v.nb = M.nb;
  for (int i = 0; i <= DCP_MAXB; ++i) v.start[i] = M.start[i < M.nb + 1 ? i : M.nb];
  for (int i = 0; i < DCP_MAXB; ++i)
    for (int j = 0; j < DCP_MAXB; ++j) {
      v.rowptr[i][j] = (const long long*)M.blk[i][j].rowptr;
      v.col[i][j] = M.blk[i][j].col;
      v.val[i][j] = M.blk[i][j].val;
    }
  return v;
}

inline CsView make_view(const DevCs& c) { return CsView{c.line_of_dof, c.line_ptr, c.entry_dof, c.entry_w, c.inhom}; }

struct OwnerPlan;  // row-owner tiles (assemble_th_owner.cu)
// position tables of the cells without constrained dofs (built in assemble_th_fast.cu)
struct FastPlan {
  int64_t n_fast = 0, n_general = 0;
  int32_t* fast_cells = nullptr;
  int32_t* general_cells = nullptr;
  uint16_t* pos = nullptr;  // [n_fast][NE*NE], NE = NU + NP
  int ne = 0;
};

// (chunk, node) items of the write-once gather pass (assemble_th_stage.cu)
struct GatherPlan {
  int64_t chunk = 0, n_chunks = 0, n_cells = 0;
  int32_t *v_g0 = nullptr, *p_g0 = nullptr;            // per item: first dof of the velocity node / pressure row
  uint32_t *v_incptr = nullptr, *p_incptr = nullptr;   // per item: range in *_inc
  uint32_t *v_inc = nullptr, *p_inc = nullptr;         // (cell index inside the chunk << 5) | local node
  uint8_t *v_flag = nullptr, *p_flag = nullptr;        // bit 0: first chunk that touches the node (store, else accumulate)
  double *v_meta = nullptr, *p_meta = nullptr;         // per incidence, in incidence order: header + row positions (18 doubles)
  uint32_t* slots = nullptr;                           // [plan cell][36]: staging slot of every local node = index of its incidence in the chunk
  std::vector<int64_t> v_chunk_ptr, p_chunk_ptr;       // host: item range of every chunk
  double* staging = nullptr;                           // node-major: [27 chunk][304] velocity-node rows, then [8 chunk][92] pressure-node rows
  double* dphi_lane = nullptr;                         // reference gradients in table-build order [7][3][108]
  // fused preconditioner (the system pass also stages m + nu k and the pressure mass block; the gather writes
  // nse_preconditioner_matrix, too): index of every system-plan cell in the preconditioner plan (-1: the cell is
  // assembled by the reduction kernels afterwards), and those cells as a list of preconditioner-plan indices
  long long* pre_w = nullptr;
  int32_t* pre_rest = nullptr;
  int64_t n_pre_rest = 0;
  bool has_pre = false;
  int pstr = 4;                                        // accumulator length of one preconditioner row
};

// masked position tables for the DMMA path (assemble_th_mma.cu): every node-blocked cell, constrained or not
struct MaskedPlan {
  int64_t n = 0, n_other = 0, n_wide = 0;
  int32_t* cells = nullptr;
  int32_t* other_cells = nullptr;  // cells that failed the layout check: general kernel, full mode
  uint16_t* pos = nullptr;         // [n][1228]: 35*35 offsets padded to 8 bytes
  uint8_t* nmask = nullptr;        // [n][48]: node masks, pressure flags, cell flag, wide-table index, 9-table index
  uint16_t* pos_wide = nullptr;    // [n_wide][3][27][27]
  uint16_t* pos9 = nullptr;        // [n_nnf][9][27][27]: preconditioner cells with no-normal-flux lines
  GatherPlan* gather = nullptr;    // write-once path (system matrix, n_other == 0)
  std::vector<int32_t> h_cells, h_nnf_idx, h_wide_idx;   // host copies: plan order, 9-table / wide-table index per plan cell
  std::vector<uint8_t> h_cflag;                          // the plan cell holds constrained velocity dofs
  int max_off_plain = 0;   // preconditioner plan: largest velocity-row offset used by cells without no-normal-flux lines
};

struct dcp_model {
  dcp_ctx* ctx = nullptr;
  int dim = 3, family = 0, strategy = DCP_STRATEGY_POSITIONS;
  int64_t n_cells = 0;
  // spaces
  int nse_n_local = 0, nse_nb = 2, temp_n_local = 0;
  int64_t nse_n_dofs = 0, temp_n_dofs = 0;
  int32_t *nse_l2g = nullptr, *temp_l2g = nullptr;
  int32_t* vel_dof = nullptr;       // [dim][ndu] cell dof of (velocity component, node), classic family
  double* cell_vertices = nullptr;  // [n_cells][2^dim][dim] (optional)
  int64_t n_owned_cells = 0;
  uint8_t* temp_bc_flag = nullptr;   // [n_cells] 1: the cell holds an inhomogeneously constrained temperature dof
  int32_t* temp_bc_cells = nullptr;  // those cells
  int64_t n_temp_bc_cells = 0;
  uint16_t* temp_pos = nullptr;  // [n_cells][nd*nd] scatter positions of the temperature matrices (0xffff row: general)
  double* temp_q2_tab = nullptr;   // Q2 temperature matrix kernel: reference table in table-build order
  int32_t *temp_fast_cells = nullptr, *temp_general_cells = nullptr;   // cells with / without a position row
  int64_t n_temp_fast = 0, n_temp_general = 0;
  int32_t *nse_local_field = nullptr, *nse_local_base = nullptr;
  std::vector<int32_t> h_local_field, h_local_base;
  DevCs nse_cs, temp_cs;
  // cells that hold at least one constrained dof (for the owner strategy's fix-up pass)
  int32_t* nse_constrained_cells = nullptr;
  int64_t n_nse_constrained_cells = 0;
  // tables
  int nq_nse = 0, nq_temp = 0, ndu = 0, ndp = 0, ndt = 0;
  double *phi_u_qn = nullptr, *dphi_u_qn = nullptr, *phi_p_qn = nullptr, *phi_t_qn = nullptr;
  double *phi_u_qt = nullptr, *phi_t_qt = nullptr, *dphi_t_qt = nullptr;
  double *phi_u_qt_T = nullptr, *phi_t_qt_T = nullptr, *dphi_t_qt_T = nullptr;  // point-fastest copies [function(,dim)][q]
  // mapping data
  double *geom_qn = nullptr, *geom_qt = nullptr;
  bool geom_shared = false;
  int gs_n = 0, gs_t = 0, gs_p = 0;  // doubles per cell record on the NSE / temperature / preconditioner rule
  // FEEC family
  int nq_pre = 0;
  double *geom_qp = nullptr, *nse_sign = nullptr;
  double *feec_w_qn = nullptr, *feec_c_qn = nullptr, *feec_u_qn = nullptr;
  double *feec_w_qn_t = nullptr, *feec_c_qn_t = nullptr, *feec_u_qn_t = nullptr;  // same, stored [function][component][q]
  double *feec_w_qp = nullptr, *feec_c_qp = nullptr, *feec_u_qp = nullptr;
  double *feec_u_qt = nullptr, *feec_div = nullptr;
  int32_t* feec_general_cells = nullptr;  // FEEC cells with constrained dofs (general scatter)
  int32_t* feec_fast_cells = nullptr;     // the others
  int64_t n_feec_general = 0, n_feec_fast = 0;
  uint16_t *feec_pos_nse = nullptr, *feec_pos_pre = nullptr;  // [n_cells][19*19] scatter positions (FEEC)
  // matrices and vectors
  BlockMat nse, pre, tmass, tstiff, tmat;
  double *nse_rhs = nullptr, *temp_rhs = nullptr;
  bool temp_matrices_ready = false;
  bool pre_fused_valid = false;     // nse_preconditioner_matrix was written by the last staged system pass ...
  double pre_fused_dt = 0.0, pre_fused_inv_re = 0.0;   // ... for these parameters
  OwnerPlan* owner_nse = nullptr;
  OwnerPlan* owner_pre = nullptr;
  FastPlan* fast_nse = nullptr;
  FastPlan* fast_pre = nullptr;
  MaskedPlan* masked_nse = nullptr;
  MaskedPlan* masked_pre = nullptr;
  std::vector<struct dcp_ilu*> ilus;  // live ILU handles of this model (destroyed with it)
};

// ---- helpers implemented in context.cu -----------------------------------------------------------
template <class T>
int dcp_upload(dcp_ctx* ctx, T** dst, const T* src, int64_t n);
int dcp_check_device_errors(dcp_ctx* ctx, const char* what);
int dcp_stage_in(dcp_ctx* ctx, int slot, const double* src, int64_t n, int mem, const double** dev);
int dcp_stage_out_alloc(dcp_ctx* ctx, int slot, double* dst, int64_t n, int mem, double** dev);
int dcp_stage_out_finish(dcp_ctx* ctx, int slot, double* dst, int64_t n, int mem);

BlockMat* dcp_select_matrix(dcp_model* m, int which);
extern "C" int dcp_ilu_destroy(struct dcp_ilu* p);

// ---- kernels' host launchers ---------------------------------------------------------------------
int dcp_feec_positions_build(dcp_model* m, const dcp_model_desc* d, bool system, uint16_t** out);
int dcp_launch_spmv(dcp_ctx* ctx, const DevCsr& A, const double* x, double* y, bool add, int64_t row_limit = -1,
                    const unsigned char* skip = nullptr, const int* list = nullptr, int64_t n_list = 0);
int dcp_build_ghost_rows(dcp_ctx* ctx, BlockMat& M, int r, const int64_t* owned_cols);
int dcp_launch_gather(dcp_ctx* ctx, int64_t n, const int32_t* idx, const double* src, double* dst, bool scatter);
int dcp_launch_extract_diag_inv(dcp_ctx* ctx, const DevCsr& A, double* diag_inv, int32_t** diag_off = nullptr);
int dcp_launch_jacobi(dcp_ctx* ctx, int64_t n, const double* diag_inv, const double* x, double* y);
int dcp_launch_axpby_values(dcp_ctx* ctx, int64_t n, const double* a, const double* b, double fb, double* out);
int dcp_launch_fill(dcp_ctx* ctx, double* p, int64_t n, double v);
int dcp_launch_velocity_extrema(dcp_model* m, const double* nse_solution, double* out2_dev);
int dcp_launch_distribute(dcp_ctx* ctx, const DevCs& cs, double* x);

int dcp_launch_th_cells(dcp_model* m, const dcp_params& p, bool system, const double* old_nse, const double* old_temp,
                        const int32_t* cell_list, int64_t n_list, bool only_constrained_entries);
int dcp_launch_temperature_matrix(dcp_model* m, const dcp_params& p);
int dcp_launch_temperature_rhs(dcp_model* m, const dcp_params& p, const double* old_temp, const double* nse_solution);

int dcp_launch_feec(dcp_model* m, const dcp_params& p, bool system, const double* old_nse, const double* old_temp);
int dcp_fast_plan_build(dcp_model* m, const dcp_model_desc* d, bool system, FastPlan** out);
void dcp_fast_plan_free(FastPlan* p);
int64_t dcp_fast_plan_counts(const FastPlan* p, int64_t* n_general);
const int32_t* dcp_fast_plan_general_cells(const FastPlan* p);
int dcp_launch_th_fast(dcp_model* m, const dcp_params& p, bool system, const FastPlan* plan, const double* old_nse,
                       const double* old_temp);
int dcp_masked_plan_build(dcp_model* m, const dcp_model_desc* d, bool system, MaskedPlan** out);
void dcp_masked_plan_free(MaskedPlan* p);
int dcp_launch_th_mma(dcp_model* m, const dcp_params& p, bool system, const MaskedPlan* plan, const double* old_nse,
                      const double* old_temp, const int32_t* wlist = nullptr, int64_t n_list = 0);
int dcp_gather_plan_attach_pre(dcp_model* m, const dcp_model_desc* d, GatherPlan* G, const MaskedPlan* nse_plan, const MaskedPlan* pre_plan);
int dcp_gather_plan_build(dcp_model* m, const dcp_model_desc* d, const std::vector<int32_t>& cells,
                          const std::vector<uint8_t>& cell_has_constraints, GatherPlan** out);
void dcp_gather_plan_free(GatherPlan* p);
int dcp_launch_th_staged(dcp_model* m, const dcp_params& p, const MaskedPlan* plan, const double* old_nse, const double* old_temp);
int dcp_owner_plan_build(dcp_model* m, bool system, const dcp_model_desc* desc);
void dcp_owner_plan_free(OwnerPlan* p);
int dcp_launch_th_owner(dcp_model* m, const dcp_params& p, bool system);
int dcp_launch_th_rhs(dcp_model* m, const dcp_params& p, const double* old_nse, const double* old_temp);
