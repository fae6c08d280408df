// dcp.hpp -- header-only C++ host mirror over the C ABI of libdcp.so (include/dcp.h).
//
// The reference is C++ (deal.II).  This header gives the device path the shapes the reference's code is written
// against, so that the adapter of INTEGRATION.md is a handful of one-line substitutions:
//   * dcp::SparseMatrix / dcp::BlockSparseMatrix / dcp::PreconditionJacobi satisfy the deal.II operator concept
//     (`void vmult(Vec &dst, const Vec &src) const`, `vmult_add`, `m()`, `n()`, `block(i,j)`) that every template
//     in include/linear_algebra/*.hpp and SolverCG/GMRES/FGMRES::solve(A, x, b, P) take as MatrixType
//     (schur_complement.hpp:50-56, inverse_matrix.hpp:36-58, boussinesq_model.tpp:1196-1199, 1437-1440);
//   * dcp::BoussinesqModel carries the member names of Standard::BoussinesqModel<dim>
//     (include/core/boussinesq_model.h:168-180, 220-250): assemble_nse_system, assemble_nse_preconditioner,
//     build_nse_preconditioner, assemble_temperature_matrix, assemble_temperature_rhs, get_maximal_velocity,
//     get_cfl_number; matrices nse_matrix, nse_preconditioner_matrix, temperature_{mass,stiffness,}_matrix;
//     preconditioners Mu_plus_A_preconditioner, Mp_preconditioner, T_preconditioner;
//   * errors: every non-zero status becomes dcp::Error (a std::runtime_error), which is what the handler in
//     source/main.cxx:128-156 catches.
// Vectors: any contiguous container of double with data()/size() is taken as HOST memory (copied through the
// library's staging buffers); dcp::DeviceVector is device memory and is passed through untouched.
// No CPU fallback: constructing a Context without a CUDA device throws.
#ifndef DCP_HPP
#define DCP_HPP
#include <dcp.h>

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace dcp {

class Error : public std::runtime_error {
 public:
  Error(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
  int code() const { return code_; }

 private:
  int code_;
};

inline void check(int rc, const char* where) {
  if (rc != DCP_OK) throw Error(rc, std::string(where) + ": " + dcp_last_error());
}

class Context {
 public:
  explicit Context(int device = 0) { check(dcp_ctx_create(device, &h_), "dcp_ctx_create"); }
  ~Context() {
    if (h_) dcp_ctx_destroy(h_);
  }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  dcp_ctx* get() const { return h_; }
  void set_stream(void* cuda_stream) { check(dcp_ctx_set_stream(h_, cuda_stream), "dcp_ctx_set_stream"); }
  void synchronize() { check(dcp_ctx_synchronize(h_), "dcp_ctx_synchronize"); }
  int64_t launch_count() const { return dcp_ctx_launch_count(h_); }

 private:
  dcp_ctx* h_ = nullptr;
};

// A vector in device memory (what LA::MPI::Vector is to the reference once the Krylov vectors live in HBM).
class DeviceVector {
 public:
  DeviceVector(Context& ctx, int64_t n) : ctx_(&ctx), n_(n) {
    void* p = nullptr;
    check(dcp_malloc(ctx.get(), 8 * (n > 0 ? n : 1), &p), "dcp_malloc");
    d_ = static_cast<double*>(p);
  }
  DeviceVector(Context& ctx, const std::vector<double>& host) : DeviceVector(ctx, (int64_t)host.size()) { upload(host); }
  ~DeviceVector() {
    if (d_) dcp_free(ctx_->get(), d_);
  }
  DeviceVector(const DeviceVector&) = delete;
  DeviceVector& operator=(const DeviceVector&) = delete;
  DeviceVector(DeviceVector&& o) noexcept : ctx_(o.ctx_), d_(o.d_), n_(o.n_) { o.d_ = nullptr; }
  int64_t size() const { return n_; }
  double* data() { return d_; }
  const double* data() const { return d_; }
  void upload(const std::vector<double>& host) {
    if ((int64_t)host.size() != n_) throw Error(DCP_ERR_ARG, "DeviceVector::upload: size mismatch");
    check(dcp_memcpy_h2d(ctx_->get(), d_, host.data(), 8 * n_), "dcp_memcpy_h2d");
  }
  std::vector<double> download() const {
    std::vector<double> out((size_t)n_);
    check(dcp_memcpy_d2h(ctx_->get(), out.data(), d_, 8 * n_), "dcp_memcpy_d2h");
    return out;
  }
  // Trilinos-vector algebra used by the deal.II solvers (operator*, add, sadd, operator*=, operator=)
  double operator*(const DeviceVector& y) const {
    double r = 0;
    check(dcp_vec_dot(ctx_->get(), n_, d_, y.d_, &r), "dcp_vec_dot");
    return r;
  }
  void add(double a, const DeviceVector& x) { check(dcp_vec_axpy(ctx_->get(), n_, a, x.d_, d_), "dcp_vec_axpy"); }
  void sadd(double s, double a, const DeviceVector& x) { check(dcp_vec_sadd(ctx_->get(), n_, s, a, x.d_, d_), "dcp_vec_sadd"); }
  DeviceVector& operator*=(double a) {
    check(dcp_vec_scale(ctx_->get(), n_, a, d_), "dcp_vec_scale");
    return *this;
  }
  void equ(const DeviceVector& x) { check(dcp_vec_copy(ctx_->get(), n_, x.d_, d_), "dcp_vec_copy"); }

 private:
  Context* ctx_;
  double* d_ = nullptr;
  int64_t n_ = 0;
};

namespace detail {
// (pointer, memory space) of a vector argument
inline std::pair<double*, int> arg(DeviceVector& v) { return {v.data(), DCP_DEVICE}; }
inline std::pair<const double*, int> arg(const DeviceVector& v) { return {v.data(), DCP_DEVICE}; }
template <class V>
std::pair<double*, int> arg(V& v) {
  return {v.data(), DCP_HOST};
}
template <class V>
std::pair<const double*, int> arg(const V& v) {
  return {v.data(), DCP_HOST};
}
inline void same_space(int a, int b) {
  if (a != b) throw Error(DCP_ERR_ARG, "source and destination vectors must live in the same memory space");
}
}  // namespace detail

// LA::SparseMatrix (one CSR block of a model matrix)
class SparseMatrix {
 public:
  SparseMatrix() = default;
  SparseMatrix(dcp_model* m, int which, int bi, int bj) : m_(m), which_(which), bi_(bi), bj_(bj) {
    check(dcp_matrix_info(m_, which_, bi_, bj_, &rows_, &cols_, &nnz_), "dcp_matrix_info");
  }
  int64_t m() const { return rows_; }
  int64_t n() const { return cols_; }
  int64_t n_nonzero_elements() const { return nnz_; }
  dcp_model* model() const { return m_; }
  int which() const { return which_; }
  int block_row() const { return bi_; }
  int block_col() const { return bj_; }
  template <class Dst, class Src>
  void vmult(Dst& dst, const Src& src) const {
    auto d = detail::arg(dst);
    auto s = detail::arg(src);
    detail::same_space(d.second, s.second);
    check(dcp_vmult(m_, which_, bi_, bj_, d.first, s.first, d.second), "dcp_vmult");
  }
  template <class Dst, class Src>
  void vmult_add(Dst& dst, const Src& src) const {
    auto d = detail::arg(dst);
    auto s = detail::arg(src);
    detail::same_space(d.second, s.second);
    check(dcp_vmult_add(m_, which_, bi_, bj_, d.first, s.first, d.second), "dcp_vmult_add");
  }
  std::vector<double> values() const {
    std::vector<double> v((size_t)nnz_);
    if (nnz_) check(dcp_matrix_download(m_, which_, bi_, bj_, v.data()), "dcp_matrix_download");
    return v;
  }

 private:
  dcp_model* m_ = nullptr;
  int which_ = 0, bi_ = 0, bj_ = 0;
  int64_t rows_ = 0, cols_ = 0, nnz_ = 0;
};

// LA::BlockSparseMatrix over block-concatenated vectors
class BlockSparseMatrix {
 public:
  using BlockType = SparseMatrix;
  BlockSparseMatrix() = default;
  BlockSparseMatrix(dcp_model* m, int which, int n_blocks) : m_(m), which_(which), nb_(n_blocks) {
    for (int i = 0; i < nb_; ++i)
      for (int j = 0; j < nb_; ++j) blocks_[i][j] = SparseMatrix(m, which, i, j);
  }
  const SparseMatrix& block(int i, int j) const { return blocks_[i][j]; }
  int n_block_rows() const { return nb_; }
  int n_block_cols() const { return nb_; }
  int64_t m() const {
    int64_t s = 0;
    for (int i = 0; i < nb_; ++i) s += blocks_[i][0].m();
    return s;
  }
  template <class Dst, class Src>
  void vmult(Dst& dst, const Src& src) const {
    auto d = detail::arg(dst);
    auto s = detail::arg(src);
    detail::same_space(d.second, s.second);
    check(dcp_block_vmult(m_, which_, d.first, s.first, d.second), "dcp_block_vmult");
  }

 private:
  dcp_model* m_ = nullptr;
  int which_ = 0, nb_ = 0;
  SparseMatrix blocks_[DCP_MAX_BLOCKS][DCP_MAX_BLOCKS];
};

// LA::PreconditionJacobi initialised with one diagonal block (boussinesq_model.tpp:520-542, 980-986)
class PreconditionJacobi {
 public:
  PreconditionJacobi() = default;
  PreconditionJacobi(dcp_model* m, int which, int bi) : m_(m), which_(which), bi_(bi) {}
  int which() const { return which_; }
  int block() const { return bi_; }
  template <class Dst, class Src>
  void vmult(Dst& dst, const Src& src) const {
    auto d = detail::arg(dst);
    auto s = detail::arg(src);
    detail::same_space(d.second, s.second);
    check(dcp_jacobi_vmult(m_, which_, bi_, d.first, s.first, d.second), "dcp_jacobi_vmult");
  }

 private:
  dcp_model* m_ = nullptr;
  int which_ = 0, bi_ = 0;
};

// SolverCG<LA::MPI::Vector>(SolverControl(max_steps, tol)).solve(A, x, b, P) on device vectors, P = PreconditionJacobi (or
// the identity with `preconditioner == nullptr`): the loop of LinearAlgebra::InverseMatrix::vmult
// (include/linear_algebra/inverse_matrix.hpp:90-121) and of solve_temperature (boussinesq_model.tpp:1426-1440), resident on
// the device (dcp_cg_solve).  Returns the last step; throws like SolverControl::NoConvergence when the limit is reached.
inline int64_t solve_cg(const SparseMatrix& A, DeviceVector& x, const DeviceVector& b, const PreconditionJacobi* preconditioner, double tol,
                        int64_t max_steps, double* last_residual = nullptr) {
  int64_t step = 0;
  double res = 0.0;
  const int rc = dcp_cg_solve(A.model(), A.which(), A.block_row(), A.block_col(), preconditioner ? DCP_PRECOND_JACOBI : DCP_PRECOND_IDENTITY,
                              preconditioner ? preconditioner->which() : 0, preconditioner ? preconditioner->block() : 0, nullptr, x.data(),
                              b.data(), tol, max_steps, 0, &step, &res);
  if (last_residual) *last_residual = res;
  if (rc != DCP_OK) throw Error(rc, "dcp_cg_solve");
  return step;
}

// Device-side state of Standard::BoussinesqModel<dim> / ExteriorCalculus::BoussinesqModel<3> for one mesh.
class BoussinesqModel {
 public:
  dcp_params parameters;  // the pointwise coefficients (dt, 1/Re, 1/Pe, ...); may be changed between calls

  BoussinesqModel(Context& ctx, const dcp_model_desc& desc, const dcp_params& prm) : parameters(prm), ctx_(&ctx) {
    check(dcp_model_create(ctx.get(), &desc, &h_), "dcp_model_create");
    const int nb = desc.nse_n_blocks;
    for (int b = 0; b < nb; ++b) n_nse_ += desc.nse_block_size[b];
    n_temp_ = desc.temp_cs.n_dofs;
    nse_matrix = BlockSparseMatrix(h_, DCP_MAT_NSE, nb);
    nse_preconditioner_matrix = BlockSparseMatrix(h_, DCP_MAT_NSE_PRECOND, nb);
    temperature_mass_matrix = SparseMatrix(h_, DCP_MAT_TEMP_MASS, 0, 0);
    temperature_stiffness_matrix = SparseMatrix(h_, DCP_MAT_TEMP_STIFF, 0, 0);
    temperature_matrix = SparseMatrix(h_, DCP_MAT_TEMP, 0, 0);
    Mu_plus_A_preconditioner = PreconditionJacobi(h_, DCP_MAT_NSE_PRECOND, 0);
    Mp_preconditioner = PreconditionJacobi(h_, DCP_MAT_NSE_PRECOND, nb - 1);
    T_preconditioner = PreconditionJacobi(h_, DCP_MAT_TEMP, 0);
  }
  ~BoussinesqModel() {
    if (h_) dcp_model_destroy(h_);
  }
  BoussinesqModel(const BoussinesqModel&) = delete;
  BoussinesqModel& operator=(const BoussinesqModel&) = delete;

  dcp_model* get() const { return h_; }
  int64_t n_nse_dofs() const { return n_nse_; }
  int64_t n_temperature_dofs() const { return n_temp_; }
  void set_strategy(int strategy) { check(dcp_model_set_strategy(h_, strategy), "dcp_model_set_strategy"); }
  int strategy() const { return dcp_model_get_strategy(h_); }
  void set_owned(const std::array<int64_t, DCP_MAX_BLOCKS>& nse_owned_per_block, int64_t temp_owned) {
    check(dcp_model_set_owned(h_, nse_owned_per_block.data(), temp_owned), "dcp_model_set_owned");
  }

  // ---- the four assemblers (boussinesq_model.h:168-180) ----
  template <class V>
  void assemble_nse_system(const V& old_nse_solution, const V& old_temperature_solution) {
    auto a = detail::arg(old_nse_solution);
    auto b = detail::arg(old_temperature_solution);
    check(dcp_assemble_nse_system(h_, &parameters, a.first, b.first, a.second), "assemble_nse_system");
  }
  void assemble_nse_preconditioner() { check(dcp_assemble_nse_preconditioner(h_, &parameters), "assemble_nse_preconditioner"); }
  // the Jacobi set-up of build_nse_preconditioner (:520-542) is part of the device call
  void build_nse_preconditioner() { assemble_nse_preconditioner(); }
  void assemble_temperature_matrix() { check(dcp_assemble_temperature_matrix(h_, &parameters), "assemble_temperature_matrix"); }
  template <class V>
  void assemble_temperature_rhs(const V& old_temperature_solution, const V& nse_solution) {
    auto a = detail::arg(old_temperature_solution);
    auto b = detail::arg(nse_solution);
    check(dcp_assemble_temperature_rhs(h_, &parameters, a.first, b.first, a.second), "assemble_temperature_rhs");
  }

  // ---- passes next to the solves (:1023-1098, 1233, 1442) ----
  template <class V>
  std::pair<double, double> velocity_extrema(const V& nse_solution) const {
    auto a = detail::arg(nse_solution);
    double out[2] = {0, 0};
    check(dcp_velocity_extrema(h_, a.first, a.second, out), "dcp_velocity_extrema");
    return {out[0], out[1]};
  }
  template <class V>
  double get_maximal_velocity(const V& nse_solution) const {
    return velocity_extrema(nse_solution).first;
  }
  template <class V>
  double get_cfl_number(const V& nse_solution) const {
    return velocity_extrema(nse_solution).second;
  }
  template <class V>
  void distribute_nse_constraints(V& x) const {
    auto a = detail::arg(x);
    check(dcp_constraints_distribute(h_, 0, a.first, a.second), "dcp_constraints_distribute");
  }
  template <class V>
  void distribute_temperature_constraints(V& x) const {
    auto a = detail::arg(x);
    check(dcp_constraints_distribute(h_, 1, a.first, a.second), "dcp_constraints_distribute");
  }

  // ---- right-hand sides ----
  std::vector<double> nse_rhs() const { return vector(DCP_VEC_NSE_RHS, n_nse_); }
  std::vector<double> temperature_rhs() const { return vector(DCP_VEC_TEMP_RHS, n_temp_); }

  BlockSparseMatrix nse_matrix, nse_preconditioner_matrix;
  SparseMatrix temperature_mass_matrix, temperature_stiffness_matrix, temperature_matrix;
  PreconditionJacobi Mu_plus_A_preconditioner, Mp_preconditioner, T_preconditioner;

 private:
  std::vector<double> vector(int which, int64_t n) const {
    std::vector<double> v((size_t)n);
    check(dcp_vector_download(h_, which, v.data()), "dcp_vector_download");
    return v;
  }
  Context* ctx_;
  dcp_model* h_ = nullptr;
  int64_t n_nse_ = 0, n_temp_ = 0;
};

}  // namespace dcp
#endif
