// C++ host mirror (include/dcp.hpp) exercised the way the reference's own code would use it, checked against the
// CPU oracle (TEST INFRASTRUCTURE: this program links oracle/_build/liboracle.so as the checker only).
//
//   host_mirror_test --no-gpu   : no CUDA device -> dcp::Context must throw dcp::Error (no CPU fallback)
//   host_mirror_test [spec]     : assemble the classic system on the stand-in problem, compare matrix, right-hand
//                                 side and a block vmult with the oracle (1e-12), run a Jacobi-preconditioned CG on
//                                 the temperature system with device-resident vectors (boussinesq_model.tpp:1417-1440)
#include <dcp.hpp>
#include <dcp_harness.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

extern "C" {
#include "../../oracle/oracle_common.h"
void orc_assemble_nse_system(const orc_params* P, int64_t n_cells, int nd, int nq, int ndu, int ndp, int ndt, const int32_t* field,
                             const int32_t* base, const double* phi_u, const double* dphi_u, const double* phi_p,
                             const double* phi_t, const double* geom, const int32_t* l2g, const int32_t* l2g_t,
                             const double* old_nse, const double* old_temp, const orc_constraints* cs, orc_csr* A, double* rhs,
                             int64_t n_rhs, int use_omp);
void orc_spmv(int64_t n_rows, const int64_t* rowptr, const int32_t* col, const double* val, const double* x, double* y, int add,
              int use_omp);
}

namespace {

struct Problem {
  dcph_problem* p;
  explicit Problem(const std::string& spec) : p(dcph_create(spec.c_str())) {
    if (!p) throw std::runtime_error(std::string("harness: ") + dcph_last_error());
  }
  ~Problem() { dcph_destroy(p); }
  template <class T>
  const T* arr(const char* name, int64_t* count = nullptr) const {
    const void* d = nullptr;
    int64_t n = 0;
    int dt = 0;
    if (dcph_array(p, name, &d, &n, &dt)) throw std::runtime_error(std::string("harness: no array ") + name);
    if (count) *count = n;
    return static_cast<const T*>(d);
  }
  int64_t scalar(const char* name) const { return dcph_scalar(p, name); }
  dcp_constraints_desc cs(const std::string& prefix) const {
    dcp_constraints_desc c{};
    c.n_dofs = scalar(prefix == "nse.cs" ? "nse.n_dofs" : "temp.n_dofs");
    int64_t nl = 0;
    c.line_dof = arr<int32_t>((prefix + ".line_dof").c_str(), &nl);
    c.n_lines = nl;
    c.line_ptr = arr<int32_t>((prefix + ".line_ptr").c_str());
    c.entry_dof = arr<int32_t>((prefix + ".entry_dof").c_str());
    c.entry_w = arr<double>((prefix + ".entry_w").c_str());
    c.inhom = arr<double>((prefix + ".inhom").c_str());
    return c;
  }
  dcp_csr_desc csr(const std::string& name) const {
    dcp_csr_desc d{};
    d.n_rows = scalar((name + ".n_rows").c_str());
    d.n_cols = scalar((name + ".n_cols").c_str());
    if (scalar((name + ".nnz").c_str()) > 0) {
      d.rowptr = arr<int64_t>((name + ".rowptr").c_str());
      d.col = arr<int32_t>((name + ".col").c_str());
    }
    return d;
  }
};

dcp_model_desc classic_desc(const Problem& P) {
  dcp_model_desc d{};
  d.dim = 3;
  d.family = DCP_FAMILY_CLASSIC;
  d.n_cells = P.scalar("n_cells");
  d.nse_n_local = (int32_t)P.scalar("nse.n_local");
  d.nse_n_blocks = 2;
  d.nse_block_size[0] = P.scalar("nse.n_u");
  d.nse_block_size[1] = P.scalar("nse.n_p");
  d.nse_l2g = P.arr<int32_t>("nse.l2g");
  d.nse_local_field = P.arr<int32_t>("nse.local_field");
  d.nse_local_base = P.arr<int32_t>("nse.local_base");
  d.nse_cs = P.cs("nse.cs");
  d.temp_n_local = (int32_t)P.scalar("temp.n_local");
  d.temp_l2g = P.arr<int32_t>("temp.l2g");
  d.temp_cs = P.cs("temp.cs");
  d.nq_nse = (int32_t)P.scalar("q_nse.nq");
  d.nq_temp = (int32_t)P.scalar("q_temp.nq");
  d.ndu = (int32_t)P.scalar("tab.u_qn.nd");
  d.ndp = (int32_t)P.scalar("tab.p_qn.nd");
  d.ndt = (int32_t)P.scalar("tab.t_qn.nd");
  d.phi_u_qn = P.arr<double>("tab.u_qn.phi");
  d.dphi_u_qn = P.arr<double>("tab.u_qn.dphi");
  d.phi_p_qn = P.arr<double>("tab.p_qn.phi");
  d.phi_t_qn = P.arr<double>("tab.t_qn.phi");
  d.phi_u_qt = P.arr<double>("tab.u_qt.phi");
  d.phi_t_qt = P.arr<double>("tab.t_qt.phi");
  d.dphi_t_qt = P.arr<double>("tab.t_qt.dphi");
  d.geom_qn = P.arr<double>("geom.qn");
  d.geom_qt = P.arr<double>("geom.qt");
  const char* nb[2][2] = {{"b00", "b01"}, {"b10", "b11"}};
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) {
      d.nse_pattern[i][j] = P.csr(std::string("nse.") + nb[i][j]);
      d.pre_pattern[i][j] = P.csr(std::string("pre.") + nb[i][j]);
    }
  d.temp_pattern = P.csr("temp.pat");
  d.cell_vertices = P.arr<double>("cell_vertices");
  return d;
}

double max_abs(const std::vector<double>& v) {
  double m = 0;
  for (double x : v) m = std::fmax(m, std::fabs(x));
  return m;
}

// SolverCG<LA::MPI::Vector>::solve(A, x, b, P) written against the operator concept only
template <class Matrix, class Precond>
int solver_cg(dcp::Context& ctx, const Matrix& A, dcp::DeviceVector& x, const dcp::DeviceVector& b, const Precond& P, double tol,
              int max_it) {
  const int64_t n = x.size();
  dcp::DeviceVector r(ctx, n), z(ctx, n), p(ctx, n), Ap(ctx, n);
  A.vmult(r, x);
  r.sadd(-1.0, 1.0, b);  // r = b - A x
  if (std::sqrt(r * r) <= tol) return 0;
  P.vmult(z, r);
  p.equ(z);
  double rz = r * z;
  for (int it = 1; it <= max_it; ++it) {
    A.vmult(Ap, p);
    const double alpha = rz / (p * Ap);
    x.add(alpha, p);
    r.add(-alpha, Ap);
    if (std::sqrt(r * r) <= tol) return it;
    P.vmult(z, r);
    const double rz_new = r * z;
    p.sadd(rz_new / rz, 1.0, z);
    rz = rz_new;
  }
  return -1;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc > 1 && !std::strcmp(argv[1], "--no-gpu")) {
    try {
      dcp::Context ctx(0);
    } catch (const dcp::Error& e) {
      std::printf("no device: dcp::Error(%d) \"%s\" -- no CPU fallback, as designed\n", e.code(), e.what());
      return e.code() == DCP_ERR_CUDA ? 0 : 2;
    }
    std::printf("a CUDA device is present; --no-gpu expects none\n");
    return 3;
  }
  try {
    const std::string spec = argc > 1 ? argv[1] : "geometry=shell,refine=1";
    Problem P(spec);
    const dcp_params prm{3, 0, 1, 0, 0.1, 0.01, 0.001, 0.2, 2.0, 1.0, 1.0, 1.0, 1.0};  // data/aqua_planet_shell_test_3d-classic.prm
    const int64_t n = P.scalar("nse.n_dofs"), nT = P.scalar("temp.n_dofs"), n_u = P.scalar("nse.n_u");
    std::vector<double> u((size_t)n), T((size_t)nT);
    for (int64_t i = 0; i < n; ++i) u[i] = 0.1 * std::sin(0.37 * (double)i) + 0.05;
    for (int64_t i = 0; i < nT; ++i) T[i] = 2.0 + 0.3 * std::cos(0.11 * (double)i);

    dcp::Context ctx(0);
    dcp::BoussinesqModel model(ctx, classic_desc(P), prm);
    model.assemble_nse_system(u, T);
    model.build_nse_preconditioner();
    model.assemble_temperature_matrix();
    model.assemble_temperature_rhs(T, u);

    // ---- oracle on the same inputs
    int64_t nnz = 0;
    const int64_t* rp = P.arr<int64_t>("nse.full.rowptr");
    const int32_t* col = P.arr<int32_t>("nse.full.col", &nnz);
    std::vector<double> ref_val((size_t)nnz), ref_rhs((size_t)n);
    orc_csr A{n, rp, col, ref_val.data()};
    orc_constraints cs{n, P.arr<int32_t>("nse.cs.line_of_dof"), P.arr<int32_t>("nse.cs.line_ptr"), P.arr<int32_t>("nse.cs.entry_dof"),
                       P.arr<double>("nse.cs.entry_w"), P.arr<double>("nse.cs.inhom")};
    orc_params op;
    static_assert(sizeof(orc_params) == sizeof(dcp_params), "parameter structs mirror each other");
    std::memcpy(&op, &prm, sizeof op);
    orc_assemble_nse_system(&op, P.scalar("n_cells"), (int)P.scalar("nse.n_local"), (int)P.scalar("q_nse.nq"), (int)P.scalar("tab.u_qn.nd"),
                            (int)P.scalar("tab.p_qn.nd"), (int)P.scalar("tab.t_qn.nd"), P.arr<int32_t>("nse.local_field"),
                            P.arr<int32_t>("nse.local_base"), P.arr<double>("tab.u_qn.phi"), P.arr<double>("tab.u_qn.dphi"),
                            P.arr<double>("tab.p_qn.phi"), P.arr<double>("tab.t_qn.phi"), P.arr<double>("geom.qn"), P.arr<int32_t>("nse.l2g"),
                            P.arr<int32_t>("temp.l2g"), u.data(), T.data(), &cs, &A, ref_rhs.data(), n, 1);

    // right-hand side
    const std::vector<double> rhs = model.nse_rhs();
    double err = 0;
    for (int64_t i = 0; i < n; ++i) err = std::fmax(err, std::fabs(rhs[i] - ref_rhs[i]));
    const double rhs_rel = err / max_abs(ref_rhs);
    // block(0,0) values: rows < n_u, columns < n_u of the concatenated pattern, in order
    const std::vector<double> v00 = model.nse_matrix.block(0, 0).values();
    size_t k = 0;
    double verr = 0, vmax = 0;
    for (int64_t r = 0; r < n_u; ++r)
      for (int64_t q = rp[r]; q < rp[r + 1]; ++q)
        if (col[q] < n_u) {
          verr = std::fmax(verr, std::fabs(v00[k++] - ref_val[q]));
          vmax = std::fmax(vmax, std::fabs(ref_val[q]));
        }
    if (k != v00.size()) throw std::runtime_error("block(0,0) pattern size mismatch");
    // whole block matrix times a vector, host vectors through the operator concept
    std::vector<double> x((size_t)n), y((size_t)n), y_ref((size_t)n), y_abs((size_t)n), ax((size_t)n), aval(ref_val);
    for (int64_t i = 0; i < n; ++i) x[i] = std::sin(1.3 * (double)i), ax[i] = std::fabs(x[i]);
    for (double& a : aval) a = std::fabs(a);
    model.nse_matrix.vmult(y, x);
    orc_spmv(n, rp, col, ref_val.data(), x.data(), y_ref.data(), 0, 1);
    orc_spmv(n, rp, col, aval.data(), ax.data(), y_abs.data(), 0, 1);
    double yerr = 0;
    for (int64_t i = 0; i < n; ++i) yerr = std::fmax(yerr, std::fabs(y[i] - y_ref[i]));
    const double y_rel = yerr / max_abs(y_abs);
    std::printf("nse_rhs rel err %.3e, block(0,0) rel err %.3e, vmult rel err %.3e\n", rhs_rel, verr / vmax, y_rel);
    if (!(rhs_rel <= 1e-12 && verr / vmax <= 1e-12 && y_rel <= 1e-12)) return 1;

    // temperature solve with device-resident Krylov vectors (boussinesq_model.tpp:1417-1440)
    const std::vector<double> trhs = model.temperature_rhs();
    dcp::DeviceVector b(ctx, trhs), xt(ctx, T);
    const double tol = 1e-12 * std::sqrt(b * b);
    const int its = solver_cg(ctx, model.temperature_matrix, xt, b, model.T_preconditioner, tol, (int)nT);
    dcp::DeviceVector res(ctx, nT);
    model.temperature_matrix.vmult(res, xt);
    res.sadd(-1.0, 1.0, b);
    const double rel_res = std::sqrt(res * res) / std::sqrt(b * b);
    // the same solve resident on the device (dcp::solve_cg -> dcp_cg_solve): same step count, same iterate
    dcp::DeviceVector xr(ctx, T);
    double last_res = 0.0;
    const int64_t its_r = dcp::solve_cg(model.temperature_matrix, xr, b, &model.T_preconditioner, tol, nT, &last_res);
    const std::vector<double> xa = xt.download(), xb = xr.download();
    double dmax = 0.0, xmax = 0.0;
    for (size_t i = 0; i < xa.size(); ++i) dmax = std::fmax(dmax, std::fabs(xa[i] - xb[i])), xmax = std::fmax(xmax, std::fabs(xa[i]));
    std::printf("resident CG: %lld iterations, last residual %.3e, max difference to the host-driven loop %.3e\n", (long long)its_r, last_res,
                dmax / xmax);
    if (its_r != its || dmax > 1e-12 * xmax || last_res > tol) return 1;
    model.distribute_temperature_constraints(xt);
    const auto vc = model.velocity_extrema(u);
    std::printf("temperature CG: %d iterations, relative residual %.3e; max |u| %.6f, CFL %.6f\n", its, rel_res, vc.first, vc.second);
    if (its <= 0 || rel_res > 1e-10) return 1;
    std::printf("OK\n");
    return 0;
  } catch (const std::exception& e) {
    std::printf("FAILED: %s\n", e.what());
    return 1;
  }
}
