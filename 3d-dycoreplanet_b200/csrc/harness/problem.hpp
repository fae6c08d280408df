// Stand-in "problem builder": produces, without deal.II, exactly the arrays the C ABI in include/dcp.h
// takes -- reference-cell tables, per-cell mapping data (what FEValues::reinit would deliver: JxW,
// inverse Jacobians, quadrature points), cell->global DoF maps, constraint lines and CSR patterns --
// for the reference's classic Taylor-Hood model (include/core/boussinesq_model.tpp:15-412).
// A real deal.II build would fill the same arrays from FEValues / DoFHandler / AffineConstraints.
#pragma once
#include <map>
#include <memory>
#include <sstream>
#include <string>

#include "dofs.hpp"

namespace dcph {

enum DType { F64 = 0, I32 = 1, I64 = 2, I8 = 3, I16 = 4 };
struct ArrayRef {
  const void* p = nullptr;
  int64_t n = 0;
  int dtype = F64;
};

struct Spec {
  std::string geometry = "shell";  // shell | cube
  std::string family = "classic";  // classic | feec
  int dim = 3;
  int refine = 2;
  double R0 = 1.0, R1 = 3.0;
  int velocity_degree = 2;
  int temperature_degree = 1;
  int patterns = 1;       // build CSR patterns on the host
  int geometry_data = 1;  // build per-cell mapping data
  int mapping_degree = 3;
  int threads = 0;
  int constraints = 1;    // 0: no boundary conditions at all (used by the oracle identity tests)
  int radial_factor = 1;  // shell only: n_r = 2^refine * radial_factor layers (weak-scaling synthetic refinement)
  std::string renumber = "none";  // none | cuthill_mckee | random: order of the NSE dofs before component_wise
  int n_ranks = 1, rank = 0;  // contiguous partition of the (tree, Morton) cell order, one ghost-cell layer
};

inline Spec parse_spec(const std::string& s) {
  Spec sp;
  std::stringstream ss(s);
  std::string kv;
  while (std::getline(ss, kv, ',')) {
    auto eq = kv.find('=');
    if (eq == std::string::npos) continue;
    std::string k = kv.substr(0, eq), v = kv.substr(eq + 1);
    if (k == "geometry") sp.geometry = v;
    else if (k == "family") sp.family = v;
    else if (k == "dim") sp.dim = std::stoi(v);
    else if (k == "refine") sp.refine = std::stoi(v);
    else if (k == "R0") sp.R0 = std::stod(v);
    else if (k == "R1") sp.R1 = std::stod(v);
    else if (k == "velocity_degree") sp.velocity_degree = std::stoi(v);
    else if (k == "temperature_degree") sp.temperature_degree = std::stoi(v);
    else if (k == "patterns") sp.patterns = std::stoi(v);
    else if (k == "geometry_data") sp.geometry_data = std::stoi(v);
    else if (k == "mapping_degree") sp.mapping_degree = std::stoi(v);
    else if (k == "threads") sp.threads = std::stoi(v);
    else if (k == "radial_factor") sp.radial_factor = std::stoi(v);
    else if (k == "constraints") sp.constraints = std::stoi(v);
    else if (k == "n_ranks") sp.n_ranks = std::stoi(v);
    else if (k == "renumber") sp.renumber = v;
    else if (k == "rank") sp.rank = std::stoi(v);
    else throw std::runtime_error("unknown spec key: " + k);
  }
  return sp;
}

// TemperatureInitialValues<3> / Cuboid (include/model_data/boussinesq_model_data.tpp:57-147,168-196)
inline double temperature_initial_shell(int dim, double R0, double R1, const double* p) {
  const double cov = 20.0 / ((R1 - R0) / 2.0);
  double c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
  if (dim == 3) {
    c1[0] = R0 + (R1 - R0) * 0.35;
    c2[1] = R0 + (R1 - R0) * 0.65;
  } else {
    // 2-D: centres are R*c*R^T-"rotated" (quirk Q15: applied as rotation*(c*rotation^T))
    const double a = M_PI / 3.0;
    double R[2][2] = {{std::cos(a), -std::sin(a)}, {std::sin(a), std::cos(a)}};
    double t1[2] = {R0 + (R1 - R0) * 0.35, 0.0}, t2[2] = {0.0, R0 + (R1 - R0) * 0.65};
    // (rotation * c) is a vector v; v * transpose(rotation) contracts v with first index of R^T: sum_i v_i R^T[i][j] = sum_i v_i R[j][i]
    for (int pass = 0; pass < 2; ++pass) {
      double* t = pass == 0 ? t1 : t2;
      double* c = pass == 0 ? c1 : c2;
      double v[2] = {R[0][0] * t[0] + R[0][1] * t[1], R[1][0] * t[0] + R[1][1] * t[1]};
      c[0] = v[0] * R[0][0] + v[1] * R[0][1];
      c[1] = v[0] * R[1][0] + v[1] * R[1][1];
    }
  }
  double det = std::pow(cov, dim);
  double q1 = 0, q2 = 0;
  for (int d = 0; d < dim; ++d) {
    q1 += cov * (p[d] - c1[d]) * (p[d] - c1[d]);
    q2 += cov * (p[d] - c2[d]) * (p[d] - c2[d]);
  }
  double nrm = std::sqrt(std::pow(2.0 * M_PI, dim));
  return std::sqrt(det) * std::exp(-0.5 * q1) / nrm + std::sqrt(det) * std::exp(-0.5 * q2) / nrm;
}
inline double temperature_initial_cuboid(int dim, const double* center, double diameter, const double* p) {
  const double cov = 1.0 / ((diameter * 0.1) * (diameter * 0.1));
  double q = 0;
  for (int d = 0; d < dim; ++d) q += cov * (p[d] - center[d]) * (p[d] - center[d]);
  return std::sqrt(std::pow(cov, dim)) * std::exp(-0.5 * q) / (2.0 * std::sqrt(std::pow(2.0 * M_PI, 2)));
}

// Per-cell mapping evaluation -------------------------------------------------------------------
struct CellMapper {
  const Mesh& mesh;
  int dim, mdeg;
  std::vector<double> gl;  // GL nodes for the high-order mapping
  CellMapper(const Mesh& m, int mapping_degree) : mesh(m), dim(m.dim), mdeg(mapping_degree) {
    gl = gauss_lobatto01(mdeg);
  }
  // number of support points for cell c and fills X[ns][dim]; returns mapping degree used
  int support_points(int64_t c, double* X) const {
    int nv = 1 << dim;
    if (mdeg == 1 || !mesh.at_boundary(c)) {
      mesh.cell_vertices(c, X);
      (void)nv;
      return 1;
    }
    int n1 = mdeg + 1;
    int ns = dim == 3 ? n1 * n1 * n1 : n1 * n1;
    for (int s = 0; s < ns; ++s) {
      double xi[3] = {gl[s % n1], gl[(s / n1) % n1], dim == 3 ? gl[s / (n1 * n1)] : 0.0};
      mesh.manifold_point(c, xi, &X[s * dim]);
    }
    return mdeg;
  }
};

// geometry record per cell: [JxW(nq) | Kinv[e][d](nq each) | xq[d](nq each)], Kinv[e][d] = d xi_e / d x_d
inline int geom_stride(int dim, int nq) { return nq * (1 + dim * dim + dim); }

inline void invert_jac(int dim, const double* J, double* K, double& det) {
  if (dim == 2) {
    det = J[0] * J[3] - J[1] * J[2];
    double id = 1.0 / det;
    K[0] = J[3] * id;
    K[1] = -J[1] * id;
    K[2] = -J[2] * id;
    K[3] = J[0] * id;
    return;
  }
  double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
  det = J[0] * c00 + J[1] * c01 + J[2] * c02;
  double id = 1.0 / det;
  K[0] = c00 * id;
  K[1] = (J[2] * J[7] - J[1] * J[8]) * id;
  K[2] = (J[1] * J[5] - J[2] * J[4]) * id;
  K[3] = c01 * id;
  K[4] = (J[0] * J[8] - J[2] * J[6]) * id;
  K[5] = (J[2] * J[3] - J[0] * J[5]) * id;
  K[6] = c02 * id;
  K[7] = (J[1] * J[6] - J[0] * J[7]) * id;
  K[8] = (J[0] * J[4] - J[1] * J[3]) * id;
}

// J[i][j] = d x_i / d xi_j from support points X[ns][dim] and basis gradients dN[ns][dim]
inline void jacobian_from(int dim, int ns, const double* X, const double* dN, double* J) {
  for (int i = 0; i < dim * dim; ++i) J[i] = 0;
  for (int s = 0; s < ns; ++s)
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j) J[i * dim + j] += X[s * dim + i] * dN[s * dim + j];
}

inline void compute_geometry(const Mesh& mesh, int mapping_degree, const QuadRule& q, std::vector<double>& out,
                             bool want_full_jac = false, std::vector<double>* jac_out = nullptr) {
  const int dim = mesh.dim, nq = q.nq;
  MappingTable t1 = tabulate_mapping(dim, 1, q);
  MappingTable tm = mapping_degree > 1 ? tabulate_mapping(dim, mapping_degree, q) : t1;
  CellMapper mapper(mesh, mapping_degree);
  const int stride = geom_stride(dim, nq);
  out.resize((size_t)mesh.n_cells * stride);
  if (want_full_jac && jac_out) jac_out->resize((size_t)mesh.n_cells * nq * (dim * dim + 1));
#pragma omp parallel
  {
    std::vector<double> X(64 * 3);
#pragma omp for schedule(dynamic, 256)
    for (int64_t c = 0; c < mesh.n_cells; ++c) {
      int deg = mapper.support_points(c, X.data());
      const MappingTable& t = deg == 1 ? t1 : tm;
      double* g = &out[(size_t)c * stride];
      for (int iq = 0; iq < nq; ++iq) {
        double J[9], K[9], det;
        jacobian_from(dim, t.ns, X.data(), &t.dN[(size_t)iq * t.ns * dim], J);
        invert_jac(dim, J, K, det);
        g[iq] = det * q.w[iq];
        for (int e = 0; e < dim * dim; ++e) g[nq * (1 + e) + iq] = K[e];
        for (int d = 0; d < dim; ++d) {
          double x = 0;
          for (int s = 0; s < t.ns; ++s) x += t.N[(size_t)iq * t.ns + s] * X[s * dim + d];
          g[nq * (1 + dim * dim + d) + iq] = x;
        }
        if (want_full_jac && jac_out) {
          double* jo = &(*jac_out)[((size_t)c * nq + iq) * (dim * dim + 1)];
          for (int e = 0; e < dim * dim; ++e) jo[e] = J[e];
          jo[dim * dim] = det;
        }
      }
    }
  }
}

// ---- lowest-order FEEC reference shape functions (deal.II 9.2 FE_Nedelec(0), FE_RaviartThomas(0), FE_DGQ(0);
// restated from memory): Nedelec: unit tangential component along its own line, oriented
// along the positive axis, phi_l = prod_{e != d} h_e(x_e) e_d ; RT: phi_f = h(x_d) e_d (normal component
// measured along the POSITIVE axis on both faces of a pair), div = -1 / +1 ; DGQ0: 1.
struct FeecTable {
  int nq = 0;
  std::vector<double> phi_w, curl_w, phi_u;  // [nq][12][3], [nq][12][3], [nq][6][3]
  std::vector<double> div_u;                 // [6]
};
inline FeecTable tabulate_feec(const QuadRule& q) {
  FeecTable t;
  t.nq = q.nq;
  t.phi_w.assign((size_t)q.nq * 36, 0.0);
  t.curl_w.assign((size_t)q.nq * 36, 0.0);
  t.phi_u.assign((size_t)q.nq * 18, 0.0);
  t.div_u.assign(6, 0.0);
  auto offs = hierarchical_offsets(3);
  auto h = [](int o, double x) { return o == 0 ? 1.0 - x : x; };
  auto dh = [](int o) { return o == 0 ? -1.0 : 1.0; };
  for (int iq = 0; iq < q.nq; ++iq) {
    const double* x = &q.pts[iq * 3];
    for (int l = 0; l < 12; ++l) {
      const auto& o = offs[8 + l];
      int d = o[0] == 1 ? 0 : (o[1] == 1 ? 1 : 2);
      int e1 = (d + 1) % 3, e2 = (d + 2) % 3;
      double f = h(o[e1], x[e1]) * h(o[e2], x[e2]);
      double grad[3] = {0, 0, 0};
      grad[e1] = dh(o[e1]) * h(o[e2], x[e2]);
      grad[e2] = h(o[e1], x[e1]) * dh(o[e2]);
      t.phi_w[((size_t)iq * 12 + l) * 3 + d] = f;
      // curl(f e_d) = grad f x e_d : component e1 = -(d_e2 f) ... with (d,e1,e2) cyclic:
      //   (grad f x e_d)_{e1} = grad_{e2} f,  (grad f x e_d)_{e2} = -grad_{e1} f
      t.curl_w[((size_t)iq * 12 + l) * 3 + e1] = grad[e2];
      t.curl_w[((size_t)iq * 12 + l) * 3 + e2] = -grad[e1];
    }
    for (int f = 0; f < 6; ++f) {
      int d = f / 2, side = f % 2;
      t.phi_u[((size_t)iq * 6 + f) * 3 + d] = side ? x[d] : 1.0 - x[d];
    }
  }
  for (int f = 0; f < 6; ++f) t.div_u[f] = (f % 2) ? 1.0 : -1.0;
  return t;
}

// extended record for the Piola transforms: [JxW | Kinv[e][d] | xq[d] | J[i][j] | detJ], nq entries each
inline int geom_stride_ext(int dim, int nq) { return nq * (1 + dim * dim + dim + dim * dim + 1); }
inline void compute_geometry_ext(const Mesh& mesh, int mapping_degree, const QuadRule& q, std::vector<double>& out) {
  const int dim = mesh.dim, nq = q.nq;
  std::vector<double> base, jac;
  compute_geometry(mesh, mapping_degree, q, base, true, &jac);
  const int s0 = geom_stride(dim, nq), s1 = geom_stride_ext(dim, nq), nj = dim * dim + 1;
  out.resize((size_t)mesh.n_cells * s1);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < mesh.n_cells; ++c) {
    double* g = &out[(size_t)c * s1];
    std::copy(&base[(size_t)c * s0], &base[(size_t)c * s0] + s0, g);
    for (int iq = 0; iq < nq; ++iq)
      for (int e = 0; e < nj; ++e) g[s0 + e * nq + iq] = jac[((size_t)c * nq + iq) * nj + e];
  }
}


inline void collect_cell_vertices(const Mesh& mesh, std::vector<double>& out) {
  const int nv = 1 << mesh.dim;
  out.resize((size_t)mesh.n_cells * nv * mesh.dim);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < mesh.n_cells; ++c) mesh.cell_vertices(c, &out[(size_t)c * nv * mesh.dim]);
}

struct Problem {
  Spec spec;
  std::unique_ptr<Mesh> base_mesh;  // whole mesh (only when partitioned)
  std::unique_ptr<Mesh> mesh;       // this rank's cells: owned chunk first, then the ghost layer
  std::vector<int64_t> cell_global;
  std::vector<int32_t> node_owner;
  int64_t n_owned_cells = 0;
  DofMap nse, temp;
  Constraints nse_cs, temp_cs;
  std::vector<int64_t> nse_block_start;
  // reference tables
  QuadRule q_nse, q_temp;
  ScalarTable tab_u_qn, tab_p_qn, tab_t_qn;  // on the NSE rule (system, preconditioner)
  ScalarTable tab_u_qt, tab_t_qt;            // on the temperature rule (T matrices, T rhs)
  std::vector<double> geom_qn, geom_qt;
  std::vector<double> cell_vertices;  // [n_cells][2^dim][dim]
  // input of the device-side mapping evaluation (dcp_geometry_create)
  std::vector<int64_t> map_ptr;       // [n_cells+1] offsets into map_points
  std::vector<double> map_points;     // support points of each cell's mapping
  std::map<std::string, MappingTable> map_tables;
  std::map<std::string, std::vector<double>> map_weights;
  // patterns
  Csr nse_full, pre_full, temp_pat;
  Csr nse_b[2][2], pre_b[2][2];
  RowAdjacency nse_adj, temp_adj;
  // dof meta
  std::vector<double> nse_dof_xyz, temp_dof_xyz;
  std::vector<int8_t> nse_dof_comp;
  std::vector<int32_t> nse_coupling, pre_coupling;
  // FEEC family
  QuadRule q_pre;
  FeecTable feec_qn, feec_qp, feec_qt;
  std::vector<double> geom_qp, nse_sign;
  Csr nse_b3[3][3], pre_b3[3][3];
  std::map<std::string, ArrayRef> arrays;
  std::map<std::string, int64_t> scalars;

  template <class T>
  void reg(const std::string& name, const std::vector<T>& v, int dtype) {
    arrays[name] = ArrayRef{v.data(), (int64_t)v.size(), dtype};
  }
  void reg_csr(const std::string& name, const Csr& A) {
    reg(name + ".rowptr", A.rowptr, I64);
    reg(name + ".col", A.col, I32);
    scalars[name + ".n_rows"] = A.n_rows;
    scalars[name + ".n_cols"] = A.n_cols;
    scalars[name + ".nnz"] = A.rowptr.empty() ? 0 : A.rowptr.back();
  }
  void reg_cs(const std::string& name, const Constraints& c) {
    reg(name + ".line_dof", c.line_dof, I32);
    reg(name + ".line_ptr", c.line_ptr, I32);
    reg(name + ".entry_dof", c.entry_dof, I32);
    reg(name + ".entry_w", c.entry_w, F64);
    reg(name + ".inhom", c.inhom, F64);
    reg(name + ".line_of_dof", c.line_of_dof, I32);
  }
  // mapping support points of every cell and the mapping basis on one quadrature rule
  void collect_mapping(const Mesh& mesh, int mapping_degree) {
    CellMapper mapper(mesh, mapping_degree);
    const int dim = mesh.dim, nlow = 1 << dim;
    int nhigh = 1;
    for (int d = 0; d < dim; ++d) nhigh *= mapping_degree + 1;
    map_ptr.assign((size_t)mesh.n_cells + 1, 0);
    for (int64_t c = 0; c < mesh.n_cells; ++c)
      map_ptr[c + 1] = map_ptr[c] + ((mapping_degree == 1 || !mesh.at_boundary(c)) ? nlow : nhigh);
    map_points.resize((size_t)map_ptr.back() * dim);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t c = 0; c < mesh.n_cells; ++c) mapper.support_points(c, &map_points[(size_t)map_ptr[c] * dim]);
    reg("map.ptr", map_ptr, I64);
    reg("map.points", map_points, F64);
    scalars["map.n_low"] = nlow;
    scalars["map.n_high"] = nhigh;
  }
  void reg_map(const std::string& name, int dim, int mapping_degree, const QuadRule& q) {
    map_tables[name + ".low"] = tabulate_mapping(dim, 1, q);
    map_tables[name + ".high"] = tabulate_mapping(dim, std::max(1, mapping_degree), q);
    map_weights[name] = q.w;
    reg(name + ".N_low", map_tables[name + ".low"].N, F64);
    reg(name + ".dN_low", map_tables[name + ".low"].dN, F64);
    reg(name + ".N_high", map_tables[name + ".high"].N, F64);
    reg(name + ".dN_high", map_tables[name + ".high"].dN, F64);
    reg(name + ".w", map_weights[name], F64);
  }
  void reg_tab(const std::string& name, const ScalarTable& t) {
    reg(name + ".phi", t.phi, F64);
    reg(name + ".dphi", t.dphi, F64);
    scalars[name + ".nd"] = t.nd;
    scalars[name + ".nq"] = t.nq;
  }
};

// positions of all dofs of a dofmap (support points under the cell mapping) and their component
inline void dof_positions(const Mesh& mesh, const DofMap& dm, int mapping_degree, std::vector<double>& xyz,
                          std::vector<int8_t>* comp) {
  const int dim = mesh.dim, nl = dm.fe.n_local;
  xyz.assign((size_t)dm.n_dofs * dim, 0.0);
  if (comp) comp->assign((size_t)dm.n_dofs, 0);
  CellMapper mapper(mesh, mapping_degree);
  auto gl = gauss_lobatto01(mapping_degree), g1 = gauss_lobatto01(1);
  // the first cell (in cell order) that holds a dof defines its support point -> deterministic output
  std::vector<int32_t> owner((size_t)dm.n_dofs, -1);
  for (int64_t c = 0; c < mesh.n_cells; ++c)
    for (int i = 0; i < nl; ++i) {
      int32_t g = dm.l2g[(size_t)c * nl + i];
      if (owner[g] < 0) owner[g] = (int32_t)c;
    }
#pragma omp parallel
  {
    std::vector<double> X(64 * 3), N(64);
#pragma omp for schedule(dynamic, 256)
    for (int64_t c = 0; c < mesh.n_cells; ++c) {
      int deg = mapper.support_points(c, X.data());
      const auto& nodes = deg == 1 ? g1 : gl;
      int n1 = deg + 1, ns = dim == 3 ? n1 * n1 * n1 : n1 * n1;
      for (int i = 0; i < nl; ++i) {
        int32_t g = dm.l2g[(size_t)c * nl + i];
        if (owner[g] != (int32_t)c) continue;
        int lx = dm.fe.local_lex[i];
        double xi[3] = {0.5 * (lx % 3), 0.5 * ((lx / 3) % 3), 0.5 * (lx / 9)};
        mapping_basis_at(dim, nodes, xi, N.data(), nullptr);
        for (int d = 0; d < dim; ++d) {
          double x = 0;
          for (int s = 0; s < ns; ++s) x += N[s] * X[s * dim + d];
          xyz[(size_t)g * dim + d] = x;
        }
        if (comp) (*comp)[g] = (int8_t)dm.fe.local_field[i];
      }
    }
  }
}

inline void add_dirichlet(const Mesh& mesh, const DofMap& dm, int bid, const std::vector<int>& fields,
                          const std::function<double(const double*)>& value, const std::vector<double>& xyz,
                          Constraints& cs) {
  const int dim = mesh.dim, nl = dm.fe.n_local, nfaces = 2 * dim;
  for (int64_t c = 0; c < mesh.n_cells; ++c)
    for (int f = 0; f < nfaces; ++f) {
      if (mesh.face_boundary_id(c, f) != bid) continue;
      int d = f / 2, side = (f % 2) * 2;
      for (int i = 0; i < nl; ++i) {
        int lx = dm.fe.local_lex[i];
        int o[3] = {lx % 3, (lx / 3) % 3, lx / 9};
        if (o[d] != side) continue;
        if (std::find(fields.begin(), fields.end(), dm.fe.local_field[i]) == fields.end()) continue;
        int32_t g = dm.l2g[(size_t)c * nl + i];
        cs.add_line(g, {}, value ? value(&xyz[(size_t)g * dim]) : 0.0);
      }
    }
}

inline void add_periodic(const Mesh& mesh, const DofMap& dm, Constraints& cs) {
  const int dim = mesh.dim;
  auto offs = hierarchical_offsets(dim);
  const int n3 = dim == 3 ? 27 : 9;
  std::vector<int64_t> ids(n3);
  const int nf = (int)dm.fe.field_degree.size();
  for (int64_t c = 0; c < mesh.n_cells; ++c) {
    mesh.cell_nodes(c, ids.data());
    for (auto& o : offs) {
      int lx = lex_index(dim, o), ed = entity_dim(dim, o);
      int64_t m = mesh.periodic_master(ids[lx]);
      if (m < 0) continue;
      for (int f = 0; f < nf; ++f) {
        int32_t s = dm.dof_at(ids[lx], ed, f), mm = dm.dof_at(m, ed, f);
        if (s < 0 || mm < 0) continue;
        cs.add_line(s, {{mm, 1.0}}, 0.0);
      }
    }
  }
}

// compute_no_normal_flux_constraints on boundary `bid` for the vector field starting at component 0
// `mesh` is the mesh the normals are averaged over: for a partitioned problem this is the WHOLE mesh, so that the
// constraint of a ghost dof does not depend on which of its faces happen to be local (deal.II needs
// AffineConstraints::is_consistent_in_parallel for the same reason); only nodes that carry local dofs get lines.
inline void add_no_normal_flux(const Mesh& mesh, const DofMap& dm, int bid, int mapping_degree, Constraints& cs) {
  const int dim = mesh.dim, nfaces = 2 * dim;
  auto offs = hierarchical_offsets(dim);
  const int n3 = dim == 3 ? 27 : 9;
  CellMapper mapper(mesh, mapping_degree);
  auto gl = gauss_lobatto01(mapping_degree), g1 = gauss_lobatto01(1);
  std::map<int64_t, std::array<double, 3>> normal_sum;  // lattice node -> summed unit normals
  std::map<int64_t, int> node_ed;
  std::vector<int64_t> ids(n3);
  std::vector<double> X(64 * 3), dN(64 * 3);
  for (int64_t c = 0; c < mesh.n_cells; ++c)
    for (int f = 0; f < nfaces; ++f) {
      if (mesh.face_boundary_id(c, f) != bid) continue;
      int fd = f / 2, side = (f % 2) * 2;
      mesh.cell_nodes(c, ids.data());
      int deg = mapper.support_points(c, X.data());
      const auto& nodes = deg == 1 ? g1 : gl;
      int n1 = deg + 1, ns = dim == 3 ? n1 * n1 * n1 : n1 * n1;
      for (auto& o : offs) {
        if (o[fd] != side) continue;
        double xi[3] = {0.5 * o[0], 0.5 * o[1], 0.5 * o[2]};
        mapping_basis_at(dim, nodes, xi, nullptr, dN.data());
        double J[9], K[9], det;
        jacobian_from(dim, ns, X.data(), dN.data(), J);
        invert_jac(dim, J, K, det);
        // n ~ J^{-T} n_hat, n_hat = +-e_fd  ->  n_i = K[fd][i] * sign
        double nn[3] = {0, 0, 0}, len = 0;
        for (int i = 0; i < dim; ++i) {
          nn[i] = K[fd * dim + i] * (side ? 1.0 : -1.0);
          len += nn[i] * nn[i];
        }
        len = std::sqrt(len);
        int64_t node = ids[lex_index(dim, o)];
        auto& acc = normal_sum[node];
        for (int i = 0; i < dim; ++i) acc[i] += nn[i] / len;
        node_ed[node] = entity_dim(dim, o);
      }
    }
  const double eps = std::numeric_limits<double>::epsilon();
  for (auto& kv : normal_sum) {
    double n[3] = {kv.second[0], kv.second[1], kv.second[2]};
    double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int i = 0; i < 3; ++i) n[i] /= len;
    int k = 0;
    for (int i = 1; i < dim; ++i)
      if (std::fabs(n[i]) > std::fabs(n[k])) k = i;
    int ed = node_ed[kv.first];
    if (dm.node_first[kv.first] < 0) continue;  // node not on this rank
    int32_t dk = dm.dof_at(kv.first, ed, k);
    std::vector<std::pair<int32_t, double>> e;
    for (int i = 0; i < dim; ++i)
      if (i != k && std::fabs(n[i] / n[k]) > eps) e.push_back({dm.dof_at(kv.first, ed, i), -n[i] / n[k]});
    cs.add_line(dk, e, 0.0);
  }
}

// FEEC model: FESystem(FE_Nedelec(0), FE_RaviartThomas(0), FE_DGQ(0)) + Q1 temperature
// (include/core/boussineq_model_FEEC.tpp:14-55, setup_dofs :231-478, setup_nse_matrices :79-130,
// setup_nse_preconditioner :133-198).  Deviations that cannot matter to the kernels: no Cuthill-McKee pass
// (:243, tie-breaking unverifiable) -- dofs keep first-visit order inside each block (block_wise, :245).
inline void build_feec(Problem* P, const Spec& sp) {
  const Mesh& mesh = *P->mesh;
  const int dim = 3;
  const bool cuboid = sp.geometry == "cube";
  const std::vector<int32_t>* owner = sp.n_ranks > 1 ? &P->node_owner : nullptr;
  if (sp.renumber != "none" && sp.n_ranks > 1) throw std::runtime_error("harness: renumber is implemented for a single rank");
  FESystemDesc fe;
  fe.dim = dim;
  fe.field_degree = {0, 0, 0};
  fe.field_block = {0, 1, 2};
  fe.field_kind = {1, 2, 3};
  P->nse = distribute_dofs(mesh, fe, owner, sp.rank, sp.renumber);
  FESystemDesc t_fe;
  t_fe.dim = dim;
  t_fe.field_degree = {sp.temperature_degree};
  t_fe.field_block = {0};
  P->temp = distribute_dofs(mesh, t_fe, owner, sp.rank);
  P->nse_block_start = {0, P->nse.block_size[0], P->nse.block_size[0] + P->nse.block_size[1],
                        P->nse.block_size[0] + P->nse.block_size[1] + P->nse.block_size[2]};
  const int mdeg = 1;  // FEValues without a mapping argument (boussineq_model_assembly_FEEC.tpp:24,67,160); T: MappingQ(1)
  dof_positions(mesh, P->nse, mdeg, P->nse_dof_xyz, &P->nse_dof_comp);
  dof_positions(mesh, P->temp, mdeg, P->temp_dof_xyz, nullptr);

  // constraints (:308-478): zero tangential (Nedelec) and zero normal (RT) boundary dofs on both boundaries
  if (sp.constraints) {
    if (cuboid) {
      add_periodic(mesh, P->nse, P->nse_cs);
      for (int bid : {4, 5}) add_dirichlet(mesh, P->nse, bid, {0, 1}, nullptr, P->nse_dof_xyz, P->nse_cs);
      add_periodic(mesh, P->temp, P->temp_cs);
      double center[3] = {0.5, 0.5, 0.5};
      double diam = std::sqrt(3.0);
      add_dirichlet(mesh, P->temp, 4, {0}, [&](const double* p) { return temperature_initial_cuboid(dim, center, diam, p); },
                    P->temp_dof_xyz, P->temp_cs);
    } else {
      for (int bid : {0, 1}) add_dirichlet(mesh, P->nse, bid, {0, 1}, nullptr, P->nse_dof_xyz, P->nse_cs);
      add_dirichlet(mesh, P->temp, 0, {0}, [&](const double* p) { return temperature_initial_shell(dim, sp.R0, sp.R1, p); },
                    P->temp_dof_xyz, P->temp_cs);
    }
  }
  P->nse_cs.close(P->nse.n_dofs);
  P->temp_cs.close(P->temp.n_dofs);

  // face sign (source/base/utilities.cc:20-46): -1 on the RT dof of an interior face the cell sees in
  // non-standard orientation.  Here: the first cell (in cell order) that meets a face defines its normal; a
  // later cell that sees the face on the same side (both as a "0" face or both as a "1" face, which only
  // happens across the seams of the six trees) has the opposite local normal.
  {
    const int nl = P->nse.fe.n_local;
    P->nse_sign.assign((size_t)mesh.n_cells * nl, 1.0);
    auto offs = hierarchical_offsets(dim);
    // "First" is decided on the whole mesh in global cell order, so that every rank of a partition (whose local
    // cell order is owned cells, then ghosts) sees the same sign on a shared face.
    const Mesh& whole = P->base_mesh ? *P->base_mesh : mesh;
    std::vector<int8_t> first_side((size_t)whole.n_nodes, -1);
    std::vector<int64_t> first_cell((size_t)whole.n_nodes, -1);
    int64_t ids[27];
    for (int64_t c = 0; c < whole.n_cells; ++c) {
      whole.cell_nodes(c, ids);
      for (int f = 0; f < 6; ++f) {
        if (whole.face_boundary_id(c, f) >= 0) continue;
        int64_t node = ids[lex_index(dim, offs[20 + f])];
        if (first_side[node] < 0) {
          first_side[node] = (int8_t)(f % 2);
          first_cell[node] = c;
        }
      }
    }
    for (int64_t c = 0; c < mesh.n_cells; ++c) {
      const int64_t gc = P->base_mesh ? P->cell_global[(size_t)c] : c;
      mesh.cell_nodes(c, ids);
      for (int f = 0; f < 6; ++f) {
        if (mesh.face_boundary_id(c, f) >= 0) continue;
        int64_t node = ids[lex_index(dim, offs[20 + f])];
        if (first_cell[node] != gc && first_side[node] == f % 2) P->nse_sign[(size_t)c * nl + 12 + f] = -1.0;
      }
    }
  }

  // rules: system QGauss(deg+2) (:843), preconditioner QGauss(deg+1) (:595), temperature QGauss(Tdeg+2) (:969,1125)
  const int deg = 1;  // "nse velocity degree" of the FEEC configs
  P->q_nse = qgauss(dim, deg + 2);
  P->q_pre = qgauss(dim, deg + 1);
  P->q_temp = qgauss(dim, sp.temperature_degree + 2);
  P->feec_qn = tabulate_feec(P->q_nse);
  P->feec_qp = tabulate_feec(P->q_pre);
  P->feec_qt = tabulate_feec(P->q_temp);
  P->tab_t_qn = tabulate_scalar(dim, sp.temperature_degree, P->q_nse);
  P->tab_t_qt = tabulate_scalar(dim, sp.temperature_degree, P->q_temp);
  if (sp.geometry_data) {
    compute_geometry_ext(mesh, mdeg, P->q_nse, P->geom_qn);
    compute_geometry_ext(mesh, mdeg, P->q_pre, P->geom_qp);
    if (P->q_temp.n1 != P->q_nse.n1) compute_geometry_ext(mesh, mdeg, P->q_temp, P->geom_qt);
  }
  // block couplings through the first nonzero component of each (non-primitive) shape function
  // (setup_nse_matrices :94-118, setup_nse_preconditioner :151-184): blocks w=0, u=1, p=2
  P->nse_coupling = {1, 1, 0, 1, 1, 1, 0, 1, 0};
  P->pre_coupling = {1, 1, 0, 1, 1, 0, 1, 0, 1};
  P->nse_adj = build_row_adjacency(P->nse, P->nse_cs);
  P->temp_adj = build_row_adjacency(P->temp, P->temp_cs);
  if (sp.patterns) {
    std::vector<int> cn(P->nse_coupling.begin(), P->nse_coupling.end()), cp(P->pre_coupling.begin(), P->pre_coupling.end());
    P->nse_full = make_sparsity_pattern(P->nse, P->nse_cs, cn, P->nse_adj);
    P->pre_full = make_sparsity_pattern(P->nse, P->nse_cs, cp, P->nse_adj);
    P->temp_pat = make_sparsity_pattern(P->temp, P->temp_cs, {1}, P->temp_adj);
    for (int bi = 0; bi < 3; ++bi)
      for (int bj = 0; bj < 3; ++bj) {
        P->nse_b3[bi][bj] = extract_block(P->nse_full, P->nse_block_start, bi, bj);
        P->pre_b3[bi][bj] = extract_block(P->pre_full, P->nse_block_start, bi, bj);
      }
  }
  auto& S = P->scalars;
  S["dim"] = dim;
  S["feec"] = 1;
  S["n_cells"] = mesh.n_cells;
  S["n_owned_cells"] = P->n_owned_cells;
  S["n_ranks"] = sp.n_ranks;
  S["rank"] = sp.rank;
  S["nse.n_w_owned"] = P->nse.owned_size[0];
  S["nse.n_u_owned"] = P->nse.owned_size[1];
  S["nse.n_p_owned"] = P->nse.owned_size[2];
  S["nse.n_dofs"] = P->nse.n_dofs;
  S["nse.n_w"] = P->nse.block_size[0];
  S["nse.n_u"] = P->nse.block_size[1];
  S["nse.n_p"] = P->nse.block_size[2];
  S["nse.n_local"] = P->nse.fe.n_local;
  S["temp.n_dofs"] = P->temp.n_dofs;
  S["temp.n_local"] = P->temp.fe.n_local;
  S["temp.n_owned"] = P->temp.owned_size[0];
  S["q_nse.nq"] = P->q_nse.nq;
  S["q_pre.nq"] = P->q_pre.nq;
  S["q_temp.nq"] = P->q_temp.nq;
  S["geom_shared"] = (P->q_temp.n1 == P->q_nse.n1) ? 1 : 0;
  S["cuboid"] = cuboid ? 1 : 0;
  P->reg("nse.l2g", P->nse.l2g, I32);
  P->reg("temp.l2g", P->temp.l2g, I32);
  P->reg("nse.local_field", P->nse.fe.local_field, I32);
  P->reg("nse.local_base", P->nse.fe.local_base, I32);
  P->reg("nse.sign", P->nse_sign, F64);
  P->reg("nse.dof_xyz", P->nse_dof_xyz, F64);
  P->reg("temp.dof_xyz", P->temp_dof_xyz, F64);
  P->reg("nse.dof_comp", P->nse_dof_comp, I8);
  P->reg("nse.dof_key", P->nse.dof_key, I64);
  P->reg("temp.dof_key", P->temp.dof_key, I64);
  P->reg("nse.dof_owner", P->nse.dof_owner, I32);
  P->reg("temp.dof_owner", P->temp.dof_owner, I32);
  P->reg("cell_global", P->cell_global, I64);
  P->reg_cs("nse.cs", P->nse_cs);
  P->reg_cs("temp.cs", P->temp_cs);
  P->reg("q_nse.w", P->q_nse.w, F64);
  P->reg("q_pre.w", P->q_pre.w, F64);
  P->reg("q_temp.w", P->q_temp.w, F64);
  auto reg_feec = [&](const std::string& n, const FeecTable& t) {
    P->reg(n + ".phi_w", t.phi_w, F64);
    P->reg(n + ".curl_w", t.curl_w, F64);
    P->reg(n + ".phi_u", t.phi_u, F64);
    P->reg(n + ".div_u", t.div_u, F64);
  };
  reg_feec("feec.qn", P->feec_qn);
  reg_feec("feec.qp", P->feec_qp);
  reg_feec("feec.qt", P->feec_qt);
  P->reg_tab("tab.t_qn", P->tab_t_qn);
  P->reg_tab("tab.t_qt", P->tab_t_qt);
  collect_cell_vertices(mesh, P->cell_vertices);
  P->reg("cell_vertices", P->cell_vertices, F64);
  P->collect_mapping(mesh, mdeg);
  P->reg_map("map.qn", mesh.dim, mdeg, P->q_nse);
  P->reg_map("map.qt", mesh.dim, mdeg, P->q_temp);
  P->reg_map("map.qp", mesh.dim, mdeg, P->q_pre);
  P->reg("geom.qn", P->geom_qn, F64);
  P->reg("geom.qp", P->geom_qp, F64);
  P->reg("geom.qt", P->q_temp.n1 == P->q_nse.n1 ? P->geom_qn : P->geom_qt, F64);
  P->reg("nse.coupling", P->nse_coupling, I32);
  P->reg("pre.coupling", P->pre_coupling, I32);
  if (sp.patterns) {
    P->reg_csr("nse.full", P->nse_full);
    P->reg_csr("pre.full", P->pre_full);
    P->reg_csr("temp.pat", P->temp_pat);
    for (int bi = 0; bi < 3; ++bi)
      for (int bj = 0; bj < 3; ++bj) {
        std::string t = std::to_string(bi) + std::to_string(bj);
        P->reg_csr("nse.b" + t, P->nse_b3[bi][bj]);
        P->reg_csr("pre.b" + t, P->pre_b3[bi][bj]);
      }
  }
}

inline std::unique_ptr<Problem> build_problem(const Spec& sp) {
  if (sp.threads > 0) omp_set_num_threads(sp.threads);
  auto P = std::make_unique<Problem>();
  P->spec = sp;
  if (sp.dim != 3 && !(sp.dim == 2 && sp.geometry == "annulus" && sp.family == "classic"))
    throw std::runtime_error("harness: dim=2 is implemented for geometry=annulus, family=classic only");
  if (sp.family != "classic" && sp.family != "feec") throw std::runtime_error("harness: unknown family " + sp.family);
  const int dim = sp.dim;
  if (sp.geometry == "shell")
    P->mesh = std::make_unique<ShellMesh3D>(sp.refine, sp.R0, sp.R1, sp.radial_factor);
  else if (sp.geometry == "cube")
    P->mesh = std::make_unique<CubeMesh3D>(sp.refine, true);
  else if (sp.geometry == "annulus")
    P->mesh = std::make_unique<AnnulusMesh2D>(sp.refine, sp.R0, sp.R1);
  else
    throw std::runtime_error("harness: unknown geometry " + sp.geometry);
  P->n_owned_cells = P->mesh->n_cells;
  if (sp.n_ranks > 1) {
    // p4est-style partition (restated): rank p owns cells [floor(pN/P), floor((p+1)N/P))
    // of the space-filling curve; a lattice node (and its dofs) belongs to the lowest rank touching it; the
    // rank also sees the one-deep layer of ghost cells (cells sharing a node with an owned cell).
    if (sp.geometry != "shell") throw std::runtime_error("harness: only the shell can be partitioned");
    P->base_mesh = std::move(P->mesh);
    const Mesh& B = *P->base_mesh;
    const int64_t N = B.n_cells;
    auto rank_of = [&](int64_t c) {
      int p = (int)(((__int128)(c + 1) * sp.n_ranks - 1) / N);
      while ((int64_t)(((__int128)p * N) / sp.n_ranks) > c) --p;
      while ((int64_t)(((__int128)(p + 1) * N) / sp.n_ranks) <= c) ++p;
      return p;
    };
    const int64_t c0 = (int64_t)(((__int128)sp.rank * N) / sp.n_ranks), c1 = (int64_t)(((__int128)(sp.rank + 1) * N) / sp.n_ranks);
    P->node_owner.assign((size_t)B.n_nodes, INT32_MAX);
    std::vector<uint8_t> touched((size_t)B.n_nodes, 0);
    int64_t ids[27];
    for (int64_t c = 0; c < N; ++c) {
      B.cell_nodes(c, ids);
      const int rk = rank_of(c);
      for (int k = 0; k < 27; ++k) {
        if (P->node_owner[ids[k]] > rk) P->node_owner[ids[k]] = rk;
        if (c >= c0 && c < c1) touched[ids[k]] = 1;
      }
    }
    std::vector<int64_t> cells;
    for (int64_t c = c0; c < c1; ++c) cells.push_back(c);
    for (int64_t c = 0; c < N; ++c) {
      if (c >= c0 && c < c1) continue;
      B.cell_nodes(c, ids);
      bool g = false;
      for (int k = 0; k < 27 && !g; ++k) g = touched[ids[k]];
      if (g) cells.push_back(c);
    }
    P->n_owned_cells = c1 - c0;
    P->cell_global = cells;
    P->mesh = std::make_unique<SubMesh>(B, cells);
  }
  if (sp.family == "feec") {
    build_feec(P.get(), sp);
    return P;
  }
  const Mesh& mesh = *P->mesh;
  const bool cuboid = sp.geometry == "cube";

  // --- DoFs (boussinesq_model.tpp:191-206)
  FESystemDesc nse_fe;
  nse_fe.dim = dim;
  for (int d = 0; d < dim; ++d) {
    nse_fe.field_degree.push_back(sp.velocity_degree);
    nse_fe.field_block.push_back(0);
  }
  nse_fe.field_degree.push_back(sp.velocity_degree - 1);
  nse_fe.field_block.push_back(1);
  const std::vector<int32_t>* owner = sp.n_ranks > 1 ? &P->node_owner : nullptr;
  if (sp.renumber != "none" && sp.n_ranks > 1) throw std::runtime_error("harness: renumber is implemented for a single rank");
  P->nse = distribute_dofs(mesh, nse_fe, owner, sp.rank, sp.renumber);
  FESystemDesc t_fe;
  t_fe.dim = dim;
  t_fe.field_degree = {sp.temperature_degree};
  t_fe.field_block = {0};
  P->temp = distribute_dofs(mesh, t_fe, owner, sp.rank);
  P->nse_block_start = {0, P->nse.block_size[0], P->nse.block_size[0] + P->nse.block_size[1]};

  dof_positions(mesh, P->nse, sp.mapping_degree, P->nse_dof_xyz, &P->nse_dof_comp);
  dof_positions(mesh, P->temp, sp.mapping_degree, P->temp_dof_xyz, nullptr);

  // --- constraints (boussinesq_model.tpp:259-387)
  std::vector<int> vel;
  for (int d = 0; d < dim; ++d) vel.push_back(d);
  if (!sp.constraints) {
    // unconstrained spaces
  } else if (cuboid) {
    add_periodic(mesh, P->nse, P->nse_cs);
    add_dirichlet(mesh, P->nse, 4, vel, nullptr, P->nse_dof_xyz, P->nse_cs);
    add_no_normal_flux(mesh, P->nse, 5, sp.mapping_degree, P->nse_cs);
    add_periodic(mesh, P->temp, P->temp_cs);
    double center[3] = {0.5, 0.5, 0.5};
    double diam = std::sqrt(3.0);
    add_dirichlet(mesh, P->temp, 4, {0},
                  [&](const double* p) { return temperature_initial_cuboid(dim, center, diam, p); }, P->temp_dof_xyz,
                  P->temp_cs);
  } else {
    add_dirichlet(mesh, P->nse, 0, vel, nullptr, P->nse_dof_xyz, P->nse_cs);
    add_no_normal_flux(P->base_mesh ? *P->base_mesh : mesh, P->nse, 1, sp.mapping_degree, P->nse_cs);
    add_dirichlet(mesh, P->temp, 0, {0}, [&](const double* p) { return temperature_initial_shell(dim, sp.R0, sp.R1, p); },
                  P->temp_dof_xyz, P->temp_cs);
  }
  P->nse_cs.close(P->nse.n_dofs);
  P->temp_cs.close(P->temp.n_dofs);

  // --- reference tables
  P->q_nse = qgauss(dim, sp.velocity_degree + 1);
  P->q_temp = qgauss(dim, sp.temperature_degree + 2);
  P->tab_u_qn = tabulate_scalar(dim, sp.velocity_degree, P->q_nse);
  P->tab_p_qn = tabulate_scalar(dim, sp.velocity_degree - 1, P->q_nse);
  P->tab_t_qn = tabulate_scalar(dim, sp.temperature_degree, P->q_nse);
  P->tab_u_qt = tabulate_scalar(dim, sp.velocity_degree, P->q_temp);
  P->tab_t_qt = tabulate_scalar(dim, sp.temperature_degree, P->q_temp);

  // --- mapping data
  if (sp.geometry_data) {
    compute_geometry(mesh, sp.mapping_degree, P->q_nse, P->geom_qn);
    if (P->q_temp.n1 != P->q_nse.n1) compute_geometry(mesh, sp.mapping_degree, P->q_temp, P->geom_qt);
  }

  // --- coupling tables and patterns (boussinesq_model.tpp:90-105, 131-146, 164-174)
  const int nf = dim + 1;
  P->nse_coupling.assign(nf * nf, 1);
  P->nse_coupling[dim * nf + dim] = 0;
  P->pre_coupling.assign(nf * nf, 0);
  for (int c = 0; c < nf; ++c) P->pre_coupling[c * nf + c] = 1;
  P->nse_adj = build_row_adjacency(P->nse, P->nse_cs);
  P->temp_adj = build_row_adjacency(P->temp, P->temp_cs);
  if (sp.patterns) {
    std::vector<int> cn(P->nse_coupling.begin(), P->nse_coupling.end()), cp(P->pre_coupling.begin(), P->pre_coupling.end());
    P->nse_full = make_sparsity_pattern(P->nse, P->nse_cs, cn, P->nse_adj);
    P->pre_full = make_sparsity_pattern(P->nse, P->nse_cs, cp, P->nse_adj);
    P->temp_pat = make_sparsity_pattern(P->temp, P->temp_cs, {1}, P->temp_adj);
    for (int bi = 0; bi < 2; ++bi)
      for (int bj = 0; bj < 2; ++bj) {
        P->nse_b[bi][bj] = extract_block(P->nse_full, P->nse_block_start, bi, bj);
        P->pre_b[bi][bj] = extract_block(P->pre_full, P->nse_block_start, bi, bj);
      }
  }

  // --- registry
  auto& S = P->scalars;
  S["dim"] = dim;
  S["n_cells"] = mesh.n_cells;
  S["n_owned_cells"] = P->n_owned_cells;
  S["n_ranks"] = sp.n_ranks;
  S["rank"] = sp.rank;
  S["nse.n_u_owned"] = P->nse.owned_size[0];
  S["nse.n_p_owned"] = P->nse.owned_size[1];
  S["temp.n_owned"] = P->temp.owned_size[0];
  P->reg("cell_global", P->cell_global, I64);
  P->reg("nse.dof_key", P->nse.dof_key, I64);
  P->reg("nse.dof_owner", P->nse.dof_owner, I32);
  P->reg("temp.dof_key", P->temp.dof_key, I64);
  P->reg("temp.dof_owner", P->temp.dof_owner, I32);
  S["nse.n_dofs"] = P->nse.n_dofs;
  S["nse.n_u"] = P->nse.block_size[0];
  S["nse.n_p"] = P->nse.block_size[1];
  S["nse.n_local"] = P->nse.fe.n_local;
  S["temp.n_dofs"] = P->temp.n_dofs;
  S["temp.n_local"] = P->temp.fe.n_local;
  S["q_nse.nq"] = P->q_nse.nq;
  S["q_temp.nq"] = P->q_temp.nq;
  S["geom_shared"] = (P->q_temp.n1 == P->q_nse.n1) ? 1 : 0;
  S["cuboid"] = cuboid ? 1 : 0;
  P->reg("nse.l2g", P->nse.l2g, I32);
  P->reg("temp.l2g", P->temp.l2g, I32);
  P->reg("nse.local_field", P->nse.fe.local_field, I32);
  P->reg("nse.local_base", P->nse.fe.local_base, I32);
  P->reg("temp.local_base", P->temp.fe.local_base, I32);
  P->reg("nse.dof_xyz", P->nse_dof_xyz, F64);
  P->reg("temp.dof_xyz", P->temp_dof_xyz, F64);
  P->reg("nse.dof_comp", P->nse_dof_comp, I8);
  P->reg_cs("nse.cs", P->nse_cs);
  P->reg_cs("temp.cs", P->temp_cs);
  P->reg("q_nse.w", P->q_nse.w, F64);
  P->reg("q_nse.pts", P->q_nse.pts, F64);
  P->reg("q_temp.w", P->q_temp.w, F64);
  P->reg("q_temp.pts", P->q_temp.pts, F64);
  P->reg_tab("tab.u_qn", P->tab_u_qn);
  P->reg_tab("tab.p_qn", P->tab_p_qn);
  P->reg_tab("tab.t_qn", P->tab_t_qn);
  P->reg_tab("tab.u_qt", P->tab_u_qt);
  P->reg_tab("tab.t_qt", P->tab_t_qt);
  collect_cell_vertices(mesh, P->cell_vertices);
  P->reg("cell_vertices", P->cell_vertices, F64);
  P->collect_mapping(mesh, sp.mapping_degree);
  P->reg_map("map.qn", mesh.dim, sp.mapping_degree, P->q_nse);
  P->reg_map("map.qt", mesh.dim, sp.mapping_degree, P->q_temp);
  P->reg("geom.qn", P->geom_qn, F64);
  P->reg("geom.qt", P->q_temp.n1 == P->q_nse.n1 ? P->geom_qn : P->geom_qt, F64);
  P->reg("nse.coupling", P->nse_coupling, I32);
  P->reg("pre.coupling", P->pre_coupling, I32);
  P->reg("nse.adj.ptr", P->nse_adj.ptr, I64);
  P->reg("nse.adj.cell", P->nse_adj.cell, I32);
  P->reg("nse.adj.loc", P->nse_adj.loc, I16);
  P->reg("nse.adj.w", P->nse_adj.w, F64);
  if (sp.patterns) {
    P->reg_csr("nse.full", P->nse_full);
    P->reg_csr("pre.full", P->pre_full);
    P->reg_csr("temp.pat", P->temp_pat);
    for (int bi = 0; bi < 2; ++bi)
      for (int bj = 0; bj < 2; ++bj) {
        std::string s = std::to_string(bi) + std::to_string(bj);
        P->reg_csr("nse.b" + s, P->nse_b[bi][bj]);
        P->reg_csr("pre.b" + s, P->pre_b[bi][bj]);
      }
  }
  return P;
}

}  // namespace dcph
