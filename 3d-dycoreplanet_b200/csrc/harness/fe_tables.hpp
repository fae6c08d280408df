// Reference-cell tables for the stand-in harness: Gauss quadrature, Lagrange Q1/Q2 shape
// functions in deal.II's hierarchical cell-DoF order, and Q1/Q3 mapping bases.
//
// This file restates conventions of deal.II (un-vendored dependency of the reference, pinned only as
// ">= 9.2.0" in /root/reference/CMakeLists.txt:26): unit cell [0,1]^dim, QGauss<dim>(n) tensor rule
// with the x index running fastest, FE_Q nodes equidistant, cell DoFs ordered vertices -> lines ->
// quads -> hex.  Uses in the reference: QGauss(deg+1) boussinesq_model.tpp:487,708; QGauss(Tdeg+2)
// :834,990; FE_Q / FESystem :21-30; MappingQ(3) :20.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace dcph {

struct Rule1D {
  std::vector<double> x, w;
};

// Gauss-Legendre rule with n points mapped to [0,1] (Newton on P_n, then affine map).
inline Rule1D gauss01(int n) {
  Rule1D r;
  r.x.resize(n);
  r.w.resize(n);
  for (int i = 0; i < n; ++i) {
    long double z = std::cos(M_PIl * (i + 0.75L) / (n + 0.5L));
    long double pp = 0;
    for (int it = 0; it < 100; ++it) {
      long double p1 = 1, p2 = 0;
      for (int j = 0; j < n; ++j) {
        long double p3 = p2;
        p2 = p1;
        p1 = ((2 * j + 1) * z * p2 - j * p3) / (j + 1);
      }
      pp = n * (z * p1 - p2) / (z * z - 1);
      long double dz = p1 / pp;
      z -= dz;
      if (std::fabs((double)dz) < 1e-19) break;
    }
    // ascending order on [0,1]
    r.x[n - 1 - i] = (double)(0.5L * (1 + z));
    r.w[n - 1 - i] = (double)(1.0L / ((1 - z * z) * pp * pp));
  }
  // enforce exact symmetry
  for (int i = 0; i < n / 2; ++i) {
    double xm = 0.5 * (r.x[i] + (1.0 - r.x[n - 1 - i]));
    r.x[i] = xm;
    r.x[n - 1 - i] = 1.0 - xm;
    double wm = 0.5 * (r.w[i] + r.w[n - 1 - i]);
    r.w[i] = r.w[n - 1 - i] = wm;
  }
  if (n % 2) r.x[n / 2] = 0.5;
  return r;
}

// 1-D Lagrange basis on the half-step lattice {0, 1/2, 1}: `o` is the lattice offset (0,1,2).
// degree 1 uses offsets {0,2}; degree 2 uses {0,1,2}.
inline double lag1d(int degree, int o, double x) {
  if (degree == 1) return o == 0 ? 1.0 - x : x;
  switch (o) {
    case 0: return (1.0 - x) * (1.0 - 2.0 * x);
    case 1: return 4.0 * x * (1.0 - x);
    default: return x * (2.0 * x - 1.0);
  }
}
inline double dlag1d(int degree, int o, double x) {
  if (degree == 1) return o == 0 ? -1.0 : 1.0;
  switch (o) {
    case 0: return 4.0 * x - 3.0;
    case 1: return 4.0 - 8.0 * x;
    default: return 4.0 * x - 1.0;
  }
}

// Lattice offsets (ox,oy,oz in {0,1,2}) of the 3^dim local nodes in deal.II hierarchical order.
// 3-D: 8 vertices (lexicographic), 12 lines (0-3 bottom: x=0,x=1 along y; y=0,y=1 along x; 4-7 top;
// 8-11 vertical), 6 quads (x0,x1,y0,y1,z0,z1), 1 hex.  2-D: 4 vertices, 4 lines (x0,x1,y0,y1), 1 quad.
inline std::vector<std::array<int, 3>> hierarchical_offsets(int dim) {
  std::vector<std::array<int, 3>> o;
  if (dim == 3) {
    for (int v = 0; v < 8; ++v) o.push_back({2 * (v & 1), 2 * ((v >> 1) & 1), 2 * ((v >> 2) & 1)});
    for (int z = 0; z <= 2; z += 2) {
      o.push_back({0, 1, z});
      o.push_back({2, 1, z});
      o.push_back({1, 0, z});
      o.push_back({1, 2, z});
    }
    o.push_back({0, 0, 1});
    o.push_back({2, 0, 1});
    o.push_back({0, 2, 1});
    o.push_back({2, 2, 1});
    o.push_back({0, 1, 1});
    o.push_back({2, 1, 1});
    o.push_back({1, 0, 1});
    o.push_back({1, 2, 1});
    o.push_back({1, 1, 0});
    o.push_back({1, 1, 2});
    o.push_back({1, 1, 1});
  } else {
    for (int v = 0; v < 4; ++v) o.push_back({2 * (v & 1), 2 * ((v >> 1) & 1), 0});
    o.push_back({0, 1, 0});
    o.push_back({2, 1, 0});
    o.push_back({1, 0, 0});
    o.push_back({1, 2, 0});
    o.push_back({1, 1, 0});
  }
  return o;
}

inline int lex_index(int dim, const std::array<int, 3>& o) {
  return dim == 3 ? o[0] + 3 * (o[1] + 3 * o[2]) : o[0] + 3 * o[1];
}

// entity dimension of a lattice offset: number of odd coordinates (0 = vertex, 1 = line, ...)
inline int entity_dim(int dim, const std::array<int, 3>& o) {
  int e = 0;
  for (int d = 0; d < dim; ++d) e += (o[d] & 1);
  return e;
}

struct QuadRule {
  int dim = 0, n1 = 0, nq = 0;
  std::vector<double> pts;  // [nq][dim]
  std::vector<double> w;    // [nq]
};

inline QuadRule qgauss(int dim, int n1) {
  Rule1D g = gauss01(n1);
  QuadRule q;
  q.dim = dim;
  q.n1 = n1;
  q.nq = dim == 3 ? n1 * n1 * n1 : n1 * n1;
  q.pts.resize((size_t)q.nq * dim);
  q.w.resize(q.nq);
  for (int i = 0; i < q.nq; ++i) {
    int ix = i % n1, iy = (i / n1) % n1, iz = i / (n1 * n1);
    q.pts[i * dim + 0] = g.x[ix];
    q.pts[i * dim + 1] = g.x[iy];
    double ww = g.w[ix] * g.w[iy];
    if (dim == 3) {
      q.pts[i * dim + 2] = g.x[iz];
      ww *= g.w[iz];
    }
    q.w[i] = ww;
  }
  return q;
}

// Scalar Lagrange element of degree 1 or 2 tabulated on a quadrature rule.
struct ScalarTable {
  int dim = 0, degree = 0, nd = 0, nq = 0;
  std::vector<int> lex;        // [nd] lexicographic 3^dim lattice index of local dof a
  std::vector<double> phi;     // [nq][nd]
  std::vector<double> dphi;    // [nq][nd][dim]  reference gradients
};

inline std::vector<std::array<int, 3>> scalar_offsets(int dim, int degree) {
  auto all = hierarchical_offsets(dim);
  if (degree == 2) return all;
  std::vector<std::array<int, 3>> v(all.begin(), all.begin() + (dim == 3 ? 8 : 4));
  return v;
}

inline double shape_value(int dim, int degree, const std::array<int, 3>& o, const double* x) {
  double v = 1;
  for (int d = 0; d < dim; ++d) v *= lag1d(degree, o[d], x[d]);
  return v;
}
inline void shape_grad(int dim, int degree, const std::array<int, 3>& o, const double* x, double* g) {
  for (int e = 0; e < dim; ++e) {
    double v = 1;
    for (int d = 0; d < dim; ++d) v *= (d == e ? dlag1d(degree, o[d], x[d]) : lag1d(degree, o[d], x[d]));
    g[e] = v;
  }
}

inline ScalarTable tabulate_scalar(int dim, int degree, const QuadRule& q) {
  ScalarTable t;
  t.dim = dim;
  t.degree = degree;
  auto offs = scalar_offsets(dim, degree);
  t.nd = (int)offs.size();
  t.nq = q.nq;
  t.lex.resize(t.nd);
  t.phi.resize((size_t)t.nq * t.nd);
  t.dphi.resize((size_t)t.nq * t.nd * dim);
  for (int a = 0; a < t.nd; ++a) t.lex[a] = lex_index(dim, offs[a]);
  for (int iq = 0; iq < q.nq; ++iq)
    for (int a = 0; a < t.nd; ++a) {
      t.phi[(size_t)iq * t.nd + a] = shape_value(dim, degree, offs[a], &q.pts[iq * dim]);
      shape_grad(dim, degree, offs[a], &q.pts[iq * dim], &t.dphi[((size_t)iq * t.nd + a) * dim]);
    }
  return t;
}

// ---- mapping bases --------------------------------------------------------------------------
// Lagrange basis of degree m on Gauss-Lobatto nodes of [0,1] (m=1: {0,1}; m=3: {0,(1-1/sqrt5)/2,
// (1+1/sqrt5)/2,1}), tensor product, lexicographic support-point order.
inline std::vector<double> gauss_lobatto01(int m) {
  if (m == 1) return {0.0, 1.0};
  if (m == 2) return {0.0, 0.5, 1.0};
  if (m == 3) {
    double s = 1.0 / std::sqrt(5.0);
    return {0.0, 0.5 * (1.0 - s), 0.5 * (1.0 + s), 1.0};
  }
  throw std::runtime_error("mapping degree must be 1, 2 or 3");
}
inline double lagrange_nodes(const std::vector<double>& t, int i, double x) {
  double v = 1;
  for (size_t j = 0; j < t.size(); ++j)
    if ((int)j != i) v *= (x - t[j]) / (t[i] - t[j]);
  return v;
}
inline double dlagrange_nodes(const std::vector<double>& t, int i, double x) {
  double s = 0;
  for (size_t k = 0; k < t.size(); ++k) {
    if ((int)k == i) continue;
    double v = 1.0 / (t[i] - t[k]);
    for (size_t j = 0; j < t.size(); ++j)
      if ((int)j != i && j != k) v *= (x - t[j]) / (t[i] - t[j]);
    s += v;
  }
  return s;
}

struct MappingTable {
  int dim = 0, m = 0, ns = 0, nq = 0;
  std::vector<double> node1d;
  std::vector<double> N;   // [nq][ns]
  std::vector<double> dN;  // [nq][ns][dim]
};

inline void mapping_basis_at(int dim, const std::vector<double>& t, const double* x, double* N, double* dN) {
  int n1 = (int)t.size();
  int ns = dim == 3 ? n1 * n1 * n1 : n1 * n1;
  double l[3][8], dl[3][8];
  for (int d = 0; d < dim; ++d)
    for (int i = 0; i < n1; ++i) {
      l[d][i] = lagrange_nodes(t, i, x[d]);
      dl[d][i] = dlagrange_nodes(t, i, x[d]);
    }
  for (int s = 0; s < ns; ++s) {
    int i[3] = {s % n1, (s / n1) % n1, s / (n1 * n1)};
    double v = 1;
    for (int d = 0; d < dim; ++d) v *= l[d][i[d]];
    if (N) N[s] = v;
    if (dN)
      for (int e = 0; e < dim; ++e) {
        double g = 1;
        for (int d = 0; d < dim; ++d) g *= (d == e ? dl[d][i[d]] : l[d][i[d]]);
        dN[s * dim + e] = g;
      }
  }
}

inline MappingTable tabulate_mapping(int dim, int m, const QuadRule& q) {
  MappingTable t;
  t.dim = dim;
  t.m = m;
  t.node1d = gauss_lobatto01(m);
  int n1 = m + 1;
  t.ns = dim == 3 ? n1 * n1 * n1 : n1 * n1;
  t.nq = q.nq;
  t.N.resize((size_t)t.nq * t.ns);
  t.dN.resize((size_t)t.nq * t.ns * dim);
  for (int iq = 0; iq < q.nq; ++iq)
    mapping_basis_at(dim, t.node1d, &q.pts[iq * dim], &t.N[(size_t)iq * t.ns], &t.dN[(size_t)iq * t.ns * dim]);
  return t;
}

}  // namespace dcph
