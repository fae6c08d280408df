// Stand-in meshes for the reference's PlanetGeometry (include/core/planet_geometry.tpp:29-95,111-120):
//   * ShellMesh3D  = GridGenerator::hyper_shell(tria, 0, R0, R1, 6, colorize=true) + refine_global(r)
//                    with a SphericalManifold on everything (planet_geometry.tpp:63-68,116)
//   * CubeMesh3D   = GridGenerator::hyper_rectangle([0,1]^3, colorize=true) + refine_global(r)
//                    (planet_geometry.tpp:31-40), x/y periodic
//
// deal.II and p4est are un-vendored dependencies of the reference; what is restated here from memory of
// deal.II 9.2 (and cannot be verified in this environment) is: (a) the topology "6 frustum cells, inner
// boundary id 0, outer id 1"; (b) refinement on a SphericalManifold: a new line vertex is the geodesic
// midpoint at the mean radius, a new quad/hex vertex is SphericalManifold::get_new_point of the
// surrounding points with the transfinite weights (-1/4 vertices, +1/2 line midpoints, ...), which is the
// weighted Riemannian centre of mass of the directions (Newton, <= 10 iterations, tol 1e-10) at the
// weighted mean radius; (c) active cells ordered (coarse cell, Morton child index, x fastest), which is
// also the p4est space-filling curve used for the rank partition.
// With those rules the shell is a product mesh: position(surface node s, radial level k) = r_k * d_s
// with uniform r_k (the transfinite weights cancel all off-level directions).
// The orientation of the six coarse cells is this file's own choice (local x,y on the cube face, local z
// radial, right-handed); deal.II's hard-coded vertex lists are not reproducible here.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace dcph {

// ---- small 3-vector helpers -----------------------------------------------------------------
struct V3 {
  double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 unit(V3 a) { return (1.0 / norm(a)) * a; }

// Weighted centre of mass of unit directions on S^2 (restatement of SphericalManifold<3>::get_new_point /
// do_get_new_point of deal.II 9.2, manifold_lib.cc; recalled, not verifiable here).
inline V3 sphere_mean(const V3* d, const double* w, int n) {
  const double tol = 1e-10;
  V3 c{0, 0, 0};
  for (int i = 0; i < n; ++i) c = c + w[i] * d[i];
  c = unit(c);
  // merge duplicates / early exit when all directions coincide with the candidate
  bool all_close = true;
  for (int i = 0; i < n; ++i)
    if (norm(d[i] - c) > tol) all_close = false;
  if (all_close) return c;
  int n_distinct = 0;
  {
    std::vector<V3> seen;
    for (int i = 0; i < n; ++i) {
      bool dup = false;
      for (auto& s : seen)
        if (norm(s - d[i]) < tol) dup = true;
      if (!dup) seen.push_back(d[i]);
    }
    n_distinct = (int)seen.size();
  }
  if (n_distinct <= 2) return c;
  for (int it = 0; it < 10; ++it) {
    // local orthonormal tangent basis at c
    V3 ref = std::fabs(c.x) <= std::fabs(c.y) && std::fabs(c.x) <= std::fabs(c.z) ? V3{1, 0, 0}
             : (std::fabs(c.y) <= std::fabs(c.z) ? V3{0, 1, 0} : V3{0, 0, 1});
    V3 ex = unit(cross(c, ref));
    V3 ey = cross(c, ex);
    double g0 = 0, g1 = 0, h00 = 0, h01 = 0, h11 = 0;
    for (int i = 0; i < n; ++i) {
      V3 vp = d[i] - dot(d[i], c) * c;
      double s2 = dot(vp, vp), s = std::sqrt(s2);
      if (s < tol) {
        h00 += w[i];
        h11 += w[i];
        continue;
      }
      double ct = dot(d[i], c);
      double th = std::atan2(s, ct);
      double sinc_inv = th / s;
      double cp = dot(vp, ex), sp = dot(vp, ey);
      g0 += w[i] * sinc_inv * cp;
      g1 += w[i] * sinc_inv * sp;
      double wt = w[i] / s2, tt = sinc_inv * ct;
      h00 += wt * (cp * cp + tt * sp * sp);
      h01 += cp * sp * wt * (1.0 - tt);
      h11 += wt * (sp * sp + tt * cp * cp);
    }
    double det = h00 * h11 - h01 * h01;
    double dx = (h11 * g0 - h01 * g1) / det, dy = (-h01 * g0 + h00 * g1) / det;
    V3 disp = dx * ex + dy * ey;
    double th = norm(disp);
    V3 cn = c;
    if (th >= 1e-10) cn = unit(std::cos(th) * c + (std::sin(th) / th) * disp);
    double moved = norm(cn - c);
    c = cn;
    if (moved < tol) break;
  }
  return c;
}

// ---- abstract structured mesh -----------------------------------------------------------------
struct Mesh {
  int dim = 3;
  int64_t n_cells = 0;
  int64_t n_nodes = 0;  // half-step lattice nodes (vertices, line/quad/hex midpoints)
  virtual ~Mesh() {}
  // ids of the 3^dim lattice nodes of cell c, lexicographic (x fastest)
  virtual void cell_nodes(int64_t c, int64_t* ids) const = 0;
  // vertex coordinates, lexicographic, [2^dim][dim]
  virtual void cell_vertices(int64_t c, double* X) const = 0;
  // boundary id of face f (deal.II face order x0,x1,y0,y1,z0,z1) or -1
  virtual int face_boundary_id(int64_t c, int f) const = 0;
  // does a MappingQ(p>1) use the high-order mapping on this cell (cell has boundary lines)?
  virtual bool at_boundary(int64_t c) const = 0;
  // point of the cell's manifold at reference coordinates xi (used for mapping support points)
  virtual void manifold_point(int64_t c, const double* xi, double* x) const = 0;
  // periodic partner: lattice node id that `node` is identified with (slave -> master), or -1
  virtual int64_t periodic_master(int64_t /*node*/) const { return -1; }
};

inline uint32_t morton3(uint32_t i, uint32_t j, uint32_t k) {
  auto spread = [](uint64_t v) {
    v &= 0x1fffff;
    v = (v | v << 32) & 0x1f00000000ffffULL;
    v = (v | v << 16) & 0x1f0000ff0000ffULL;
    v = (v | v << 8) & 0x100f00f00f00f00fULL;
    v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
    v = (v | v << 2) & 0x1249249249249249ULL;
    return v;
  };
  return (uint32_t)(spread(i) | (spread(j) << 1) | (spread(k) << 2));
}
inline void demorton3(uint32_t m, uint32_t& i, uint32_t& j, uint32_t& k) {
  auto compact = [](uint64_t v) {
    v &= 0x1249249249249249ULL;
    v = (v ^ (v >> 2)) & 0x10c30c30c30c30c3ULL;
    v = (v ^ (v >> 4)) & 0x100f00f00f00f00fULL;
    v = (v ^ (v >> 8)) & 0x1f0000ff0000ffULL;
    v = (v ^ (v >> 16)) & 0x1f00000000ffffULL;
    v = (v ^ (v >> 32)) & 0x1fffff;
    return (uint32_t)v;
  };
  i = compact(m);
  j = compact(m >> 1);
  k = compact(m >> 2);
}

// ---- cubed-sphere hypershell -------------------------------------------------------------------
struct ShellMesh3D : Mesh {
  int r, n, M;
  int nr, Mr;  // radial layers (n * radial_factor) and radial half-step lattice size
  double R0, R1;
  int64_t n_surf = 0;
  std::vector<int32_t> sid;   // (M+1)^3 -> surface node id or -1
  std::vector<V3> sdir;       // [n_surf] unit direction (valid on the vertex lattice = even coords)
  // tree frames on the integer cube [0,M]^3: P = O + a*A + b*B, A x B = outward normal
  int O[6][3], A[6][3], B[6][3];

  // radial_factor > 1 gives the "synthetic refinement" used for weak scaling: n_r = n * radial_factor layers
  ShellMesh3D(int refinements, double r0, double r1, int radial_factor = 1) : r(refinements), R0(r0), R1(r1) {
    dim = 3;
    n = 1 << r;
    M = 2 * n;
    nr = n * radial_factor;
    Mr = 2 * nr;
    n_cells = 6LL * n * n * nr;
    const int o[6][3] = {{1, 0, 0}, {0, 0, 0}, {0, 1, 0}, {0, 0, 0}, {0, 0, 1}, {0, 0, 0}};
    const int a[6][3] = {{0, 1, 0}, {0, 0, 1}, {0, 0, 1}, {1, 0, 0}, {1, 0, 0}, {0, 1, 0}};
    const int b[6][3] = {{0, 0, 1}, {0, 1, 0}, {1, 0, 0}, {0, 0, 1}, {0, 1, 0}, {1, 0, 0}};
    for (int t = 0; t < 6; ++t)
      for (int d = 0; d < 3; ++d) {
        O[t][d] = o[t][d] * M;
        A[t][d] = a[t][d];
        B[t][d] = b[t][d];
      }
    const int64_t L = M + 1;
    sid.assign((size_t)(L * L * L), -1);
    int32_t next = 0;
    for (int64_t z = 0; z < L; ++z)
      for (int64_t y = 0; y < L; ++y)
        for (int64_t x = 0; x < L; ++x)
          if (x == 0 || x == M || y == 0 || y == M || z == 0 || z == M) sid[(size_t)((z * L + y) * L + x)] = next++;
    n_surf = next;
    n_nodes = n_surf * (int64_t)(Mr + 1);
    build_directions();
  }

  inline int32_t surf(int t, int a, int b) const {
    int64_t L = M + 1;
    int64_t x = O[t][0] + a * A[t][0] + b * B[t][0];
    int64_t y = O[t][1] + a * A[t][1] + b * B[t][1];
    int64_t z = O[t][2] + a * A[t][2] + b * B[t][2];
    return sid[(size_t)((z * L + y) * L + x)];
  }

  void build_directions() {
    sdir.assign((size_t)n_surf, V3{0, 0, 0});
    const double c = 1.0 / std::sqrt(3.0);
    const int64_t L = M + 1;
    for (int cz = 0; cz < 2; ++cz)
      for (int cy = 0; cy < 2; ++cy)
        for (int cx = 0; cx < 2; ++cx) {
          int32_t s = sid[(size_t)(((int64_t)cz * M * L + (int64_t)cy * M) * L + (int64_t)cx * M)];
          sdir[s] = V3{cx ? c : -c, cy ? c : -c, cz ? c : -c};
        }
    for (int lev = 0; lev < r; ++lev) {
      int s = M >> lev, h = s / 2, nc = 1 << lev;
      for (int t = 0; t < 6; ++t)
        for (int J = 0; J < nc; ++J)
          for (int I = 0; I < nc; ++I) {
            int a0 = I * s, b0 = J * s;
            V3 c00 = sdir[surf(t, a0, b0)], c10 = sdir[surf(t, a0 + s, b0)];
            V3 c01 = sdir[surf(t, a0, b0 + s)], c11 = sdir[surf(t, a0 + s, b0 + s)];
            V3 eb0 = unit(c00 + c10), eb1 = unit(c01 + c11), ea0 = unit(c00 + c01), ea1 = unit(c10 + c11);
            sdir[surf(t, a0 + h, b0)] = eb0;
            sdir[surf(t, a0 + h, b0 + s)] = eb1;
            sdir[surf(t, a0, b0 + h)] = ea0;
            sdir[surf(t, a0 + s, b0 + h)] = ea1;
            V3 pts[8] = {c00, c10, c01, c11, eb0, eb1, ea0, ea1};
            double w[8] = {-0.25, -0.25, -0.25, -0.25, 0.5, 0.5, 0.5, 0.5};
            sdir[surf(t, a0 + h, b0 + h)] = sphere_mean(pts, w, 8);
          }
    }
  }

  // cell order: tree, then radial super-block (for radial_factor > 1), then Morton within the n^3 cube
  inline void decode(int64_t c, int& t, int& i, int& j, int& k) const {
    int64_t cube = (int64_t)n * n * n, per = cube * (nr / n);
    t = (int)(c / per);
    int64_t rem = c % per;
    int khi = (int)(rem / cube);
    uint32_t ii, jj, kk;
    demorton3((uint32_t)(rem % cube), ii, jj, kk);
    i = (int)ii;
    j = (int)jj;
    k = khi * n + (int)kk;
  }
  inline int64_t encode(int t, int i, int j, int k) const {
    int64_t cube = (int64_t)n * n * n, per = cube * (nr / n);
    return (int64_t)t * per + (int64_t)(k / n) * cube + morton3((uint32_t)i, (uint32_t)j, (uint32_t)(k % n));
  }
  inline double radius(int kk) const { return R0 + (R1 - R0) * ((double)kk / (double)Mr); }

  void cell_nodes(int64_t c, int64_t* ids) const override {
    int t, i, j, k;
    decode(c, t, i, j, k);
    for (int oz = 0; oz < 3; ++oz)
      for (int oy = 0; oy < 3; ++oy)
        for (int ox = 0; ox < 3; ++ox)
          ids[ox + 3 * (oy + 3 * oz)] = (int64_t)surf(t, 2 * i + ox, 2 * j + oy) * (Mr + 1) + (2 * k + oz);
  }
  void cell_vertices(int64_t c, double* X) const override {
    int t, i, j, k;
    decode(c, t, i, j, k);
    for (int v = 0; v < 8; ++v) {
      int ox = 2 * (v & 1), oy = 2 * ((v >> 1) & 1), oz = 2 * ((v >> 2) & 1);
      V3 d = sdir[surf(t, 2 * i + ox, 2 * j + oy)];
      double rr = radius(2 * k + oz);
      X[3 * v + 0] = rr * d.x;
      X[3 * v + 1] = rr * d.y;
      X[3 * v + 2] = rr * d.z;
    }
  }
  int face_boundary_id(int64_t c, int f) const override {
    int t, i, j, k;
    decode(c, t, i, j, k);
    if (f == 4 && k == 0) return 0;
    if (f == 5 && k == nr - 1) return 1;
    return -1;
  }
  bool at_boundary(int64_t c) const override {
    int t, i, j, k;
    decode(c, t, i, j, k);
    return k == 0 || k == nr - 1;
  }
  void manifold_point(int64_t c, const double* xi, double* x) const override {
    int t, i, j, k;
    decode(c, t, i, j, k);
    V3 d[4];
    double w[4];
    for (int v = 0; v < 4; ++v) {
      int ox = 2 * (v & 1), oy = 2 * ((v >> 1) & 1);
      d[v] = sdir[surf(t, 2 * i + ox, 2 * j + oy)];
      w[v] = ((v & 1) ? xi[0] : 1.0 - xi[0]) * (((v >> 1) & 1) ? xi[1] : 1.0 - xi[1]);
    }
    V3 dm = sphere_mean(d, w, 4);
    double rr = (1.0 - xi[2]) * radius(2 * k) + xi[2] * radius(2 * k + 2);
    x[0] = rr * dm.x;
    x[1] = rr * dm.y;
    x[2] = rr * dm.z;
  }
};

// ---- unit cube, periodic in x and y ----------------------------------------------------------
// hyper_rectangle colorize=true: boundary ids 0/1 = x min/max, 2/3 = y, 4/5 = z (planet_geometry.tpp:31-40)
struct CubeMesh3D : Mesh {
  int r, n, M;
  bool periodic_xy;
  CubeMesh3D(int refinements, bool periodic) : r(refinements), periodic_xy(periodic) {
    dim = 3;
    n = 1 << r;
    M = 2 * n;
    n_cells = (int64_t)n * n * n;
    int64_t L = M + 1;
    n_nodes = L * L * L;
  }
  inline void decode(int64_t c, int& i, int& j, int& k) const {
    uint32_t ii, jj, kk;
    demorton3((uint32_t)c, ii, jj, kk);
    i = (int)ii;
    j = (int)jj;
    k = (int)kk;
  }
  void cell_nodes(int64_t c, int64_t* ids) const override {
    int i, j, k;
    decode(c, i, j, k);
    int64_t L = M + 1;
    for (int oz = 0; oz < 3; ++oz)
      for (int oy = 0; oy < 3; ++oy)
        for (int ox = 0; ox < 3; ++ox) ids[ox + 3 * (oy + 3 * oz)] = ((int64_t)(2 * k + oz) * L + (2 * j + oy)) * L + (2 * i + ox);
  }
  void cell_vertices(int64_t c, double* X) const override {
    int i, j, k;
    decode(c, i, j, k);
    for (int v = 0; v < 8; ++v) {
      X[3 * v + 0] = (double)(i + (v & 1)) / n;
      X[3 * v + 1] = (double)(j + ((v >> 1) & 1)) / n;
      X[3 * v + 2] = (double)(k + ((v >> 2) & 1)) / n;
    }
  }
  int face_boundary_id(int64_t c, int f) const override {
    int i, j, k;
    decode(c, i, j, k);
    int idx[3] = {i, j, k};
    int d = f / 2, side = f % 2;
    if (side == 0 && idx[d] == 0) return 2 * d;
    if (side == 1 && idx[d] == n - 1) return 2 * d + 1;
    return -1;
  }
  bool at_boundary(int64_t) const override { return false; }  // flat: high-order mapping == trilinear
  void manifold_point(int64_t c, const double* xi, double* x) const override {
    int i, j, k;
    decode(c, i, j, k);
    x[0] = (i + xi[0]) / n;
    x[1] = (j + xi[1]) / n;
    x[2] = (k + xi[2]) / n;
  }
  int64_t periodic_master(int64_t node) const override {
    if (!periodic_xy) return -1;
    int64_t L = M + 1;
    int64_t x = node % L, y = (node / L) % L, z = node / (L * L);
    if (x != M && y != M) return -1;
    if (x == M) x = 0;
    if (y == M) y = 0;
    return (z * L + y) * L + x;
  }
};

// ---- 2-D annulus: GridGenerator::hyper_shell<2>(tria, 0, R0, R1, 12, colorize=true) + refine_global(r) -------
// (planet_geometry.tpp:63-68 with dim == 2).  On the polar manifold every refinement halves angle and radius
// increments, so the refined mesh is the uniform polar grid: position(a, k) = r_k (cos th_a, sin th_a).
// Orientation chosen here (deal.II's vertex lists are not reproducible): local x = radial (outward), local
// y = counter-clockwise, which is right-handed.  Boundary ids: inner 0 (face 0), outer 1 (face 1).
inline uint32_t morton2(uint32_t i, uint32_t j) {
  auto spread = [](uint32_t v) {
    v &= 0xffff;
    v = (v | (v << 8)) & 0x00ff00ff;
    v = (v | (v << 4)) & 0x0f0f0f0f;
    v = (v | (v << 2)) & 0x33333333;
    v = (v | (v << 1)) & 0x55555555;
    return v;
  };
  return spread(i) | (spread(j) << 1);
}
inline void demorton2(uint32_t m, uint32_t& i, uint32_t& j) {
  auto compact = [](uint32_t v) {
    v &= 0x55555555;
    v = (v | (v >> 1)) & 0x33333333;
    v = (v | (v >> 2)) & 0x0f0f0f0f;
    v = (v | (v >> 4)) & 0x00ff00ff;
    v = (v | (v >> 8)) & 0x0000ffff;
    return v;
  };
  i = compact(m);
  j = compact(m >> 1);
}

struct AnnulusMesh2D : Mesh {
  int r, n, Mr, Ma;  // n = 2^r cells per coarse cell edge; radial / angular half-step lattice sizes
  double R0, R1;
  AnnulusMesh2D(int refinements, double r0, double r1) : r(refinements), R0(r0), R1(r1) {
    dim = 2;
    n = 1 << r;
    Mr = 2 * n;
    Ma = 2 * 12 * n;
    n_cells = 12LL * n * n;
    n_nodes = (int64_t)(Mr + 1) * Ma;
  }
  // cell = (tree t in 0..11, i radial, j angular within the tree), Morton inside the tree
  inline void decode(int64_t c, int& t, int& i, int& j) const {
    int64_t per = (int64_t)n * n;
    t = (int)(c / per);
    uint32_t ii, jj;
    demorton2((uint32_t)(c % per), ii, jj);
    i = (int)ii;
    j = (int)jj;
  }
  inline double radius(int k) const { return R0 + (R1 - R0) * ((double)k / (double)Mr); }
  inline double angle(int a) const { return 2.0 * M_PI * ((double)a / (double)Ma); }
  void cell_nodes(int64_t c, int64_t* ids) const override {
    int t, i, j;
    decode(c, t, i, j);
    for (int oy = 0; oy < 3; ++oy)
      for (int ox = 0; ox < 3; ++ox) {
        int k = 2 * i + ox, a = (2 * (t * n + j) + oy) % Ma;
        ids[ox + 3 * oy] = (int64_t)k * Ma + a;
      }
  }
  void cell_vertices(int64_t c, double* X) const override {
    int t, i, j;
    decode(c, t, i, j);
    for (int v = 0; v < 4; ++v) {
      int k = 2 * (i + (v & 1)), a = 2 * (t * n + j + ((v >> 1) & 1));
      X[2 * v + 0] = radius(k) * std::cos(angle(a));
      X[2 * v + 1] = radius(k) * std::sin(angle(a));
    }
  }
  int face_boundary_id(int64_t c, int f) const override {
    int t, i, j;
    decode(c, t, i, j);
    if (f == 0 && i == 0) return 0;
    if (f == 1 && i == n - 1) return 1;
    return -1;
  }
  bool at_boundary(int64_t c) const override {
    int t, i, j;
    decode(c, t, i, j);
    return i == 0 || i == n - 1;
  }
  void manifold_point(int64_t c, const double* xi, double* x) const override {
    int t, i, j;
    decode(c, t, i, j);
    double rr = (1.0 - xi[0]) * radius(2 * i) + xi[0] * radius(2 * i + 2);
    double th = (1.0 - xi[1]) * angle(2 * (t * n + j)) + xi[1] * angle(2 * (t * n + j + 1));
    x[0] = rr * std::cos(th);
    x[1] = rr * std::sin(th);
  }
};

// ---- a subset of the cells of a base mesh (one rank's owned + ghost cells) -------------------------------
struct SubMesh : Mesh {
  const Mesh& base;
  std::vector<int64_t> cells;  // local -> base cell id
  SubMesh(const Mesh& b, std::vector<int64_t> c) : base(b), cells(std::move(c)) {
    dim = b.dim;
    n_cells = (int64_t)cells.size();
    n_nodes = b.n_nodes;
  }
  void cell_nodes(int64_t c, int64_t* ids) const override { base.cell_nodes(cells[c], ids); }
  void cell_vertices(int64_t c, double* X) const override { base.cell_vertices(cells[c], X); }
  int face_boundary_id(int64_t c, int f) const override { return base.face_boundary_id(cells[c], f); }
  bool at_boundary(int64_t c) const override { return base.at_boundary(cells[c]); }
  void manifold_point(int64_t c, const double* xi, double* x) const override { base.manifold_point(cells[c], xi, x); }
  int64_t periodic_master(int64_t node) const override { return base.periodic_master(node); }
};

}  // namespace dcph
