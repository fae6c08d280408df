// Stand-in for the deal.II DoFHandler / AffineConstraints / DoFTools calls made by
// Standard::BoussinesqModel::setup_dofs (include/core/boussinesq_model.tpp:184-412) and the three
// setup_*_matrices (:79-180).  Restated from memory of deal.II 9.2 (un-vendored; unverifiable here):
//   * distribute_dofs: cells in active order, per cell vertices -> lines -> quads -> hex, all DoFs of
//     an object numbered together when the object is first met; FESystem interleaves its base
//     elements per object.
//   * DoFRenumbering::component_wise(dh, blocks): stable partition by target block (:204).
//   * interpolate_boundary_values -> Dirichlet lines (:313-318, 377-383).
//   * compute_no_normal_flux_constraints (:321-329): normal from the mapping averaged over adjacent
//     faces, component of largest |n_k| eliminated, u_k = -sum_{i!=k} n_i/n_k u_i.
//   * make_sparsity_pattern(dh, coupling, sp, constraints, keep_constrained_dofs=false) (:98-104).
#pragma once
#include <omp.h>

#include <algorithm>
#include <cstring>
#include <functional>
#include <numeric>

#include "fe_tables.hpp"
#include "mesh.hpp"

namespace dcph {

// A system of continuous Lagrange fields (degree 1 or 2 each) on one mesh.
struct FESystemDesc {
  int dim = 3;
  std::vector<int> field_degree;  // per field (= component)
  std::vector<int> field_block;   // target block per field (component_wise)
  // 0 = continuous Lagrange of `field_degree`; lowest-order FEEC spaces (one dof per entity):
  // 1 = Nedelec (lines), 2 = Raviart-Thomas (faces), 3 = DGQ0 (cell)
  std::vector<int> field_kind;
  bool present(size_t f, int ed) const {
    const int kind = field_kind.empty() ? 0 : field_kind[f];
    if (kind == 0) return field_degree[f] == 2 || ed == 0;
    return ed == (kind == 1 ? 1 : (kind == 2 ? dim - 1 : dim));
  }
  // derived
  int n_local = 0;
  std::vector<int> local_field;  // [n_local] component
  std::vector<int> local_lex;    // [n_local] lexicographic 3^dim lattice index
  std::vector<int> local_base;   // [n_local] index within the scalar base element (hierarchical)

  void finalize() {
    auto offs = hierarchical_offsets(dim);
    local_field.clear();
    local_lex.clear();
    local_base.clear();
    // per-field running base index: base elements are themselves hierarchical, so the base index of
    // a field at hierarchical node h is h for Q2, and the vertex number for Q1.
    for (size_t h = 0; h < offs.size(); ++h) {
      int ed = entity_dim(dim, offs[h]);
      for (size_t f = 0; f < field_degree.size(); ++f) {
        if (!present(f, ed)) continue;
        local_field.push_back((int)f);
        local_lex.push_back(lex_index(dim, offs[h]));
        local_base.push_back((int)h);
      }
    }
    n_local = (int)local_field.size();
  }
};

struct Constraints {
  int64_t n_dofs = 0;
  std::vector<int32_t> line_of_dof;  // [n_dofs] -> line or -1
  std::vector<int32_t> line_dof;     // [n_lines]
  std::vector<int32_t> line_ptr;     // [n_lines+1]
  std::vector<int32_t> entry_dof;
  std::vector<double> entry_w;
  std::vector<double> inhom;  // [n_lines]

  // lines must be added in any order; close() sorts them by dof.
  struct Tmp {
    int32_t dof;
    std::vector<std::pair<int32_t, double>> e;
    double inhom;
  };
  std::vector<Tmp> tmp;
  void add_line(int32_t dof, std::vector<std::pair<int32_t, double>> e, double ih) {
    tmp.push_back({dof, std::move(e), ih});
  }
  void close(int64_t n) {
    n_dofs = n;
    std::stable_sort(tmp.begin(), tmp.end(), [](const Tmp& a, const Tmp& b) { return a.dof < b.dof; });
    // first line for a dof wins (AffineConstraints::add_line ignores re-adding an existing line)
    std::vector<Tmp> u;
    for (auto& t : tmp)
      if (u.empty() || u.back().dof != t.dof) u.push_back(t);
    tmp.swap(u);
    line_of_dof.assign((size_t)n, -1);
    for (size_t l = 0; l < tmp.size(); ++l) line_of_dof[tmp[l].dof] = (int32_t)l;
    // resolve chains: masters that are themselves constrained (bounded depth)
    for (int pass = 0; pass < 8; ++pass) {
      bool changed = false;
      for (auto& t : tmp) {
        std::vector<std::pair<int32_t, double>> ne;
        for (auto& e : t.e) {
          int32_t l2 = line_of_dof[e.first];
          if (l2 < 0) {
            ne.push_back(e);
            continue;
          }
          changed = true;
          t.inhom += e.second * tmp[l2].inhom;
          for (auto& e2 : tmp[l2].e) ne.push_back({e2.first, e.second * e2.second});
        }
        // merge duplicates
        std::sort(ne.begin(), ne.end());
        std::vector<std::pair<int32_t, double>> m;
        for (auto& e : ne) {
          if (!m.empty() && m.back().first == e.first)
            m.back().second += e.second;
          else
            m.push_back(e);
        }
        t.e.swap(m);
      }
      if (!changed) break;
    }
    line_dof.clear();
    line_ptr.assign(1, 0);
    entry_dof.clear();
    entry_w.clear();
    inhom.clear();
    for (auto& t : tmp) {
      line_dof.push_back(t.dof);
      for (auto& e : t.e) {
        entry_dof.push_back(e.first);
        entry_w.push_back(e.second);
      }
      line_ptr.push_back((int32_t)entry_dof.size());
      inhom.push_back(t.inhom);
    }
    tmp.clear();
    tmp.shrink_to_fit();
  }
  int64_t n_lines() const { return (int64_t)line_dof.size(); }
};

struct DofMap {
  FESystemDesc fe;
  int64_t n_cells = 0, n_dofs = 0;
  std::vector<int64_t> block_size;    // dofs per block (owned + ghost on this rank)
  std::vector<int64_t> owned_size;    // dofs per block owned by this rank (numbered first within the block)
  std::vector<int64_t> dof_key;       // [n_dofs] global key (lattice node * 8 + rank of the field on the node)
  std::vector<int32_t> dof_owner;     // [n_dofs] owning rank
  std::vector<int32_t> l2g;           // [n_cells][n_local], block-concatenated global numbering
  std::vector<int32_t> node_first;    // [n_nodes] first (pre-renumbering) dof at lattice node or -1
  std::vector<int32_t> renumber;      // [n_dofs] pre -> final
  std::vector<int8_t> node_nfields;   // helper: number of fields present per node type, by entity dim
  // rank of field f among the fields present on an entity of dimension ed, or -1
  std::vector<std::array<int, 4>> field_rank;

  int32_t dof_at(int64_t node, int ed, int field) const {
    int rk = field_rank[field][ed];
    if (rk < 0 || node_first[node] < 0) return -1;
    return renumber[node_first[node] + rk];
  }
};

// node_owner (optional): owning rank per lattice node; dofs owned by other ranks are numbered after the owned
// ones inside each block (Trilinos-style local numbering: owned rows first, then ghost columns).
// Order in which the pre-dofs (numbered by cell / object order, like DoFHandler::distribute_dofs) are visited by the
// component-wise stable partition: identity, Cuthill-McKee (what the reference applies before component_wise when the
// Schur-complement solver is selected, boussinesq_model.tpp:198-202; restated from memory: breadth-first from a dof of
// least coordination, neighbours in order of increasing coordination), or a seeded random permutation (a numbering
// with no locality at all -- the device path must not care).
inline std::vector<int32_t> pre_dof_order(const std::string& mode, int64_t n, int64_t n_cells, int n_pre_per_cell,
                                          const std::vector<int32_t>& cell_pre /* [n_cells][n_pre_per_cell] */) {
  std::vector<int32_t> order((size_t)n);
  for (int64_t i = 0; i < n; ++i) order[(size_t)i] = (int32_t)i;
  if (mode == "none" || mode.empty()) return order;
  if (mode == "random") {
    uint64_t st = 0x9E3779B97F4A7C15ull;
    for (int64_t i = n - 1; i > 0; --i) {
      st ^= st << 13;
      st ^= st >> 7;
      st ^= st << 17;
      std::swap(order[(size_t)i], order[(size_t)(st % (uint64_t)(i + 1))]);
    }
    return order;
  }
  if (mode != "cuthill_mckee") throw std::runtime_error("harness: unknown renumber mode " + mode);
  // dof graph: two dofs are adjacent when a cell holds both
  std::vector<std::vector<int32_t>> adj((size_t)n);
  for (int64_t c = 0; c < n_cells; ++c) {
    const int32_t* d = &cell_pre[(size_t)c * n_pre_per_cell];
    for (int i = 0; i < n_pre_per_cell; ++i)
      for (int j = 0; j < n_pre_per_cell; ++j)
        if (i != j) adj[(size_t)d[i]].push_back(d[j]);
  }
  for (auto& a : adj) {
    std::sort(a.begin(), a.end());
    a.erase(std::unique(a.begin(), a.end()), a.end());
  }
  std::vector<uint8_t> seen((size_t)n, 0);
  std::vector<int32_t> by_degree(order);
  std::stable_sort(by_degree.begin(), by_degree.end(), [&](int32_t x, int32_t y) { return adj[(size_t)x].size() < adj[(size_t)y].size(); });
  size_t out = 0, head = 0, next_seed = 0;
  while (out < (size_t)n) {
    while (seen[(size_t)by_degree[next_seed]]) ++next_seed;
    const int32_t seed = by_degree[next_seed];
    seen[(size_t)seed] = 1;
    order[out++] = seed;
    while (head < out) {
      const int32_t v = order[head++];
      std::vector<int32_t> nb;
      for (int32_t w : adj[(size_t)v])
        if (!seen[(size_t)w]) nb.push_back(w);
      std::stable_sort(nb.begin(), nb.end(), [&](int32_t x, int32_t y) { return adj[(size_t)x].size() < adj[(size_t)y].size(); });
      for (int32_t w : nb) {
        seen[(size_t)w] = 1;
        order[out++] = w;
      }
    }
  }
  return order;
}

inline DofMap distribute_dofs(const Mesh& mesh, FESystemDesc fe, const std::vector<int32_t>* node_owner = nullptr,
                              int rank = 0, const std::string& renumber = "none") {
  DofMap dm;
  fe.dim = mesh.dim;
  fe.finalize();
  dm.fe = fe;
  dm.n_cells = mesh.n_cells;
  const int dim = mesh.dim;
  const int nf = (int)fe.field_degree.size();
  dm.field_rank.assign(nf, {-1, -1, -1, -1});
  int count_by_ed[4] = {0, 0, 0, 0};
  for (int ed = 0; ed <= dim; ++ed)
    for (int f = 0; f < nf; ++f)
      if (fe.present(f, ed)) dm.field_rank[f][ed] = count_by_ed[ed]++;
  auto offs = hierarchical_offsets(dim);
  const int n3 = dim == 3 ? 27 : 9;
  std::vector<int> lex_ed(n3);
  for (auto& o : offs) lex_ed[lex_index(dim, o)] = entity_dim(dim, o);

  dm.node_first.assign((size_t)mesh.n_nodes, -1);
  int64_t next = 0;
  std::vector<int64_t> ids(n3);
  // pass 1: number objects in cell order / hierarchical object order
  for (int64_t c = 0; c < mesh.n_cells; ++c) {
    mesh.cell_nodes(c, ids.data());
    for (auto& o : offs) {
      int lx = lex_index(dim, o);
      int ed = lex_ed[lx];
      if (count_by_ed[ed] == 0) continue;
      int64_t node = ids[lx];
      if (dm.node_first[node] < 0) {
        dm.node_first[node] = (int32_t)next;
        next += count_by_ed[ed];
      }
    }
  }
  dm.n_dofs = next;
  // component_wise: stable partition by block.  Need field of each pre-dof -> walk nodes in dof order.
  int n_blocks = 0;
  for (int b : fe.field_block) n_blocks = std::max(n_blocks, b + 1);
  dm.block_size.assign(n_blocks, 0);
  dm.owned_size.assign(n_blocks, 0);
  dm.dof_key.assign((size_t)next, -1);
  dm.dof_owner.assign((size_t)next, rank);
  std::vector<int8_t> pre_block((size_t)next, 0);   // pseudo block = 2*block + is_ghost
  std::vector<int64_t> pre_key((size_t)next, -1);
  std::vector<int32_t> pre_owner((size_t)next, rank);
  {
    std::vector<int8_t> node_ed((size_t)mesh.n_nodes, -1);
    for (int64_t c = 0; c < mesh.n_cells; ++c) {
      mesh.cell_nodes(c, ids.data());
      for (int lx = 0; lx < n3; ++lx) node_ed[ids[lx]] = (int8_t)lex_ed[lx];
    }
    for (int64_t node = 0; node < mesh.n_nodes; ++node) {
      int32_t f0 = dm.node_first[node];
      if (f0 < 0) continue;
      int ed = node_ed[node];
      const int own = node_owner ? (*node_owner)[node] : rank;
      for (int f = 0; f < nf; ++f) {
        int rk = dm.field_rank[f][ed];
        if (rk >= 0) {
          pre_block[f0 + rk] = (int8_t)(2 * fe.field_block[f] + (own != rank ? 1 : 0));
          pre_key[f0 + rk] = node * 8 + rk;
          pre_owner[f0 + rk] = own;
        }
      }
    }
  }
  std::vector<int64_t> pseudo_size(2 * n_blocks, 0);
  for (int64_t d = 0; d < next; ++d) pseudo_size[pre_block[d]]++;
  std::vector<int64_t> pseudo_start(2 * n_blocks + 1, 0);
  for (int b = 0; b < 2 * n_blocks; ++b) pseudo_start[b + 1] = pseudo_start[b] + pseudo_size[b];
  for (int b = 0; b < n_blocks; ++b) {
    dm.owned_size[b] = pseudo_size[2 * b];
    dm.block_size[b] = pseudo_size[2 * b] + pseudo_size[2 * b + 1];
  }
  dm.renumber.resize((size_t)next);
  {
    std::vector<int32_t> order;
    if (renumber != "none" && !renumber.empty()) {
      // pre-dofs of every cell (for the dof graph of Cuthill-McKee)
      std::vector<int32_t> cell_pre((size_t)mesh.n_cells * fe.n_local);
      std::vector<int64_t> lids(n3);
      for (int64_t c = 0; c < mesh.n_cells; ++c) {
        mesh.cell_nodes(c, lids.data());
        for (int i = 0; i < fe.n_local; ++i) {
          const int lx = fe.local_lex[i];
          cell_pre[(size_t)c * fe.n_local + i] = dm.node_first[lids[lx]] + dm.field_rank[fe.local_field[i]][lex_ed[lx]];
        }
      }
      order = pre_dof_order(renumber, next, mesh.n_cells, fe.n_local, cell_pre);
    }
    std::vector<int64_t> cur(pseudo_start.begin(), pseudo_start.end() - 1);
    for (int64_t k = 0; k < next; ++k) {
      const int64_t d = order.empty() ? k : order[(size_t)k];
      dm.renumber[d] = (int32_t)cur[pre_block[d]]++;
      dm.dof_key[dm.renumber[d]] = pre_key[d];
      dm.dof_owner[dm.renumber[d]] = pre_owner[d];
    }
  }
  // pass 2: cell -> global
  dm.l2g.resize((size_t)mesh.n_cells * fe.n_local);
#pragma omp parallel
  {
    std::vector<int64_t> lids(n3);
#pragma omp for schedule(static)
    for (int64_t c = 0; c < mesh.n_cells; ++c) {
      mesh.cell_nodes(c, lids.data());
      for (int i = 0; i < fe.n_local; ++i) {
        int lx = fe.local_lex[i];
        int ed = lex_ed[lx];
        dm.l2g[(size_t)c * fe.n_local + i] = dm.renumber[dm.node_first[lids[lx]] + dm.field_rank[fe.local_field[i]][ed]];
      }
    }
  }
  return dm;
}

// ---- CSR pattern ---------------------------------------------------------------------------------
struct Csr {
  int64_t n_rows = 0, n_cols = 0;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> col;
};

// row -> (cell, local dof, weight) adjacency after constraint resolution
struct RowAdjacency {
  std::vector<int64_t> ptr;   // [n_dofs+1]
  std::vector<int32_t> cell;  // entries
  std::vector<int16_t> loc;
  std::vector<double> w;
};

inline RowAdjacency build_row_adjacency(const DofMap& dm, const Constraints& cs) {
  RowAdjacency ra;
  const int nl = dm.fe.n_local;
  ra.ptr.assign((size_t)dm.n_dofs + 1, 0);
  auto visit = [&](auto&& fn) {
    for (int64_t c = 0; c < dm.n_cells; ++c)
      for (int i = 0; i < nl; ++i) {
        int32_t g = dm.l2g[(size_t)c * nl + i];
        int32_t ln = cs.line_of_dof[g];
        if (ln < 0)
          fn(g, c, i, 1.0);
        else
          for (int32_t e = cs.line_ptr[ln]; e < cs.line_ptr[ln + 1]; ++e) fn(cs.entry_dof[e], c, i, cs.entry_w[e]);
      }
  };
  visit([&](int32_t g, int64_t, int, double) { ra.ptr[g + 1]++; });
  for (int64_t d = 0; d < dm.n_dofs; ++d) ra.ptr[d + 1] += ra.ptr[d];
  ra.cell.resize((size_t)ra.ptr.back());
  ra.loc.resize((size_t)ra.ptr.back());
  ra.w.resize((size_t)ra.ptr.back());
  std::vector<int64_t> cur(ra.ptr.begin(), ra.ptr.end() - 1);
  visit([&](int32_t g, int64_t c, int i, double w) {
    int64_t p = cur[g]++;
    ra.cell[p] = (int32_t)c;
    ra.loc[p] = (int16_t)i;
    ra.w[p] = w;
  });
  return ra;
}

// coupling[fi*nf+fj] != 0  <=>  DoFTools::always
inline Csr make_sparsity_pattern(const DofMap& dm, const Constraints& cs, const std::vector<int>& coupling,
                                 const RowAdjacency& ra) {
  Csr A;
  const int nl = dm.fe.n_local;
  const int nf = (int)dm.fe.field_degree.size();
  A.n_rows = A.n_cols = dm.n_dofs;
  A.rowptr.assign((size_t)dm.n_dofs + 1, 0);
  std::vector<std::vector<int32_t>> rows;  // per-thread chunks would be heavy; do two passes
  auto row_cols = [&](int64_t g, std::vector<int32_t>& out) {
    out.clear();
    if (cs.line_of_dof[g] >= 0) {
      out.push_back((int32_t)g);  // constrained row keeps only its diagonal
      return;
    }
    for (int64_t p = ra.ptr[g]; p < ra.ptr[g + 1]; ++p) {
      int64_t c = ra.cell[p];
      int fi = dm.fe.local_field[ra.loc[p]];
      const int32_t* lg = &dm.l2g[(size_t)c * nl];
      for (int j = 0; j < nl; ++j) {
        if (!coupling[fi * nf + dm.fe.local_field[j]]) continue;
        int32_t gj = lg[j];
        int32_t ln = cs.line_of_dof[gj];
        if (ln < 0)
          out.push_back(gj);
        else
          for (int32_t e = cs.line_ptr[ln]; e < cs.line_ptr[ln + 1]; ++e) out.push_back(cs.entry_dof[e]);
      }
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
  };
  // one pass: every thread appends the rows it computes to its own arena, then the arenas are gathered
  const int nthreads = omp_get_max_threads();
  std::vector<std::vector<int32_t>> arena(nthreads);
  std::vector<int64_t> row_off((size_t)dm.n_dofs);
  std::vector<uint8_t> row_thr((size_t)dm.n_dofs);
#pragma omp parallel
  {
    const int t = omp_get_thread_num();
    std::vector<int32_t> tmp;
    std::vector<int32_t>& mine = arena[t];
    int64_t prev = -1;  // row whose columns `tmp` holds
#pragma omp for schedule(dynamic, 4096)
    for (int64_t g = 0; g < dm.n_dofs; ++g) {
      // consecutive unconstrained rows that live on the same cells and couple to the same fields (the velocity
      // components of one node in the system pattern) have the same columns: build them once
      bool same = prev == g - 1 && prev >= 0 && cs.line_of_dof[g] < 0 && cs.line_of_dof[prev] < 0 &&
                  ra.ptr[g + 1] - ra.ptr[g] == ra.ptr[prev + 1] - ra.ptr[prev];
      for (int64_t k = 0, n = ra.ptr[g + 1] - ra.ptr[g]; same && k < n; ++k) {
        const int64_t p = ra.ptr[g] + k, q = ra.ptr[prev] + k;
        same = ra.cell[p] == ra.cell[q];
        if (same) {
          const int fa = dm.fe.local_field[ra.loc[p]], fb = dm.fe.local_field[ra.loc[q]];
          for (int f = 0; f < nf && same; ++f) same = coupling[fa * nf + f] == coupling[fb * nf + f];
        }
      }
      if (!same) row_cols(g, tmp);
      prev = g;
      A.rowptr[g + 1] = (int64_t)tmp.size();
      row_off[g] = (int64_t)mine.size();
      row_thr[g] = (uint8_t)t;
      mine.insert(mine.end(), tmp.begin(), tmp.end());
    }
  }
  for (int64_t g = 0; g < dm.n_dofs; ++g) A.rowptr[g + 1] += A.rowptr[g];
  A.col.resize((size_t)A.rowptr.back());
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < dm.n_dofs; ++g) {
    const int64_t len = A.rowptr[g + 1] - A.rowptr[g];
    if (len) std::memcpy(&A.col[(size_t)A.rowptr[g]], &arena[row_thr[g]][(size_t)row_off[g]], (size_t)len * sizeof(int32_t));
  }
  return A;
}

// split a square block-concatenated CSR into block (bi,bj); columns become block-local
inline Csr extract_block(const Csr& A, const std::vector<int64_t>& block_start, int bi, int bj) {
  Csr B;
  int64_t r0 = block_start[bi], r1 = block_start[bi + 1], c0 = block_start[bj], c1 = block_start[bj + 1];
  B.n_rows = r1 - r0;
  B.n_cols = c1 - c0;
  B.rowptr.assign((size_t)B.n_rows + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t r = r0; r < r1; ++r) {
    const int32_t* b = &A.col[(size_t)A.rowptr[r]];
    const int32_t* e = b + (A.rowptr[r + 1] - A.rowptr[r]);
    B.rowptr[r - r0 + 1] = std::lower_bound(b, e, (int32_t)c1) - std::lower_bound(b, e, (int32_t)c0);
  }
  for (int64_t r = 0; r < B.n_rows; ++r) B.rowptr[r + 1] += B.rowptr[r];
  B.col.resize((size_t)B.rowptr.back());
#pragma omp parallel for schedule(static)
  for (int64_t r = r0; r < r1; ++r) {
    const int32_t* b = &A.col[(size_t)A.rowptr[r]];
    const int32_t* e = b + (A.rowptr[r + 1] - A.rowptr[r]);
    const int32_t* lo = std::lower_bound(b, e, (int32_t)c0);
    const int32_t* hi = std::lower_bound(b, e, (int32_t)c1);
    int32_t* dst = &B.col[(size_t)B.rowptr[r - r0]];
    for (const int32_t* p = lo; p < hi; ++p) *dst++ = *p - (int32_t)c0;
  }
  return B;
}

}  // namespace dcph
