// Row-owner ("written once") assembly of the classic 3-D NSE system / preconditioner matrices
// (DCP_STRATEGY_OWNER).
//
// Same integrals as assemble_th.cu / assemble_th_fast.cu (reference: include/core/boussinesq_model.tpp:421-464,
// 550-687), different ownership: instead of cells adding into shared rows with atomics, a CTA OWNS a contiguous
// range of matrix rows (a "tile" of velocity nodes, 3 rows each, or of pressure rows), accumulates every
// contribution to those rows in shared memory and streams the finished CSR values to HBM exactly once:
// no red.global, no zero-fill pass, no read-modify-write traffic.  Rows of one tile are contiguous in the CSR
// value arrays, so the write-out is one coalesced stream per block.
//
// Work list of a tile (built once per mesh on the host, dcp_owner_plan_build): for every cell touching the
// tile ("group") the local nodes a that lie in the tile ("entries"), each with a uint16 position list
// (offset of column node b inside the row, 3-bit mask of the unconstrained components).  Per group the CTA
// stages the cell's mapping record, forms the physical-gradient table once, then each warp takes entries:
// lane = column node b, 27 quadrature points, 10 accumulators (m_ab and g_ab^{dc}), and adds its 3x3 block
// into the tile accumulator.  Entries of one group have distinct row nodes, so warps never collide.
// Only contributions between unconstrained dofs are handled here; everything that involves a constrained
// dof (Dirichlet / no-normal-flux rows and columns, constrained diagonals) is added afterwards by the general
// kernel restricted to the constrained cells (assemble_th.cu, only_constrained mode).
#include <omp.h>

#include <algorithm>
#include <numeric>

#include "scatter.cuh"

namespace {
constexpr int ONU = 27, ONP = 8, ONQ = 27, OND = 89, OGS = ONQ * 13;
constexpr int POS_STRIDE = 36;      // uint16 per entry: 27 velocity column nodes + 8 pressure columns + pad
constexpr int PRE_POS_STRIDE = 84;  // preconditioner velocity entries: 3 x 27 (+pad)
constexpr int TILE_BUDGET = 23500;  // doubles of accumulator per tile
constexpr int MAX_TILE_NODES = 400;
constexpr int OTHREADS = 256;
}  // namespace

struct OwnerTiles {
  int64_t n_tiles = 0, n_groups = 0, n_entries = 0;
  int rows_per_node = 3, pos_stride = POS_STRIDE;
  int32_t* tile_node0 = nullptr;     // [n_tiles+1] first row-node of each tile
  int32_t* tile_group_ptr = nullptr; // [n_tiles+1]
  int32_t* group_cell = nullptr;     // [n_groups]
  int32_t* group_entry_ptr = nullptr;// [n_groups+1]
  uint32_t* entry_info = nullptr;    // [n_entries] a | rmask<<8 | node_local<<16
  uint16_t* entry_pos = nullptr;     // [n_entries][pos_stride]
};

struct OwnerPlan {
  OwnerTiles vel, prs;
  bool system = true;
};

namespace {

using namespace dcpdev;

struct OwnerArgs {
  const int* tile_node0;
  const int* tile_group_ptr;
  const int* group_cell;
  const int* group_entry_ptr;
  const unsigned* entry_info;
  const unsigned short* entry_pos;
  long long n_tiles;
  const double* geom;
  const double* phi_u;
  const double* dphi_u;
  const double* phi_p;
  const long long* rpA;  // velocity tiles: block(0,0); pressure tiles: block(1,0) [system] / block(1,1) [precond]
  const long long* rpB;  // velocity tiles, system: block(0,1)
  double* valA;
  double* valB;
  long long row_offset;  // first row of the tile set inside rpA (0)
  double nu;
};

struct __align__(16) dbl2 {
  double x, y;
};

// shared-memory layout (doubles unless noted)
struct OwnerSmem {
  dbl2* P0;      // [ONQ][32] (phi, G0)
  dbl2* P1;      // [ONQ][32] (G1, G2)
  double* psi;   // [ONQ][ONP]
  double* geo;   // OGS
  double* acc;   // TILE_BUDGET
  int* rbA;      // [3*MAX_TILE_NODES+1]
  int* rbB;      // [3*MAX_TILE_NODES+1]
};
constexpr size_t owner_smem_bytes() {
  return sizeof(dbl2) * ONQ * 32 * 2 + sizeof(double) * (ONQ * ONP + OGS + TILE_BUDGET) + sizeof(int) * 2 * (3 * MAX_TILE_NODES + 2);
}
__device__ __forceinline__ OwnerSmem carve(unsigned char* raw) {
  OwnerSmem s;
  s.P0 = reinterpret_cast<dbl2*>(raw);
  s.P1 = s.P0 + ONQ * 32;
  s.psi = reinterpret_cast<double*>(s.P1 + ONQ * 32);
  s.geo = s.psi + ONQ * ONP;
  s.acc = s.geo + OGS;
  s.rbA = reinterpret_cast<int*>(s.acc + TILE_BUDGET);
  s.rbB = s.rbA + 3 * MAX_TILE_NODES + 1;
  return s;
}

__device__ __forceinline__ void st_stream_f64(double* p, double v) {
  asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// stage the mapping record of `cell` and build the gradient table; all threads, three barriers
__device__ __forceinline__ void build_table(const OwnerArgs& a, const OwnerSmem& s, long long cell, int tid, int nt) {
  const double* g = a.geom + cell * OGS;
  __syncthreads();  // every warp is done with the previous group's table
  for (int i = tid; i < ONQ * 10; i += nt) s.geo[i] = g[i];   // JxW + Kinv (xq not needed for the matrix)
  __syncthreads();
  for (int i = tid; i < ONQ * ONU; i += nt) {
    const int q = i / ONU, b = i - q * ONU;
    const double r0 = __ldg(a.dphi_u + i * 3), r1 = __ldg(a.dphi_u + i * 3 + 1), r2 = __ldg(a.dphi_u + i * 3 + 2);
    double G[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      G[d] = s.geo[ONQ * (1 + d) + q] * r0 + s.geo[ONQ * (4 + d) + q] * r1 + s.geo[ONQ * (7 + d) + q] * r2;
    s.P0[q * 32 + b] = dbl2{__ldg(a.phi_u + i), G[0]};
    s.P1[q * 32 + b] = dbl2{G[1], G[2]};
  }
  __syncthreads();
}

// ---- velocity-row tiles ----------------------------------------------------------------------------------
template <bool SYSTEM>
__global__ void __launch_bounds__(OTHREADS, 1) owner_velocity_kernel(OwnerArgs a) {
  extern __shared__ __align__(16) unsigned char raw[];
  const OwnerSmem s = carve(raw);
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  for (int i = tid; i < ONQ * ONP; i += nt) s.psi[i] = a.phi_p[i];
  for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int n0 = a.tile_node0[tile], n1 = a.tile_node0[tile + 1];
    const int nrows = 3 * (n1 - n0);
    const long long baseA = a.rpA[3LL * n0];
    const long long baseB = SYSTEM ? a.rpB[3LL * n0] : 0;
    const int sizeA = (int)(a.rpA[3LL * n1] - baseA);
    const int sizeB = SYSTEM ? (int)(a.rpB[3LL * n1] - baseB) : 0;
    __syncthreads();
    for (int i = tid; i <= nrows; i += nt) {
      s.rbA[i] = (int)(a.rpA[3LL * n0 + i] - baseA);
      if (SYSTEM) s.rbB[i] = sizeA + (int)(a.rpB[3LL * n0 + i] - baseB);
    }
    for (int i = tid; i < sizeA + sizeB; i += nt) s.acc[i] = 0.0;
    for (int grp = a.tile_group_ptr[tile]; grp < a.tile_group_ptr[tile + 1]; ++grp) {
      build_table(a, s, a.group_cell[grp], tid, nt);
      const int e0 = a.group_entry_ptr[grp], e1 = a.group_entry_ptr[grp + 1];
      for (int e = e0 + warp; e < e1; e += nwarps) {
        const unsigned info = a.entry_info[e];
        const int na = info & 0xff, rmask = (info >> 8) & 7, nl = info >> 16;
        const unsigned short* pos = a.entry_pos + (size_t)e * (SYSTEM ? POS_STRIDE : PRE_POS_STRIDE);
        const int nb = lane < ONU ? lane : ONU - 1;
        double m = 0.0, g00 = 0, g01 = 0, g02 = 0, g10 = 0, g11 = 0, g12 = 0, g20 = 0, g21 = 0, g22 = 0;
#pragma unroll 3
        for (int q = 0; q < ONQ; ++q) {
          const double w = s.geo[q];
          const dbl2 a0 = s.P0[q * 32 + na], a1 = s.P1[q * 32 + na];
          const dbl2 b0 = s.P0[q * 32 + nb], b1 = s.P1[q * 32 + nb];
          const double pa = a0.x * w, ga0 = a0.y * w, ga1 = a1.x * w, ga2 = a1.y * w;
          m += pa * b0.x;
          g00 += ga0 * b0.y; g01 += ga0 * b1.x; g02 += ga0 * b1.y;
          g10 += ga1 * b0.y; g11 += ga1 * b1.x; g12 += ga1 * b1.y;
          g20 += ga2 * b0.y; g21 += ga2 * b1.x; g22 += ga2 * b1.y;
        }
        const double diag = m + a.nu * (g00 + g11 + g22);
        if (lane < ONU) {
          if (SYSTEM) {
            const unsigned p = pos[lane];
            const int off = p & 0x1fff, cmask = p >> 13;
            // value of L[(a,c),(b,d)] = delta_cd diag + nu * g[d][c]
            const double v[3][3] = {{diag + a.nu * g00, a.nu * g10, a.nu * g20},
                                    {a.nu * g01, diag + a.nu * g11, a.nu * g21},
                                    {a.nu * g02, a.nu * g12, diag + a.nu * g22}};
#pragma unroll
            for (int c = 0; c < 3; ++c)
              if (rmask & (1 << c)) {
                double* row = s.acc + s.rbA[3 * nl + c] + off;
                int k = 0;
#pragma unroll
                for (int d = 0; d < 3; ++d)
                  if (cmask & (1 << d)) row[k++] += v[c][d];
              }
          } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const unsigned p = pos[c * ONU + lane];
              if (p != 0xffffu) s.acc[s.rbA[3 * nl + c] + p] += diag;
            }
          }
        }
        if (SYSTEM && lane < ONP * 3) {
          // velocity-pressure coupling: L[(a,c),p_b] = -sum_q w d_c phi_a psi_b
          const int pb = lane / 3, c = lane - pb * 3;
          double sp = 0.0;
          for (int q = 0; q < ONQ; ++q) {
            const dbl2 a0 = s.P0[q * 32 + na], a1 = s.P1[q * 32 + na];
            const double ga = c == 0 ? a0.y : (c == 1 ? a1.x : a1.y);
            sp += s.geo[q] * ga * s.psi[q * ONP + pb];
          }
          const unsigned p = pos[ONU + pb];
          if ((rmask & (1 << c)) && p != 0xffffu) s.acc[s.rbB[3 * nl + c] + p] -= sp;
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < sizeA; i += nt) st_stream_f64(a.valA + baseA + i, s.acc[i]);
    if (SYSTEM)
      for (int i = tid; i < sizeB; i += nt) st_stream_f64(a.valB + baseB + i, s.acc[sizeA + i]);
  }
}

// ---- pressure-row tiles ----------------------------------------------------------------------------------
// system: block(1,0) rows, L[p_a,(b,d)] = -sum_q w psi_a d_d phi_b ; preconditioner: block(1,1), sum_q w psi_a psi_b
template <bool SYSTEM>
__global__ void __launch_bounds__(OTHREADS, 1) owner_pressure_kernel(OwnerArgs a) {
  extern __shared__ __align__(16) unsigned char raw[];
  const OwnerSmem s = carve(raw);
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  for (int i = tid; i < ONQ * ONP; i += nt) s.psi[i] = a.phi_p[i];
  for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int n0 = a.tile_node0[tile], n1 = a.tile_node0[tile + 1];
    const int nrows = n1 - n0;
    const long long baseA = a.rpA[n0];
    const int sizeA = (int)(a.rpA[n1] - baseA);
    __syncthreads();
    for (int i = tid; i <= nrows; i += nt) s.rbA[i] = (int)(a.rpA[n0 + i] - baseA);
    for (int i = tid; i < sizeA; i += nt) s.acc[i] = 0.0;
    for (int grp = a.tile_group_ptr[tile]; grp < a.tile_group_ptr[tile + 1]; ++grp) {
      build_table(a, s, a.group_cell[grp], tid, nt);
      const int e0 = a.group_entry_ptr[grp], e1 = a.group_entry_ptr[grp + 1];
      for (int e = e0 + warp; e < e1; e += nwarps) {
        const unsigned info = a.entry_info[e];
        const int pa = info & 0xff, nl = info >> 16;
        const unsigned short* pos = a.entry_pos + (size_t)e * POS_STRIDE;
        if (SYSTEM) {
          if (lane < ONU) {
            double s0 = 0, s1 = 0, s2 = 0;
            for (int q = 0; q < ONQ; ++q) {
              const double wp = s.geo[q] * s.psi[q * ONP + pa];
              const dbl2 b0 = s.P0[q * 32 + lane], b1 = s.P1[q * 32 + lane];
              s0 += wp * b0.y;
              s1 += wp * b1.x;
              s2 += wp * b1.y;
            }
            const unsigned p = pos[lane];
            const int off = p & 0x1fff, cmask = p >> 13;
            double* row = s.acc + s.rbA[nl] + off;
            int k = 0;
            if (cmask & 1) row[k++] -= s0;
            if (cmask & 2) row[k++] -= s1;
            if (cmask & 4) row[k++] -= s2;
          }
        } else if (lane < ONP) {
          double sp = 0.0;
          for (int q = 0; q < ONQ; ++q) sp += s.geo[q] * s.psi[q * ONP + pa] * s.psi[q * ONP + lane];
          const unsigned p = pos[ONU + lane];
          if (p != 0xffffu) s.acc[s.rbA[nl] + p] += sp;
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < sizeA; i += nt) st_stream_f64(a.valA + baseA + i, s.acc[i]);
  }
}

// ---- host-side plan builder ---------------------------------------------------------------------------------
struct HCsr {
  int64_t n_rows = 0;
  const int64_t* rp = nullptr;
  const int32_t* col = nullptr;
  int64_t len(int64_t r) const { return rp ? rp[r + 1] - rp[r] : 0; }
  int64_t find(int64_t r, int32_t c) const {
    if (!rp) return -1;
    const int32_t* b = col + rp[r];
    const int32_t* e = col + rp[r + 1];
    const int32_t* p = std::lower_bound(b, e, c);
    return (p == e || *p != c) ? -1 : p - b;
  }
  bool at(int64_t r, int64_t off, int32_t c) const { return rp && off >= 0 && rp[r] + off < rp[r + 1] && col[rp[r] + off] == c; }
};

struct HostTiles {
  std::vector<int32_t> tile_node0, tile_group_ptr, group_cell, group_entry_ptr;
  std::vector<uint32_t> entry_info;
  std::vector<uint16_t> entry_pos;
};

template <class T>
int up(dcp_ctx* ctx, T** dst, const std::vector<T>& v) {
  *dst = nullptr;
  if (v.empty()) return DCP_OK;
  if (cudaMalloc((void**)dst, v.size() * sizeof(T)) != cudaSuccess ||
      cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
    dcp_set_error("owner plan: device allocation / copy failed");
    return DCP_ERR_CUDA;
  }
  return DCP_OK;
}

int upload_tiles(dcp_ctx* ctx, const HostTiles& h, OwnerTiles& t, int rows_per_node, int pos_stride) {
  t.n_tiles = (int64_t)h.tile_node0.size() - 1;
  t.n_groups = (int64_t)h.group_cell.size();
  t.n_entries = (int64_t)h.entry_info.size();
  t.rows_per_node = rows_per_node;
  t.pos_stride = pos_stride;
  DCP_TRY(up(ctx, &t.tile_node0, h.tile_node0));
  DCP_TRY(up(ctx, &t.tile_group_ptr, h.tile_group_ptr));
  DCP_TRY(up(ctx, &t.group_cell, h.group_cell));
  DCP_TRY(up(ctx, &t.group_entry_ptr, h.group_entry_ptr));
  DCP_TRY(up(ctx, &t.entry_info, h.entry_info));
  DCP_TRY(up(ctx, &t.entry_pos, h.entry_pos));
  DCP_CUDA(cudaStreamSynchronize(ctx->stream));
  return DCP_OK;
}

void free_tiles(OwnerTiles& t) {
  cudaFree(t.tile_node0);
  cudaFree(t.tile_group_ptr);
  cudaFree(t.group_cell);
  cudaFree(t.group_entry_ptr);
  cudaFree(t.entry_info);
  cudaFree(t.entry_pos);
  t = OwnerTiles();
}

// Generic tiler: `n_nodes` row-nodes, node_size[n] accumulator doubles, adjacency node -> (cell, local) lists.
// fill_pos(entry index, node, cell, local, out pos*) returns the row mask (0 = skip entry) or -1 on failure.
template <class Fill>
bool make_tiles(int64_t n_nodes, const std::vector<int64_t>& node_size, const std::vector<int64_t>& adj_ptr,
                const std::vector<int32_t>& adj_cell, const std::vector<uint8_t>& adj_loc, int pos_stride, Fill fill_pos,
                HostTiles& out) {
  out.tile_node0.assign(1, 0);
  int64_t acc = 0;
  for (int64_t n = 0; n < n_nodes; ++n) {
    if (node_size[n] > TILE_BUDGET) return false;
    const int64_t cur_nodes = n - out.tile_node0.back();
    if (acc + node_size[n] > TILE_BUDGET || cur_nodes >= MAX_TILE_NODES) {
      out.tile_node0.push_back((int32_t)n);
      acc = 0;
    }
    acc += node_size[n];
  }
  out.tile_node0.push_back((int32_t)n_nodes);
  const int64_t nt = (int64_t)out.tile_node0.size() - 1;
  // per tile: entries sorted by cell
  std::vector<std::vector<int32_t>> t_group_cell(nt), t_group_ptr(nt);
  std::vector<std::vector<uint32_t>> t_info(nt);
  std::vector<std::vector<uint16_t>> t_pos(nt);
  bool ok = true;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t t = 0; t < nt; ++t) {
    const int64_t n0 = out.tile_node0[t], n1 = out.tile_node0[t + 1];
    struct E {
      int32_t cell;
      uint8_t loc;
      int32_t node;
    };
    std::vector<E> es;
    for (int64_t n = n0; n < n1; ++n)
      for (int64_t p = adj_ptr[n]; p < adj_ptr[n + 1]; ++p) es.push_back({adj_cell[p], adj_loc[p], (int32_t)n});
    std::stable_sort(es.begin(), es.end(), [](const E& x, const E& y) { return x.cell < y.cell; });
    std::vector<uint16_t> pos(pos_stride);
    t_group_ptr[t].push_back(0);
    for (size_t i = 0; i < es.size(); ++i) {
      std::fill(pos.begin(), pos.end(), (uint16_t)0xffff);
      const int rmask = fill_pos(es[i].node, es[i].cell, es[i].loc, pos.data());
      if (rmask < 0) {
#pragma omp atomic write
        ok = false;
        continue;
      }
      if (rmask == 0) continue;
      if (t_group_cell[t].empty() || t_group_cell[t].back() != es[i].cell) {
        if (!t_group_cell[t].empty()) t_group_ptr[t].push_back((int32_t)t_info[t].size());
        t_group_cell[t].push_back(es[i].cell);
      }
      t_info[t].push_back((uint32_t)es[i].loc | ((uint32_t)rmask << 8) | ((uint32_t)(es[i].node - n0) << 16));
      t_pos[t].insert(t_pos[t].end(), pos.begin(), pos.end());
    }
    if (!t_group_cell[t].empty()) t_group_ptr[t].push_back((int32_t)t_info[t].size());
  }
  if (!ok) return false;
  out.tile_group_ptr.assign(1, 0);
  out.group_entry_ptr.assign(1, 0);
  for (int64_t t = 0; t < nt; ++t) {
    const int32_t ebase = (int32_t)out.entry_info.size();
    for (size_t g = 0; g < t_group_cell[t].size(); ++g) {
      out.group_cell.push_back(t_group_cell[t][g]);
      out.group_entry_ptr.push_back(ebase + t_group_ptr[t][g + 1]);
    }
    out.tile_group_ptr.push_back((int32_t)out.group_cell.size());
    out.entry_info.insert(out.entry_info.end(), t_info[t].begin(), t_info[t].end());
    out.entry_pos.insert(out.entry_pos.end(), t_pos[t].begin(), t_pos[t].end());
  }
  return true;
}

}  // namespace

void dcp_owner_plan_free(OwnerPlan* p) {
  if (!p) return;
  free_tiles(p->vel);
  free_tiles(p->prs);
  delete p;
}

// Builds the tile plan for the system (nse pattern) or preconditioner (pre pattern) matrix.  Returns DCP_ERR_STATE
// (with a message) when the numbering is not node-blocked or a position check fails: the caller keeps POSITIONS.
int dcp_owner_plan_build(dcp_model* m, bool system, const dcp_model_desc* d) {
  OwnerPlan** slot = system ? &m->owner_nse : &m->owner_pre;
  *slot = nullptr;
  if (d->dim != 3 || d->family != DCP_FAMILY_CLASSIC) {
    dcp_set_error("row-owner strategy: classic 3-D family only");
    return DCP_ERR_STATE;
  }
  const int64_t nc = d->n_cells, n_u = d->nse_block_size[0], n_p = d->nse_block_size[1];
  const dcp_csr_desc(*pat)[DCP_MAX_BLOCKS] = system ? d->nse_pattern : d->pre_pattern;
  const HCsr A00{pat[0][0].n_rows, pat[0][0].rowptr, pat[0][0].col}, A01{pat[0][1].n_rows, pat[0][1].rowptr, pat[0][1].col};
  const HCsr A10{pat[1][0].n_rows, pat[1][0].rowptr, pat[1][0].col}, A11{pat[1][1].n_rows, pat[1][1].rowptr, pat[1][1].col};
  std::vector<int> sys_u(3 * ONU), sys_p(ONP);
  for (int i = 0; i < OND; ++i) {
    const int f = d->nse_local_field[i], b = d->nse_local_base[i];
    if (f < 3) sys_u[f * ONU + b] = i; else sys_p[b] = i;
  }
  std::vector<int32_t> lod((size_t)d->nse_cs.n_dofs, -1);
  for (int64_t l = 0; l < d->nse_cs.n_lines; ++l) lod[d->nse_cs.line_dof[l]] = (int32_t)l;
  if (n_u % 3 != 0) {
    dcp_set_error("row-owner strategy: velocity block is not a multiple of 3");
    return DCP_ERR_STATE;
  }
  // node-blocked numbering check + adjacency
  const int64_t n_vnodes = n_u / 3;
  bool blocked = true;
  std::vector<int64_t> vptr((size_t)n_vnodes + 1, 0), pptr((size_t)n_p + 1, 0);
  for (int64_t c = 0; c < nc && blocked; ++c) {
    const int32_t* idx = d->nse_l2g + c * OND;
    for (int a = 0; a < ONU && blocked; ++a) {
      const int32_t g0 = idx[sys_u[a]];
      blocked = g0 % 3 == 0 && g0 < n_u && idx[sys_u[ONU + a]] == g0 + 1 && idx[sys_u[2 * ONU + a]] == g0 + 2;
      if (blocked) vptr[g0 / 3 + 1]++;
    }
    for (int a = 0; a < ONP && blocked; ++a) {
      blocked = idx[sys_p[a]] >= n_u;
      if (blocked) pptr[idx[sys_p[a]] - n_u + 1]++;
    }
  }
  if (!blocked) {
    dcp_set_error("row-owner strategy: velocity components of a node are not adjacent dofs (numbering not node-blocked)");
    return DCP_ERR_STATE;
  }
  for (int64_t n = 0; n < n_vnodes; ++n) vptr[n + 1] += vptr[n];
  for (int64_t n = 0; n < n_p; ++n) pptr[n + 1] += pptr[n];
  std::vector<int32_t> vcell((size_t)vptr.back()), pcell((size_t)pptr.back());
  std::vector<uint8_t> vloc((size_t)vptr.back()), ploc((size_t)pptr.back());
  {
    std::vector<int64_t> vc(vptr.begin(), vptr.end() - 1), pc(pptr.begin(), pptr.end() - 1);
    for (int64_t c = 0; c < nc; ++c) {
      const int32_t* idx = d->nse_l2g + c * OND;
      for (int a = 0; a < ONU; ++a) {
        const int64_t p = vc[idx[sys_u[a]] / 3]++;
        vcell[p] = (int32_t)c;
        vloc[p] = (uint8_t)a;
      }
      for (int a = 0; a < ONP; ++a) {
        const int64_t p = pc[idx[sys_p[a]] - n_u]++;
        pcell[p] = (int32_t)c;
        ploc[p] = (uint8_t)a;
      }
    }
  }
  std::vector<int64_t> vsize((size_t)n_vnodes), psize((size_t)n_p);
  for (int64_t n = 0; n < n_vnodes; ++n) {
    int64_t s = 0;
    for (int c = 0; c < 3; ++c) s += A00.len(3 * n + c) + (system ? A01.len(3 * n + c) : 0);
    vsize[n] = s;
  }
  for (int64_t n = 0; n < n_p; ++n) psize[n] = system ? A10.len(n) : A11.len(n);

  auto fill_vel_system = [&](int32_t node, int32_t cell, uint8_t a, uint16_t* pos) -> int {
    const int32_t* idx = d->nse_l2g + (int64_t)cell * OND;
    int rmask = 0, c0 = -1;
    for (int c = 0; c < 3; ++c)
      if (lod[3 * node + c] < 0) {
        rmask |= 1 << c;
        if (c0 < 0) c0 = c;
      }
    if (!rmask) return 0;
    (void)a;
    for (int b = 0; b < ONU; ++b) {
      const int32_t g0 = idx[sys_u[b]];
      int cmask = 0, d0 = -1;
      for (int dd = 0; dd < 3; ++dd)
        if (lod[g0 + dd] < 0) {
          cmask |= 1 << dd;
          if (d0 < 0) d0 = dd;
        }
      if (!cmask) {
        pos[b] = 0;  // mask 0: nothing is added
        continue;
      }
      const int64_t off = A00.find(3LL * node + c0, g0 + d0);
      if (off < 0 || off > 0x1fff - 3) return -1;
      for (int c = 0; c < 3; ++c)
        if (rmask & (1 << c)) {
          int k = 0;
          for (int dd = 0; dd < 3; ++dd)
            if (cmask & (1 << dd)) {
              if (!A00.at(3LL * node + c, off + k, g0 + dd)) return -1;
              ++k;
            }
        }
      pos[b] = (uint16_t)(off | (cmask << 13));
    }
    for (int b = 0; b < ONP; ++b) {
      const int32_t gp = idx[sys_p[b]];
      if (lod[gp] >= 0) continue;  // stays 0xffff
      const int64_t off = A01.find(3LL * node + c0, (int32_t)(gp - n_u));
      if (off < 0 || off >= 0xffff) return -1;
      for (int c = 0; c < 3; ++c)
        if ((rmask & (1 << c)) && !A01.at(3LL * node + c, off, (int32_t)(gp - n_u))) return -1;
      pos[ONU + b] = (uint16_t)off;
    }
    return rmask;
  };
  auto fill_vel_precond = [&](int32_t node, int32_t cell, uint8_t, uint16_t* pos) -> int {
    const int32_t* idx = d->nse_l2g + (int64_t)cell * OND;
    int rmask = 0;
    for (int c = 0; c < 3; ++c) {
      if (lod[3 * node + c] >= 0) continue;
      rmask |= 1 << c;
      for (int b = 0; b < ONU; ++b) {
        const int32_t gc = idx[sys_u[b]] + c;
        if (lod[gc] >= 0) continue;
        const int64_t off = A00.find(3LL * node + c, gc);
        if (off < 0 || off >= 0xffff) return -1;
        pos[c * ONU + b] = (uint16_t)off;
      }
    }
    return rmask;
  };
  auto fill_prs = [&](int32_t node, int32_t cell, uint8_t, uint16_t* pos) -> int {
    const int32_t* idx = d->nse_l2g + (int64_t)cell * OND;
    if (lod[n_u + node] >= 0) return 0;
    if (system) {
      for (int b = 0; b < ONU; ++b) {
        const int32_t g0 = idx[sys_u[b]];
        int cmask = 0, d0 = -1;
        for (int dd = 0; dd < 3; ++dd)
          if (lod[g0 + dd] < 0) {
            cmask |= 1 << dd;
            if (d0 < 0) d0 = dd;
          }
        if (!cmask) {
          pos[b] = 0;
          continue;
        }
        const int64_t off = A10.find(node, g0 + d0);
        if (off < 0 || off > 0x1fff - 3) return -1;
        int k = 0;
        for (int dd = 0; dd < 3; ++dd)
          if (cmask & (1 << dd)) {
            if (!A10.at(node, off + k, g0 + dd)) return -1;
            ++k;
          }
        pos[b] = (uint16_t)(off | (cmask << 13));
      }
    } else {
      for (int b = 0; b < ONP; ++b) {
        const int32_t gp = idx[sys_p[b]];
        if (lod[gp] >= 0) continue;
        const int64_t off = A11.find(node, (int32_t)(gp - n_u));
        if (off < 0 || off >= 0xffff) return -1;
        pos[ONU + b] = (uint16_t)off;
      }
    }
    return 1;
  };

  HostTiles hv, hp;
  const int vstride = system ? POS_STRIDE : PRE_POS_STRIDE;
  bool ok = system ? make_tiles(n_vnodes, vsize, vptr, vcell, vloc, vstride, fill_vel_system, hv)
                   : make_tiles(n_vnodes, vsize, vptr, vcell, vloc, vstride, fill_vel_precond, hv);
  ok = ok && make_tiles(n_p, psize, pptr, pcell, ploc, POS_STRIDE, fill_prs, hp);
  if (!ok) {
    dcp_set_error("row-owner strategy: position verification failed (rows of a node do not share one column layout)");
    return DCP_ERR_STATE;
  }
  OwnerPlan* P = new OwnerPlan;
  P->system = system;
  int rc = upload_tiles(m->ctx, hv, P->vel, 3, vstride);
  if (rc == DCP_OK) rc = upload_tiles(m->ctx, hp, P->prs, 1, POS_STRIDE);
  if (rc != DCP_OK) {
    dcp_owner_plan_free(P);
    return rc;
  }
  *slot = P;
  return DCP_OK;
}

int dcp_launch_th_owner(dcp_model* m, const dcp_params& p, bool system) {
  OwnerPlan* P = system ? m->owner_nse : m->owner_pre;
  if (!P) {
    dcp_set_error("row-owner strategy: no plan for this model");
    return DCP_ERR_STATE;
  }
  dcp_ctx* ctx = m->ctx;
  const BlockMat& M = system ? m->nse : m->pre;
  const size_t smem = owner_smem_bytes();
  static bool attr = false;
  if (!attr) {
    DCP_CUDA(cudaFuncSetAttribute(owner_velocity_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DCP_CUDA(cudaFuncSetAttribute(owner_velocity_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DCP_CUDA(cudaFuncSetAttribute(owner_pressure_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DCP_CUDA(cudaFuncSetAttribute(owner_pressure_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  auto args_for = [&](const OwnerTiles& t) {
    OwnerArgs a{};
    a.tile_node0 = t.tile_node0;
    a.tile_group_ptr = t.tile_group_ptr;
    a.group_cell = t.group_cell;
    a.group_entry_ptr = t.group_entry_ptr;
    a.entry_info = t.entry_info;
    a.entry_pos = t.entry_pos;
    a.n_tiles = t.n_tiles;
    a.geom = m->geom_qn;
    a.phi_u = m->phi_u_qn;
    a.dphi_u = m->dphi_u_qn;
    a.phi_p = m->phi_p_qn;
    a.nu = p.dt * p.inv_re;
    return a;
  };
  {
    OwnerArgs a = args_for(P->vel);
    a.rpA = (const long long*)M.blk[0][0].rowptr;
    a.valA = M.blk[0][0].val;
    a.rpB = (const long long*)M.blk[0][1].rowptr;
    a.valB = M.blk[0][1].val;
    if (a.n_tiles) {
      const unsigned grid = (unsigned)std::min<long long>(a.n_tiles, ctx->sm_count);
      if (system)
        owner_velocity_kernel<true><<<grid, OTHREADS, smem, ctx->stream>>>(a);
      else
        owner_velocity_kernel<false><<<grid, OTHREADS, smem, ctx->stream>>>(a);
      ctx->launches++;
      DCP_CUDA(cudaGetLastError());
    }
  }
  {
    OwnerArgs a = args_for(P->prs);
    const DevCsr& B = system ? M.blk[1][0] : M.blk[1][1];
    a.rpA = (const long long*)B.rowptr;
    a.valA = B.val;
    if (a.n_tiles && B.nnz) {
      const unsigned grid = (unsigned)std::min<long long>(a.n_tiles, ctx->sm_count);
      if (system)
        owner_pressure_kernel<true><<<grid, OTHREADS, smem, ctx->stream>>>(a);
      else
        owner_pressure_kernel<false><<<grid, OTHREADS, smem, ctx->stream>>>(a);
      ctx->launches++;
      DCP_CUDA(cudaGetLastError());
    }
  }
  return DCP_OK;
}
