// Device-side AffineConstraints::distribute_local_to_global (general path).
//
// Restates the scatter the reference performs in its serial WorkStream copier
// (include/core/boussinesq_model.tpp:468-476, 677-687, 804-817, 955-964) for a local matrix that lives in
// shared memory: exact-zero local entries are skipped, constrained rows/columns are redistributed to their
// master dofs with weights, inhomogeneities are moved to the right-hand side, and every constrained dof
// receives |L_ii| (mean |diag| if zero) on its own diagonal.  Adds go through red.global.add.f64; the
// column position is found by binary search in the CSR row.
#pragma once
#include "dcp_internal.cuh"

namespace dcpdev {

// cp.async (LDGSTS): asynchronous global -> shared copies of 4 or 8 bytes per thread
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// D(8x8) += A(8x4) B(4x8) on the FP64 tensor cores.  Fragments: a = A[lane/4][lane%4], b = B[lane%4][lane/4],
// c0/c1 = D[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void red_add_f64(double* addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

__device__ __forceinline__ int block_of(const BlockView& A, int g) {
  int b = 0;
#pragma unroll
  for (int k = 1; k < DCP_MAXB; ++k)
    if (k < A.nb && g >= A.start[k]) b = k;
  return b;
}

// add v at global (gr, gc); *err is bumped when the pattern has no such entry
__device__ __forceinline__ void block_csr_add(const BlockView& A, int gr, int gc, double v, int* err) {
  const int bi = block_of(A, gr), bj = block_of(A, gc);
  const long long* rp = A.rowptr[bi][bj];
  if (rp == nullptr) {
    atomicAdd(err, 1);
    return;
  }
  const long long r = gr - A.start[bi];
  const int c = gc - (int)A.start[bj];
  const int* cols = A.col[bi][bj];
  long long lo = rp[r], hi = rp[r + 1];
  const long long end = hi;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (cols[mid] < c) lo = mid + 1; else hi = mid;
  }
  if (lo == end || cols[lo] != c) {
    atomicAdd(err, 1);
    return;
  }
  red_add_f64(A.val[bi][bj] + lo, v);
}

// Cooperative distribute of an n x n local matrix (row-major, leading dimension ld) and optional local rhs.
// `lines` is an n-int scratch array in shared memory.  `only_constrained`: add only the contributions that
// involve at least one constrained local dof (used as the fix-up pass of the row-owner strategy).
// Must be called by all `nthreads` threads of the group; contains __syncthreads-free code only when
// nthreads <= 32 (GROUP_SYNC selects __syncwarp vs __syncthreads).
template <bool WARP_GROUP>
__device__ __forceinline__ void group_sync() {
  if (WARP_GROUP) __syncwarp(); else __syncthreads();
}

template <bool WARP_GROUP>
__device__ void distribute_local_matrix(const CsView& cs, int n, int ld, const double* L, const double* l,
                                        const int* idx, int* lines, const BlockView& A, double* b, int tid,
                                        int nthreads, bool only_constrained, int* err) {
  for (int i = tid; i < n; i += nthreads) lines[i] = cs.line_of_dof[idx[i]];
  group_sync<WARP_GROUP>();
  bool any = false;
  for (int i = 0; i < n; ++i) any |= (lines[i] >= 0);
  for (int e = tid; e < n * n; e += nthreads) {
    const int i = e / n, j = e - i * n;
    const double v = L[i * ld + j];
    if (v == 0.0) continue;
    const int li = lines[i], lj = lines[j];
    if (li < 0 && lj < 0) {
      if (!only_constrained) block_csr_add(A, idx[i], idx[j], v, err);
      continue;
    }
    int one_r = idx[i], one_c = idx[j];
    double one_w = 1.0;
    const int* rd = &one_r;
    const double* rw = &one_w;
    int nr = 1;
    if (li >= 0) {
      nr = cs.line_ptr[li + 1] - cs.line_ptr[li];
      rd = cs.entry_dof + cs.line_ptr[li];
      rw = cs.entry_w + cs.line_ptr[li];
    }
    const int* cd = &one_c;
    const double* cw = &one_w;
    int nc = 1;
    if (lj >= 0) {
      nc = cs.line_ptr[lj + 1] - cs.line_ptr[lj];
      cd = cs.entry_dof + cs.line_ptr[lj];
      cw = cs.entry_w + cs.line_ptr[lj];
    }
    for (int a = 0; a < nr; ++a)
      for (int c = 0; c < nc; ++c) block_csr_add(A, rd[a], cd[c], rw[a] * cw[c] * v, err);
    if (b != nullptr && lj >= 0) {
      const double ih = cs.inhom[lj];
      if (ih != 0.0)
        for (int a = 0; a < nr; ++a) red_add_f64(b + rd[a], -rw[a] * v * ih);
    }
  }
  if (b != nullptr && l != nullptr)
    for (int i = tid; i < n; i += nthreads) {
      const int li = lines[i];
      if (li < 0) {
        if (!only_constrained) red_add_f64(b + idx[i], l[i]);
      } else
        for (int k = cs.line_ptr[li]; k < cs.line_ptr[li + 1]; ++k) red_add_f64(b + cs.entry_dof[k], cs.entry_w[k] * l[i]);
    }
  if (any) {
    double avg = 0.0;
    for (int i = 0; i < n; ++i) avg += fabs(L[i * ld + i]);
    avg /= n;
    for (int i = tid; i < n; i += nthreads)
      if (lines[i] >= 0) {
        const double d = fabs(L[i * ld + i]);
        block_csr_add(A, idx[i], idx[i], d != 0.0 ? d : avg, err);
      }
  }
}

// AffineConstraints::distribute_local_to_global(local_vector, indices, global_vector, local_matrix) --
// the matrix_for_bc overload (boussinesq_model.tpp:960-963).  Lbc is n x n row-major (ld).
template <bool WARP_GROUP>
__device__ void distribute_local_vector_bc(const CsView& cs, int n, int ld, const double* l, const double* Lbc,
                                           const int* idx, const int* lines, double* b, int tid, int nthreads) {
  for (int i = tid; i < n; i += nthreads) {
    const int li = lines[i];
    if (li < 0) {
      red_add_f64(b + idx[i], l[i]);
      continue;
    }
    for (int k = cs.line_ptr[li]; k < cs.line_ptr[li + 1]; ++k) red_add_f64(b + cs.entry_dof[k], l[i] * cs.entry_w[k]);
  }
  // eliminate the columns of inhomogeneously constrained dofs
  for (int e = tid; e < n * n; e += nthreads) {
    const int i = e / n, j = e - i * n;  // i: constrained column, j: row
    const int li = lines[i];
    if (li < 0) continue;
    const double val = cs.inhom[li];
    if (val == 0.0) continue;
    const double mji = Lbc[j * ld + i];
    const int lj = lines[j];
    if (lj < 0) {
      red_add_f64(b + idx[j], -val * mji);
      continue;
    }
    if (mji == 0.0) continue;
    for (int k = cs.line_ptr[lj]; k < cs.line_ptr[lj + 1]; ++k)
      red_add_f64(b + cs.entry_dof[k], -val * cs.entry_w[k] * mji);
  }
}

}  // namespace dcpdev
