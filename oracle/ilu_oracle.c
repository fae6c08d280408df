/* ilu_oracle.c -- CPU restatement of ILU(0) and its triangular solves (TEST INFRASTRUCTURE ONLY).
 *
 * What the reference gets from LA::PreconditionILU (Trilinos Ifpack ILU, level of fill 0, no diagonal shift, no
 * relaxation; deal.II defaults) at include/core/boussinesq_model.tpp:1265-1275 and
 * include/linear_algebra/approximate_schur_complement.hpp:118-141.  Ifpack is not part of /root/reference; its
 * published algorithm is the textbook IKJ incomplete factorisation on the sparsity pattern of A (Saad, Iterative
 * Methods, Alg. 10.4) followed by forward / backward substitution -- restated here, sequentially.
 * Parity unpinned by the reference; pinned by the defining identity (LU)_ij = A_ij on the pattern and by exactness
 * on patterns without fill (tests/test_ilu.py). */
#include <stdint.h>
#include <string.h>

/* lu: output, same pattern as (rowptr, col); n: size of the (rank-local) square block, entries with col >= n are ignored.
 * returns 0, or 1 + row if a diagonal entry is missing */
int orc_ilu0_factor(int64_t n, const int64_t* rowptr, const int32_t* col, const double* val, double* lu) {
  memcpy(lu, val, sizeof(double) * (size_t)rowptr[n]);
  for (int64_t i = 0; i < n; ++i) {
    for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
      const int32_t k = col[p];
      if (k >= i) break;
      int64_t dk = -1;
      for (int64_t q = rowptr[k]; q < rowptr[k + 1]; ++q)
        if (col[q] == k) { dk = q; break; }
      if (dk < 0) return (int)(1 + k);
      const double lik = lu[p] / lu[dk];
      lu[p] = lik;
      int64_t t = p + 1;
      for (int64_t q = dk + 1; q < rowptr[k + 1]; ++q) {
        const int32_t j = col[q];
        if (j >= n) break;
        while (t < rowptr[i + 1] && col[t] < j) ++t;
        if (t < rowptr[i + 1] && col[t] == j) lu[t] -= lik * lu[q];
      }
    }
  }
  return 0;
}

/* y = (LU)^-1 x */
void orc_ilu0_solve(int64_t n, const int64_t* rowptr, const int32_t* col, const double* lu, const double* x, double* y) {
  for (int64_t i = 0; i < n; ++i) {
    double s = x[i];
    for (int64_t p = rowptr[i]; p < rowptr[i + 1] && col[p] < i; ++p) s -= lu[p] * y[col[p]];
    y[i] = s;
  }
  for (int64_t i = n - 1; i >= 0; --i) {
    double s = y[i], d = 1.0;
    for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
      const int32_t c = col[p];
      if (c == i) d = lu[p];
      else if (c > i && c < n) s -= lu[p] * y[c];
    }
    y[i] = s / d;
  }
}
