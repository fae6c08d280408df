// Two ranks, one GPU each, through the C ABI only (include/dcp.h): NCCL communicator from a broadcast id, ghost
// exchange, all-reduced inner product, Utilities::MPI::max -- what a C++ deal.II host would call in place of the
// Epetra_Import / MPI_Allreduce the reference hides inside Trilinos (schur_complement.hpp:143-150,
// boussinesq_model.tpp:1165, 1050).  Needs two GPUs; `halo_test` forks the second rank itself.
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dcp.h"

#define CHECK(call)                                                                      \
  do {                                                                                   \
    int rc__ = (call);                                                                   \
    if (rc__ != DCP_OK) {                                                                \
      std::fprintf(stderr, "rank %d: %s -> %d: %s\n", rank, #call, rc__, dcp_last_error()); \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

static int run_rank(int rank, int n_ranks, int fd_read, int fd_write) {
  dcp_ctx* ctx = nullptr;
  CHECK(dcp_ctx_create(rank, &ctx));
  unsigned char id[DCP_UNIQUE_ID_BYTES];
  if (rank == 0) {
    CHECK(dcp_comm_unique_id(id));
    if (write(fd_write, id, sizeof(id)) != (ssize_t)sizeof(id)) return 1;   // the host's broadcast (MPI_Bcast in deal.II)
  } else if (read(fd_read, id, sizeof(id)) != (ssize_t)sizeof(id))
    return 1;
  dcp_comm* comm = nullptr;
  CHECK(dcp_comm_create(ctx, id, rank, n_ranks, &comm));
  int r2 = -1, n2 = -1;
  CHECK(dcp_comm_info(comm, &r2, &n2));
  if (r2 != rank || n2 != n_ranks) return 1;

  // local layout: 5 owned entries then 3 ghosts; the other rank owns the ghosts (its entries 1, 3, 4)
  const int64_t n_local = 8, n_owned = 5;
  const int other = 1 - rank;
  std::vector<int32_t> send_idx = {1, 3, 4}, recv_idx = {5, 6, 7};
  std::vector<int64_t> send_counts(n_ranks, 0), recv_counts(n_ranks, 0);
  send_counts[other] = 3;
  recv_counts[other] = 3;
  dcp_halo* halo = nullptr;
  CHECK(dcp_halo_create(comm, n_local, send_idx.data(), send_counts.data(), recv_idx.data(), recv_counts.data(), &halo));
  std::vector<double> x(n_local, -1.0), y(n_local, 0.0);
  for (int i = 0; i < n_owned; ++i) {
    x[i] = 100.0 * (rank + 1) + i;
    y[i] = 0.5 * (i + 1) + rank;
  }
  double *dx = nullptr, *dy = nullptr;
  CHECK(dcp_malloc(ctx, sizeof(double) * n_local, (void**)&dx));
  CHECK(dcp_malloc(ctx, sizeof(double) * n_local, (void**)&dy));
  CHECK(dcp_memcpy_h2d(ctx, dx, x.data(), sizeof(double) * n_local));
  CHECK(dcp_memcpy_h2d(ctx, dy, y.data(), sizeof(double) * n_local));
  CHECK(dcp_halo_exchange(halo, dx));
  CHECK(dcp_ctx_synchronize(ctx));
  CHECK(dcp_memcpy_d2h(ctx, x.data(), dx, sizeof(double) * n_local));
  const double expect[3] = {100.0 * (other + 1) + 1, 100.0 * (other + 1) + 3, 100.0 * (other + 1) + 4};
  for (int k = 0; k < 3; ++k)
    if (x[5 + k] != expect[k]) {
      std::fprintf(stderr, "rank %d: ghost %d = %g, expected %g\n", rank, k, x[5 + k], expect[k]);
      return 1;
    }
  // inner product over the owned entries of both ranks
  const int64_t rb[1] = {0}, re[1] = {n_owned};
  double dot = 0.0, ref = 0.0;
  CHECK(dcp_vec_dot_allreduce(comm, 1, rb, re, dx, dy, &dot));
  for (int r = 0; r < n_ranks; ++r)
    for (int i = 0; i < n_owned; ++i) ref += (100.0 * (r + 1) + i) * (0.5 * (i + 1) + r);
  if (std::fabs(dot - ref) > 1e-12 * std::fabs(ref)) {
    std::fprintf(stderr, "rank %d: dot %.17g, expected %.17g\n", rank, dot, ref);
    return 1;
  }
  double mx[2] = {(double)rank, -(double)rank};
  CHECK(dcp_allreduce_max(comm, 2, mx));
  if (mx[0] != n_ranks - 1 || mx[1] != 0.0) return 1;
  // argument checks
  if (dcp_halo_exchange(nullptr, dx) != DCP_ERR_ARG) return 1;
  send_counts[rank] = 1;
  dcp_halo* bad = nullptr;
  if (dcp_halo_create(comm, n_local, send_idx.data(), send_counts.data(), recv_idx.data(), recv_counts.data(), &bad) != DCP_ERR_ARG) return 1;
  CHECK(dcp_halo_destroy(halo));
  CHECK(dcp_free(ctx, dx));
  CHECK(dcp_free(ctx, dy));
  CHECK(dcp_comm_destroy(comm));
  CHECK(dcp_ctx_destroy(ctx));
  std::printf("halo_test rank %d of %d: OK\n", rank, n_ranks);
  return 0;
}

int main() {
  int p[2];
  if (pipe(p) != 0) return 2;
  const pid_t child = fork();   // before any CUDA call
  if (child < 0) return 2;
  if (child == 0) {
    close(p[1]);
    return run_rank(1, 2, p[0], -1);
  }
  close(p[0]);
  const int rc0 = run_rank(0, 2, -1, p[1]);
  int status = 0;
  waitpid(child, &status, 0);
  const int rc1 = WIFEXITED(status) ? WEXITSTATUS(status) : 3;
  if (rc0 == 0 && rc1 == 0) std::printf("halo_test: OK\n");
  return rc0 != 0 ? rc0 : rc1;
}
