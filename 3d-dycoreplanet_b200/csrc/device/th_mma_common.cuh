// Shared definitions of the classic 3-D Taylor-Hood tensor-core kernels (assemble_th_mma.cu: reductions into the CSR
// values; assemble_th_stage.cu: write-once path through a node-major staging buffer and a TMA-fed row-owner gather).
#pragma once
#include "scatter.cuh"

namespace thmma {

using namespace dcpdev;

constexpr int NU = 27, NP = 8, NQ = 27, ND = 89, NE = 35, GS = NQ * 13;
constexpr int LDB = 148;        // row stride of X (148 mod 16 == 4: conflict-free fragment loads)
constexpr int KQ = 28;          // quadrature points padded to a multiple of 4
constexpr int PSI0 = 128;       // first psi column (node columns 32*alpha + b, b < 32)
constexpr int PSTR = 1228;       // plan row: NE*NE offsets padded to a multiple of 8 bytes (cp.async granularity)
constexpr int MSTR = 48;         // mask row: 27 node masks, 8 pressure flags, cell flag, int32 index of the wide table,
                                 // int32 index of the 9-combination table (no-normal-flux cells of the preconditioner)
constexpr int IDS = 92;          // dof index buffer stride
constexpr int MTHREADS = 128;    // 4 warps per CTA, 4 CTAs per SM: several cells in flight per SM hide the per-cell load latency

struct MmaArgs {
  long long n_fast;
  const int* wlist;                 // optional: plan indices to process (n_fast of them) instead of 0 .. n_fast-1
  const int* cells;
  const unsigned short* pos;
  const unsigned char* nmask;       // [n][MSTR]: unconstrained-component mask of the 27 velocity nodes, 8 pressure flags,
                                    // cell flag, then (preconditioner) int32: -1 or index into pos_wide
  const unsigned short* pos_wide;   // [n_wide][3][27][27]
  const unsigned short* pos9;       // [n_nnf][9][27][27]: preconditioner cells with no-normal-flux lines
  const double* geom;
  const int* l2g;
  const int* l2g_t;
  const int* local_field;
  const int* local_base;
  const double* phi_u;
  const double* dphi_u;
  const double* phi_p;
  const double* phi_t;
  int ndt;
  const double* old_nse;
  const double* old_temp;
  double* rhs;
  long long n_u;
  dcp_params prm;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// shared by the node x node and node x psi epilogues: constraint data of one velocity node
struct NodeCs {
  int mask;   // free components
  int k;      // component constrained with masters (no-normal-flux), 3 = none
  double w0, w1, w2;
};

}  // namespace thmma
