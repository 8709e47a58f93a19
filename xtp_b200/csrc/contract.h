// Host API of the contraction engine (see contract.cuh for the index model).
#pragma once
#include "common.h"
#include "contract.cuh"

namespace xtpb {

struct Workspace {
  DBuf buf;
  double* get(size_t count) { buf.ensure(count); return buf.p; }
};

// Fills a_vec/b_vec/splits/ws and launches.  `force_cfg`: -1 auto, 0=128x128, 1=128x64, 2=128x32.
// `force_splits`: 0 auto.  Returns the number of kernels launched.
int contract(GemmParams p, Workspace& ws, cudaStream_t stream, int force_cfg = -1, int force_splits = 0);

// C[j,i] = C[i,j] for i>j (column-major n x n, leading dimension ld), optionally adds `diag_add` to the diagonal.
void symmetrize_from_lower(double* C, int n, long long ld, double diag_add, cudaStream_t stream);

// operand helpers: matrix stored column-major with leading dimension ld
inline GemmOperand op_rows_contig(const double* p, long long ld) { return GemmOperand{p, 1, ld, 0, 0}; }   // A(row,k)=p[row+k*ld]
inline GemmOperand op_k_contig(const double* p, long long ld) { return GemmOperand{p, ld, 1, 0, 0}; }      // A(row,k)=p[k+row*ld]

extern long long g_launch_count;   // kernels launched by this library (bench.py's gpu_launches)

}  // namespace xtpb
