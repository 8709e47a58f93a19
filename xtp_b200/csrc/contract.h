// Host API of the contraction engine (see contract.cuh for the index model).
#pragma once
#include <atomic>

#include "common.h"
#include "contract.cuh"

namespace xtpb {

struct Workspace {
  DBuf buf;
  double* get(size_t count) { buf.ensure(count); return buf.p; }
};

// Fills a_vec/b_vec/splits/ws and launches.  `force_cfg`: -1 auto, 0=128x128, 1=128x64, 2=128x32 (TMA instance when
// the operands allow it and the problem is large); 4..6 the same tiles on the cp.async instance, 7 automatic tile on
// cp.async; 8..10 tile 0..2 on the TMA instance whatever the size.
// `force_splits`: 0 auto.  Returns the number of kernels launched.
int contract(GemmParams p, Workspace& ws, cudaStream_t stream, int force_cfg = -1, int force_splits = 0);
// The launcher's choices for a shape on a device with `sms` SMs (host arithmetic only): cfg / splits as in contract();
// values >= 0 / > 0 on entry are kept (forced), others are chosen.
void contract_plan(int M, int N, int K, int n_outer, int n_batch, int lower, int sms, int& cfg, int& splits);

// C[j,i] = C[i,j] for i>j (column-major n x n, leading dimension ld), optionally adds `diag_add` to the diagonal.
void symmetrize_from_lower(double* C, int n, long long ld, double diag_add, cudaStream_t stream);

// operand helpers: matrix stored column-major with leading dimension ld
inline GemmOperand op_rows_contig(const double* p, long long ld) { return GemmOperand{p, 1, ld, 0, 0}; }   // A(row,k)=p[row+k*ld]
inline GemmOperand op_k_contig(const double* p, long long ld) { return GemmOperand{p, ld, 1, 0, 0}; }      // A(row,k)=p[k+row*ld]

extern std::atomic<long long> g_launch_count;       // kernels launched by this library (bench.py's gpu_launches)
extern std::atomic<long long> g_tma_launch_count;   // of which contraction launches on the TMA instance
extern std::atomic<long long> g_tma5d_launch_count; // of which with a single-box (5-D map) row-contiguous operand

// ---- in-library kernel timing (bench.py's roofline object): CUDA-event pairs on the launching stream around
// every launch of a tagged kernel family; summed per tag.  Off by default (no events recorded).
enum ProfTag {
  PROF_OTHER = 0, PROF_FILL = 1, PROF_ROTATE = 2, PROF_EPSILON = 3, PROF_SIGMA_X = 4, PROF_SIGMA_OFFDIAG = 5,
  PROF_BSE_MATMUL = 6, PROF_DAVIDSON = 7, PROF_DENSE_AUX = 8, PROF_SIGMA_GRID = 9, PROF_SIGMA_PAIRS = 10,
  PROF_SOLVER = 11, PROF_UNPACK = 12, PROF_CDA = 13, PROF_EXACT = 14, PROF_COMM = 15,
  PROF_SIGMA_POINTS = 16, PROF_NTAGS = 17
};
void prof_enable(bool on);
void prof_reset();
bool prof_enabled();
// returns a slot (or -1 when disabled / pool exhausted); `work` = algorithmic flops (or bytes) of the launch
int prof_begin(int tag, double work, cudaStream_t s);
void prof_end(int slot, cudaStream_t s);
// device must be idle (caller synchronises); sums elapsed ms / work / launches for `tag`
void prof_get(int tag, double* ms, double* work, long long* launches);
struct ProfScope {     // tags every contract() call issued while it is alive (thread-local)
  int prev;
  explicit ProfScope(int tag);
  ~ProfScope();
};

}  // namespace xtpb
