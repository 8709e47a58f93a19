// Internal C++ objects behind the C ABI handles of include/xtpb200/xtpb200.h.
#pragma once
#include <cusolverDn.h>

#include <exception>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/xtpb200/xtpb200.h"
#include "common.h"
#include "contract.h"

namespace xtpb {

// ---------------------------------------------------------------- context
struct Context {
  int device = 0;
  cudaStream_t stream = nullptr;
  Workspace ws;                 // split-K partials
  cusolverDnHandle_t solver = nullptr;
  DBuf solver_work;
  DBuf scratch_a, scratch_b;    // transient operands (fill blocks, rotation targets)
  int* dev_info = nullptr;
  double solver_seconds = 0.0;
  int solver_prof_slot = -1;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // multi-GPU: one process per GPU, this context is rank `rank` of `world` (comm.cu).  world == 1: no collectives.
  int rank = 0, world = 1;
  void* nccl = nullptr;         // ncclComm_t
  cudaStream_t comm_stream = nullptr;
  double comm_seconds = 0.0;    // not timed per call; collectives are stream-ordered

  // helper-thread resources (created on first use): a second, high-priority stream with its own cuSOLVER handle, on
  // which a dense eigensolver runs underneath independent work of the main stream
  // (TCMatrix::metric_prefetch_begin under Fill3cMO; eigh_async under the second epsilon of the plasmon-pole model)
  cudaStream_t side_stream = nullptr;
  cusolverDnHandle_t side_solver = nullptr;
  DBuf side_work;
  int* side_info = nullptr;
  cudaEvent_t side_ready = nullptr;
  void side_init();
  struct AsyncEigh {
    std::thread th;
    std::exception_ptr err;
    bool active = false;
  } async_eigh;
  // A (n x n, ld = lda, on the device, produced by work already enqueued on `stream`) <- eigenvectors, w <- eigenvalues
  // (device), lam_host <- eigenvalues; returns at once, eigh_async_join() waits and rethrows
  void eigh_async_begin(int n, double* A, long long lda, double* w, double* lam_host);
  void eigh_async_join();
  explicit Context(int dev);
  ~Context();
  void sync() { XTPB_CUDA(cudaStreamSynchronize(stream)); }
  void h2d(double* dst, const double* src, size_t n) {
    XTPB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, stream));
  }
  void d2h(double* dst, const double* src, size_t n) {
    XTPB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    sync();
  }
  void h2d_2d(double* dst, long long ldd, const double* src, long long lds, long long rows, long long cols) {
    XTPB_CUDA(cudaMemcpy2DAsync(dst, ldd * 8, src, lds * 8, rows * 8, cols, cudaMemcpyHostToDevice, stream));
  }
  void d2h_2d(double* dst, long long ldd, const double* src, long long lds, long long rows, long long cols) {
    XTPB_CUDA(cudaMemcpy2DAsync(dst, ldd * 8, src, lds * 8, rows * 8, cols, cudaMemcpyDeviceToHost, stream));
    sync();
  }
  // dense solver calls (cuSOLVER; timed separately from the contraction kernels)
  void eigh(int n, double* A, long long lda, double* w);            // A <- eigenvectors (ascending w)
  void spd_inverse(int n, double* A, long long lda);                // in place, full symmetric output
  bool cholesky(int n, double* A, long long lda, bool upper);       // A <- factor in that triangle; false: not pos. def.
  void tri_inverse(int n, double* T, long long lda, bool upper);    // that triangle <- its inverse (rest untouched)
  void general_inverse(int n, double* A, long long lda, double* Ainv, long long ldi);   // A destroyed
  void lu_solve_vector(int n, double* A, long long lda, double* x);                     // x <- A^-1 x, A destroyed
  void solver_begin();
  void solver_end();
  // collectives on `st` (default: the compute stream), FP64, in place; no-ops when world == 1
  void comm_init(const char* unique_id_128, int rank_, int world_);
  void comm_destroy();
  void allreduce_sum(double* buf, size_t count, cudaStream_t st = nullptr);
  void allgather(const double* send, double* recv, size_t count_per_rank, cudaStream_t st = nullptr);
  void reduce_sum(double* buf, size_t count, int root, cudaStream_t st = nullptr);     // result on `root` only
  void bcast(double* buf, size_t count, int root, cudaStream_t st = nullptr);
  void group_start();           // NCCL group: the collectives in between are issued as one batch
  void group_end();
};
void comm_unique_id(char* out_128);
// hostlinalg.cu: symmetric eigenproblem on the host (A: n x n col-major, lower triangle read; A <- eigenvectors,
// w ascending).  false = QL iteration did not converge.
bool host_eigh(int n, double* A, double* w);
// hostlinalg.cu: Anderson mixing (upstream xtp/src/libxtp/anderson_mixing.cc)
class Anderson {
 public:
  void configure(int order, double alpha);
  void update_input(const std::vector<double>& x);
  void update_output(const std::vector<double>& x);
  std::vector<double> mix_history() const;
 private:
  int order_ = 1;
  double alpha_ = 0.7;
  std::vector<std::vector<double>> input_, output_;
};

// ---------------------------------------------------------------- small kernels (kernels.cu)
void k_chi0_weights(double* d, const double* e_m, const double* e_n, int n_occ, int n_occ_n, int a0, int K,
                    const double* omegas, int n_omega, bool imag, double eta, cudaStream_t s);
void k_set_identity(double* A, int n, long long ld, cudaStream_t s);
void k_add_diagonal(double* A, int n, long long ld, double v, cudaStream_t s);
void k_scale_columns(double* A, int rows, int cols, long long ld, const double* scale, cudaStream_t s);  // A(:,j)*=scale[j]
void k_extract_diagonal(const double* A, int n, long long ld, double* out, cudaStream_t s);
void k_zero_strict_lower(double* A, int n, long long ld, cudaStream_t s);   // A(i,j) = 0 for i > j
void k_copy_2d(double* dst, long long ldd, const double* src, long long lds, int rows, long long cols, cudaStream_t s);
void k_scale(double* x, long long n, double a, cudaStream_t s);
void k_axpby(double* y, const double* x, long long n, double a, double b, cudaStream_t s);   // y = a*x + b*y
// window extraction with optional per-P scaling: dst[i][P][j] = scale[P] * M[(m0+i)][P][n0+j]
void k_extract_window(double* dst, long long dst_ld, long long dst_slab, const double* M, long long ldn, long long slab,
                      int m0, int mcnt, int n0, int ncnt, int naux, const double* scale, cudaStream_t s);
// Sigma_c (PPM): see kernels.cu
void k_sigma_ppm_grid(const double* M, long long ldn, long long slab, int ntotal, int naux, int n_occ,
                      const double* energies, const double* ppm_freq, const double* ppm_fac, const int* level_slab,
                      const double* omega0, double domega, int n_omega, int n_levels, double* values, cudaStream_t s);
// compressed grid scan (kernels.cu (1b)): far bins of the pole axis through Chebyshev moments, near bins pole by pole;
// edges_host: nb+1 bin edges, near_host: [level][chunk][4] inclusive near-bin and inner-bin ranges (ppm_grid_plan).
// The target-independent part (bin table, moments, equivalent poles) is kept in a PpmScanState so that further points
// of the same levels (bisection rounds, final Sigma_c) are evaluated without streaming the slabs again.
struct PpmScanState {
  bool valid = false;
  unsigned long long generation = 0;   // TCMatrix::generation the moments were built from
  int nb = 0, n_levels = 0;
  std::vector<double> edges;           // host copy of the bin edges (nb + 1)
  DBuf edges_dev, table, mom, eq;
};
void k_ppm_scan_prepare(const double* M, long long ldn, long long slab, int ntotal, int naux, int n_occ,
                        const double* energies, const double* ppm_freq, const double* ppm_fac, const int* level_slab,
                        int n_levels, const double* edges_host, int nb, PpmScanState& st, cudaStream_t s);
void k_ppm_scan_evaluate(const double* M, long long ldn, long long slab, int naux, const double* energies,
                         const double* ppm_freq, const double* ppm_fac, const int* level_slab, const int* level_mom,
                         const double* omega0, double domega, int n_omega, int n_items, const int* near_host,
                         int n_chunks, const PpmScanState& st, double* values, double* direct_evaluations,
                         cudaStream_t s);
void k_sigma_ppm_grid_compressed(const double* M, long long ldn, long long slab, int ntotal, int naux, int n_occ,
                                 const double* energies, const double* ppm_freq, const double* ppm_fac,
                                 const int* level_slab, const double* omega0, double domega, int n_omega, int n_levels,
                                 const double* edges_host, int nb, const int* near_host, int n_chunks, double* values,
                                 double* direct_evaluations, PpmScanState& st, cudaStream_t s);
// host plan of the compressed scan: bins of the pole axis and, per level and chunk of 32 grid points, the bins that
// are too close for the series.  grid_start[level] = first grid frequency; [zmin, zmax] = range of the live poles.
// Returns false when the scan should use the direct kernel (no poles, too many bins).
struct PpmGridPlan {
  std::vector<double> edges;    // nb + 1, ascending
  std::vector<int> near;        // [level][chunk][4]: near bins lo..hi (inclusive), of which inner bins ilo..ihi
  int nb = 0, n_chunks = 0;
};
constexpr double kPpmDampingWindow = 0.25;   // |x| below which Sigma_PPM::Stabilize damps 1/x (sigma_ppm.cc)
constexpr double kPpmGridBinWidth = 0.125;   // width of the core bins of the pole axis [Ha]
constexpr int kPpmGridChunk = 8;             // grid points per warp of the compressed scan (see kernels.cu (1b))
bool ppm_grid_plan(const double* grid_start, long long n_levels, double spacing, long long steps, double zmin,
                   double zmax, PpmGridPlan& plan);
// near / inner bin ranges (out[0..3], as in PpmGridPlan::near) of the targets [wa, wb] for the bins `edges` (nb + 1)
void ppm_grid_near(const double* edges, int nb, double wa, double wb, int* out);
void k_sigma_ppm_pairs(const double* M, long long ldn, long long slab, int ntotal, int naux, int n_occ,
                       const double* energies, const double* ppm_freq, const double* ppm_fac, const int* pair_slab,
                       const double* pair_omega, int n_pairs, double* values, double* derivs, double* partial,
                       cudaStream_t s);
void k_sigma_ppm_weighted_slab(double* W, const double* M, long long ldn, long long slab, int ntotal, int p0, int pcnt,
                               int n_occ, const double* energies, const double* ppm_freq, const double* ppm_fac,
                               int slab0, int n_levels, const double* level_omega, cudaStream_t s);
void k_bse_diagonal(double* diag, int vt, int ct, int naux, const double* Mvc, long long ldvc, long long slabvc,
                    const double* Mvv, long long ldvv, long long slabvv, const double* Mcc, long long ldcc,
                    long long slabcc, const double* Mcv, long long ldcv, long long slabcv, const double* eps_inv,
                    const double* hqp_diag, int cqp, int cx, int cd, int cd2, cudaStream_t s);
// d[i] += alpha * sum_{p < rows} F[p*ld + i]^2, i < n   (exchange part of the BSE diagonal from the flat operand)
void k_add_column_square_sums(double* d, const double* F, long long ld, int rows, long long n, double alpha,
                              cudaStream_t s);
void k_col_norms(const double* A, long long ld, long long rows, int cols, double* out, cudaStream_t s);
// out[j] = sum_i A[i + j*lda] * B[i + j*ldb], j < cols
void k_column_dots(double* out, const double* A, long long lda, const double* B, long long ldb, long long rows, int cols,
                   cudaStream_t s);
void k_residuals(double* res, long long ldr, const double* q, long long ldq, const double* lambda, long long rows,
                 int cols, cudaStream_t s);                            // res(:,j) -= lambda[j]*q(:,j)
void k_davidson_correction(double* out, const double* r, const double* x, const double* D, double lambda, long long n,
                           int olsen, double* scratch2, cudaStream_t s);
void k_unit_vectors(double* V, long long ld, long long n, const long long* idx, int cols, cudaStream_t s);
// dense BSE Hamiltonian: H[(v1,c1),(v2l,c2)] += cqp (Hqp[vt+c1,vt+c2] d(v1,v2) - Hqp[v1,v2] d(c1,c2)) for the local columns
void k_bse_add_hqp(double* H, long long ld, int vt, int ct, int v2lo, int ns, const double* hqp, long long hs,
                   double cqp, cudaStream_t s);
// dense BSE direct term over the occupied pairs v1 <= v2 only (t = v2(v2+1)/2 + v1): packed pair columns of the flat
// operand [P][v2][v1], and the scatter of the contracted pair blocks T[(c1,c2)][t] into H and its mirror image
void k_bse_pack_pairs(double* Ftri, long long ldt, const double* F, long long ldf, int vt, int naux, cudaStream_t s);
void k_bse_scatter_pairs(double* H, long long ld, int ct, const double* T, long long t0, long long tcnt, cudaStream_t s);
// full[b](mu,nu) = full[b](nu,mu) = packed[b][mu(mu+1)/2 + nu] (nu <= mu), b < count
void k_unpack_symmetric(double* full, long long ld, long long full_slice, const double* packed, long long pk_slice,
                        int n, int count, cudaStream_t s);
// dst[j] = src[first + j*stride], j < n  /  dst[first + j*stride] = src[j]
void k_gather_strided(double* dst, const double* src, long long first, long long stride, long long n, cudaStream_t s);
// cyclic column maps between a full row layout (ld_full, global columns) and a local one (ld_loc): rows = naux*count
void k_cols_full_to_local(double* loc, long long ld_loc, const double* full, long long ld_full, long long rows,
                          long long ncols_loc, int rank, int world, cudaStream_t s);
void k_cols_local_to_full(double* full, long long ld_full, const double* loc, long long ld_loc, long long rows,
                          long long ncols_loc, int rank, int world, cudaStream_t s);
// collective Fill3cMO, second half: M[m][P(slot)][n] = stage[m][slot][n] for slot = s*B + idx < world*B with
// P = naux*s/world + round*B + idx inside rank s's aux range (other slots are padding of the last round)
void k_scatter_fill_slots(double* M, long long ldn, long long slab, const double* stage, int mtotal, int ntotal,
                          int naux, int world, int B, int round, cudaStream_t s);
// BSE window re-shard: dst[i][Ql][j] = G[s(j)][i][P0+Ql][jl(j)], gathered blocks of [mcnt][naux][ldl]
void k_window_from_gathered(double* dst, long long dst_ld, long long dst_slab, const double* G, long long ldl,
                            int mcnt, int naux, int P0, int pcnt, int n0, int ncnt, int world, cudaStream_t s);

// ---------------------------------------------------------------- TCMatrix_gwbse
struct TCMatrix {
  Context* ctx;
  long long naux, mmin, mmax, nmin, nmax, mtotal, ntotal;
  long long ldn, slab;          // device layout [m][P][ldn], ldn = ntotal rounded up to even
  unsigned long long generation = 0;   // bumped by everything that writes the tensor (caches keyed on its contents)
  unsigned long long content_gen = 0;  // ... except a change of the pending factor only (the stored elements stay)
  // Multi-GPU: the second index is distributed cyclically, rank r holds the columns n = r, r + world, ... of
  // nmin..nmax.  `ntotal` is the LOCAL column count (== ntotal_glob when world == 1); every stage that sums over
  // the second index produces a partial result that is all-reduced.
  int rank = 0, world = 1;
  long long ntotal_glob;
  long long nloc_below(long long g) const { return g > rank ? (g - rank + world - 1) / world : 0; }
  long long nglob(long long jl) const { return rank + jl * world; }
  DBuf M;
  // Fill state
  long long n_basis = 0, ldc = 0;
  DBuf Cm, Cn;                  // MO coefficient blocks (n_basis x mtotal / ntotal, ld = ldc)
  DBuf stage2[2], unpacked;     // host->device staging of AO slices (double-buffered) / unpacked symmetric slices
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};

  TCMatrix(Context* c, long long auxsize, long long mmin_, long long mmax_, long long nmin_, long long nmax_);
  double* slab_ptr(long long m) { return M.p + m * slab; }
  void set_raw(const double* host);
  void get_slab(long long m, double* host);
  void fill_begin(long long nb, const double* C_host, long long ldc_host);
  void fill_block_dev(long long P0, long long nP, const double* ao_dev, long long ld_ao);
  void fill_block_host(long long P0, long long nP, const double* ao_host, long long ld_ao, bool packed);
  void fill_block_packed_dev(long long P0, long long nP, const double* packed_dev);
  // collective Fill3cMO: this rank contributes the packed AO slices of its canonical aux range
  // [naux*rank/world, naux*(rank+1)/world) (host or device pointer); half-transformed blocks are all-gathered
  void fill_sharded_packed(const double* packed, bool on_device);
  void aux_range(int r, long long& lo, long long& hi) const { lo = naux * r / world; hi = naux * (r + 1) / world; }
  // energies of the local second-index columns (returns e_glob_dev itself when world == 1)
  const double* local_energies(const double* e_glob_dev, DBuf& tmp);
  ~TCMatrix();
  TCMatrix(TCMatrix&&) = delete;
  // M[m] <- M[m] * R for all m (R on the device, naux x naux, ld = ldr)
  // covariant = true: R was derived from the tensor THROUGH the pending factor (the eigenvectors of an epsilon formed
  // with it), so M (Rp R) is the same for every Rp with Rp Rp^T = V^-1 and a pending Cholesky factor may be folded in;
  // any other R (a caller's matrix) meets the reference's symmetric factor: a pending Cholesky factor is flushed first
  void rotate(const double* R_dev, long long ldr, bool covariant = false);
  // Deferred aux rotation.  The Coulomb-metric factor V^-1/2 (xtpb_tc_apply_coulomb_metric) is not applied to the
  // 29 GB tensor at once: the next full rotation (the PPM eigenvectors) is folded into it (M <- M (Rp R), one pass
  // instead of two), epsilon is formed from the un-rotated tensor and sandwiched (Rp^T E Rp, two N_aux^3 products),
  // and every other consumer calls flush() first, so the observable tensor is unchanged.
  DBuf pendR;
  bool pending = false;
  void set_pending(const double* R_dev, long long ldr);
  void flush();
  // After the PPM rotation (Sigma_PPM::PrepareScreening) the aux basis of the tensor IS the eigenbasis of eps(0) for
  // the RPA input energies of that call: eps(0) = diag(lambda).  G0W0 hands the same energies to
  // BSE::SetupDirectInteractionOperator, whose eps(0) + eigensolver + rotation then reduce to reading lambda
  // (bse_setup_screening).  Cleared by anything that changes the tensor's aux basis or contents.
  struct Eps0Basis {
    bool valid = false;
    std::vector<double> energies, lambda;
    double eta = 0.0;
    long long n_occ = 0;
  } eps0;
  // First eigendecomposition of the Coulomb-metric step (of the aux overlap if one is given, else of the Coulomb
  // matrix), started on a helper thread with its own stream and cuSOLVER handle BEFORE Fill3cMO so that the two
  // overlap: their inputs are independent.  xtpb_tc_apply_coulomb_metric joins it.
  struct MetricPrefetch {
    std::thread th;
    std::exception_ptr err;
    DBuf U, w;
    std::vector<double> lam;
    bool active = false, of_overlap = false;
    const double* src = nullptr;
  } prefetch;
  void metric_prefetch_begin(const double* X_host, long long ldx, bool of_overlap);
  bool metric_prefetch_join();      // true when a prefetched decomposition is available in prefetch.U / lam
  // AOCoulomb::Pseudo_InvSqrt_GWBSE + MultiplyRightWithAuxMatrix (second half of TCMatrix_gwbse::Fill).  Returns the
  // number of removed functions.  The factor R (R R^T = V^-1 on the kept space) becomes the pending right factor.
  //  * eigen path: R = [S^-1/2] (S^-1/2 V S^-1/2)^-1/2 from one or two N_aux eigendecompositions, eigenvalues below
  //    etol dropped -- the reference's construction;
  //  * Cholesky path (default whenever it is exact): when S - etol and V - etol S (V - etol without an overlap) are
  //    positive definite -- two Cholesky factorisations decide that rigorously -- no function is removed by either
  //    decomposition and ANY R with R R^T = V^-1 gives the same epsilon spectrum, the same PPM-rotated tensor
  //    M R Phi (Phi absorbs the orthogonal factor between two choices of R) and the same Sigma_x; R = L^-T from
  //    V = L L^T replaces ~0.2 s of latency-bound eigensolver per decomposition by ~0.04 s of Cholesky work.  What the
  //    host can observe is unchanged: if anything needs the metric-rotated tensor itself (flush()), the symmetric
  //    factor is computed then, from retained copies of V and S.  XTPB_METRIC_CHOLESKY=0 disables the path.
  long long apply_coulomb_metric(const double* V_host, long long ldv, const double* S_host, long long lds, double etol);
  void metric_hint(const double* V_host, long long ldv, const double* S_host, long long lds);
  long long metric_factor_eig(double* V_dev, double* S_dev, double etol, bool prefetched, DBuf& R_out);
  struct MetricSources {            // Cholesky path: what flush() needs to build the symmetric factor after all
    bool cholesky = false, has_S = false;
    DBuf V, S;
    double etol = 0.0;
  } metric_src;
  // Optional: accumulate the two epsilon matrices of Sigma_PPM::PrepareScreening (w = 0 on the real axis, w = 0.5 on the
  // imaginary axis) WHILE Fill3cMO runs: as soon as the tensor rows of another 256 aux functions are complete, the panel
  // E[P0..P1) x [0..P1) is contracted (same flops as the one SYRK-shaped launch per frequency afterwards).  When the
  // fill is fed from host memory it is PCIe-bound and the GPU idles a third of the time: the 0.6 s of epsilon work at
  // C60 size then hide underneath the transfers.  Single rank, aux blocks filled in ascending order.  The raw
  // accumulations are handed to rpa_epsilon_dev when energies, eta, frequency and tensor contents match exactly.
  struct PpmPrefetch {
    bool armed = false, complete = false;
    std::vector<double> energies;
    long long n_occ = 0, filled_upto = 0, done_upto = 0;
    double eta = 0.0;
    DBuf E, d;                         // E: [2][naux][naux];  d: chi0 weights [2][n_occ][K]
    unsigned long long content_gen = 0;
    long long taken = 0;               // matrices handed to rpa_epsilon_dev so far (tests, reports)
  } ppm_pre;
  void ppm_prefetch_begin(const double* rpa_energies_host, long long n_occ, double eta);
  void ppm_prefetch_advance(long long P0, long long nP);
  bool ppm_prefetch_take(const std::vector<double>& energies, long long n_occ, double eta, double omega, bool imag,
                         double* out_dev);
  struct MetricHint { bool given = false; const double* V = nullptr; const double* S = nullptr; } hint;
  long long metric_cholesky_count = 0, metric_eig_count = 0;   // which path apply_coulomb_metric took (tests, reports)
  // dst[i][Q][j] = sum_P M[m0+i][P][n0+j] R[P,Q]   (window rotation into a caller-owned buffer)
  void rotate_window(double* dst, long long dst_ld, long long dst_slab, int m0, int mcnt, int n0, int ncnt,
                     const double* R_dev, long long ldr);
};

// eps(w) for n_omega frequencies on the device: out[w] (naux x naux, ld = naux).  energies_dev: rpatotal.
// owner_shift < 0: every rank receives every matrix (all-reduce).  owner_shift >= 0 (frequency sharding, world > 1):
// matrix w is summed onto rank (w + owner_shift) % world only -- the other ranks' copies hold partial sums and must not
// be used; the caller inverts / consumes matrix w on its owner (Sigma_CDA's quadrature nodes and residue poles).
// energies_host (optional): the same energies on the host, for the exact comparison with a PPM prefetch (see TCMatrix).
void rpa_epsilon_dev(TCMatrix& tc, const double* energies_dev, long long n_occ, double eta, const double* omegas_host,
                     int n_omega, bool imag, double gamma_extra, double* out_dev, int owner_shift = -1,
                     const std::vector<double>* energies_host = nullptr);

// E <- R^T E R (E symmetric, full storage in; lower triangle out on one rank, full matrix out when the product is
// split over the ranks); T: n x n scratch.  tc.cu.
void congruence_sym(Context* ctx, double* E, const double* R, double* T, long long n);

// ---------------------------------------------------------------- GW
struct GW {
  Context* ctx;
  TCMatrix* tc;
  xtpb_gw_options opt;
  long long qptotal, rpatotal, n_occ, q0;     // q0 = qpmin - rpamin (slab offset of gw level 0)
  long long n_occ_loc = 0;                    // occupied levels among this rank's second-index columns
  DBuf energies_loc_dev;                      // RPA energies of the local columns (multi-GPU only)
  const double* e_loc = nullptr;              // == energies_dev.p on a single GPU
  std::vector<double> vxc, dft_energies, rpa_energies, sigma_x, sigma_c;   // host copies (q x q col-major)
  DBuf energies_dev;
  bool screening_ready = false;
  long long unconverged = 0;
  // PPM
  std::vector<double> ppm_weight, ppm_freq;
  // last grid scan: 1 = compressed (far poles through Chebyshev moments), bins of its plan, pole evaluations it
  // performed one by one and the number the direct sum would have needed (this rank's share)
  int grid_compressed = 0;
  long long grid_bins = 0;
  double grid_direct_evals = 0.0, grid_equiv_evals = 0.0;
  DBuf ppm_freq_dev, ppm_fac_dev;
  // target-independent state of the last compressed grid scan (bin table, moments, equivalent poles of every QP
  // level): single (level, frequency) evaluations without derivatives go through it while it matches the tensor,
  // the energies and the PPM parameters (XTPB_SIGMA_POINTS=direct forces the slab-streaming pair kernel)
  PpmScanState scan;
  bool points_compressed(long long n, const long long* levels, const double* om0, double domega, int n_omega,
                         double* values);
  long long points_compressed_calls = 0, points_direct_calls = 0;
  // exact
  std::vector<double> rpa_omegas;
  DBuf residues;                // [level][s][m]  (m fastest, ld = tc->ldn)
  DBuf exact_omega_dev;
  long long rpasize = 0;
  // CDA
  std::vector<double> quad_points, quad_weights;
  DBuf cda_kernels;             // (order+1) matrices naux x naux: dielinv_j ..., kappa0 last
  DBuf cda_points_dev;          // points, then weights
  DBuf cda_q;                   // quadratic forms Q[level][j][m] (j = order: kappa0), ld = tc->ldn
  DBuf backup;                  // un-rotated tensor for evGW (stands in for TCMatrix_gwbse::Rebuild)

  GW(Context* c, TCMatrix* t, const xtpb_gw_options& o, const double* vxc_host, long long ldv, const double* e,
     long long ne);
  void set_rpa_energies(const double* e);
  void exchange(double* out_host);                 // unscaled Sigma_x
  void prepare_screening();
  void sigma_c_diag_elements(long long n, const long long* levels, const double* freqs, double* values, double* derivs);
  void sigma_c_offdiag(const double* freqs, double* out_host);
  void calculate_gw_perturbation();
  void calculate_hqp();
  std::vector<double> gwa_results() const;
  std::vector<double> hqp() const;
  std::vector<double> solve_qp(const std::vector<double>& frequencies);
  void grid_scan(const std::vector<double>& f0, std::vector<double>& values);   // values[level*steps + j]

 private:
  void prepare_ppm();
  void prepare_exact();
  void prepare_cda();
  void sigma_c_diag_elements_other(long long n, const long long* levels, const double* freqs, double* values,
                                   double* derivs);   // exact / CDA
  void sigma_c_offdiag_other(const double* freqs, double* out_host);
  void cda_values(long long n, const long long* levels, const double* freqs, double* values);
};
// GaussianQuadrature (gaussian_quadrature.cc): scaled points / weights on (0, inf); host code
void gaussian_quadrature(int scheme, long long order, std::vector<double>& points, std::vector<double>& weights);

// ---------------------------------------------------------------- operators + Davidson
struct Operator {
  Context* ctx;
  long long size = 0;
  virtual ~Operator() = default;
  virtual void matmul_dev(const double* X, long long ldx, int k, double* Y, long long ldy) = 0;
  virtual void diagonal_dev(double* d) = 0;
};

struct BSE {
  Context* ctx;
  TCMatrix* tc;
  xtpb_bse_options opt;
  long long vt, ct, size, naux;
  std::vector<double> hqp;          // (vt+ct)^2 host, col-major
  std::vector<double> eps_inv;      // host
  bool eps0_reused = false;         // eps(0) eigenvalues taken from the PPM rotation instead of being recomputed
  BSE(Context* c, TCMatrix* t, const xtpb_bse_options& o, const double* rpa_e, const double* hqp_in, long long ldh,
      bool rotate_full);
};

struct BseOperator : Operator {
  long long vt, ct, naux;
  int cqp, cx, cd, cd2;
  long long ldvc, slabvc, ldvv, slabvv, ldcc, slabcc, ldcv, slabcv;
  DBuf Mvc, Mvv, Mcc, Mcv;          // rotated windows; Mvv/Mcv carry eps_inv (scaled per P)
  DBuf Mvv_raw_diag, Mcc_diag;      // not used (diagonal kernel reads windows directly)
  DBuf eps_inv_dev, hqp_dev, hqp_diag_dev;   // hqp_dev: (vt+ct)^2 col-major
  DBuf T, U;
  // dense mode: the BSE Hamiltonian itself, built once from the windows (three long-K contractions) and kept in HBM;
  // matmul is then a single HBM-bound GEMM.  Rank r owns the columns (v2, c2) with v2 in [v2lo, v2lo + ns).
  bool dense = false;
  DBuf H;
  long long h_ld = 0, v2lo = 0, ns = 0;
  // dense mode keeps the exchange term factorised (rank N_aux): flat operand Fvc[P][(v,c)] (ld ldvcF, all P, all rows)
  bool hx_factorised = false;
  DBuf Fvc;
  long long ldvcF = 0, tc_naux_glob = 0;
  // windows are built from tc (optionally rotated by R_dev)
  BseOperator(Context* c, TCMatrix* tc, long long homo, long long rpamin, long long vmin, long long cmax,
              const double* eps_inv_host, const double* hqp_host, long long ldh, int cqp_, int cx_, int cd_, int cd2_,
              const double* R_dev, bool force_factorised = false);
  void matmul_dev(const double* X, long long ldx, int k, double* Y, long long ldy) override;
  void diagonal_dev(double* d) override;
};

struct DenseOperator : Operator {
  DBuf A;
  long long lda;
  DenseOperator(Context* c, const double* A_host, long long n, long long lda_host);
  void matmul_dev(const double* X, long long ldx, int k, double* Y, long long ldy) override;
  void diagonal_dev(double* d) override;
};

struct DavidsonResult {
  std::vector<double> evals;
  DBuf evecs;          // size x neigen
  int info = 1;
  long long iterations = 0;
};
void davidson_solve(Operator& A, long long neigen, const xtpb_davidson_options& opt, DavidsonResult& out);
struct BtdaResult {
  std::vector<double> evals;
  DBuf X, Y;           // size x neigen each, X^T X - Y^T Y = 1
  int info = 1;
  long long iterations = 0;
};
// full BSE [[A, B], [-B, -A]] (upstream BSE::Solve_nonhermitian_Davidson); A, B symmetric operators of equal size
void btda_solve(Operator& A, Operator& B, long long neigen, const xtpb_davidson_options& opt, BtdaResult& out);

}  // namespace xtpb

struct xtpb_ctx { xtpb::Context impl; explicit xtpb_ctx(int d) : impl(d) {} };
struct xtpb_tc { xtpb::TCMatrix impl; };
struct xtpb_gw { xtpb::GW impl; };
struct xtpb_bse { xtpb::BSE impl; std::unique_ptr<xtpb::DBuf> R; std::vector<double> rpa_energies; };
struct xtpb_op { std::unique_ptr<xtpb::Operator> impl; };
