/* libxtpb200 -- C ABI of the B200-native GW-BSE tensor-contraction path of VOTCA-XTP.
 *
 * Drop-in boundary.  The reference has no C ABI for this path:
 * its boundary is a set of C++/Eigen classes called from GWBSE::Evaluate
 * (upstream xtp/src/libxtp/gwbse/gwbse.cc).  Each entry point below names the
 * reference member function it replaces; upstream paths are relative to
 * votca/votca and carry NO line numbers because the mounted reference
 * (/root/reference/README.md:1) is a one-line redirect stub -- see SURVEY.md section 0.
 * The header-only C++ facade in xtp_facade.hpp re-creates the reference's class
 * names on top of these functions; INTEGRATION.md shows the binding a
 * maintainer would add on the XTP side.
 *
 * Conventions
 *   - every function returns 0 on success; otherwise xtpb_last_error() holds a
 *     thread-local message (the reference throws std::runtime_error instead).
 *   - all matrices are FP64 column-major (Eigen::MatrixXd default) with an
 *     explicit leading dimension where one is taken; indices are 64-bit signed
 *     (Eigen::Index) in the option structs, absolute DFT level indices.
 *   - pointers named *_host are caller-owned host memory; the library owns all
 *     device memory.  Handles are not thread-safe; one host thread per context.
 *   - no CPU fallback: without a CUDA device xtpb_ctx_create fails.
 */
#ifndef XTPB200_H
#define XTPB200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef long long xtpb_index;
typedef struct xtpb_ctx xtpb_ctx;
typedef struct xtpb_tc xtpb_tc;
typedef struct xtpb_gw xtpb_gw;
typedef struct xtpb_bse xtpb_bse;
typedef struct xtpb_op xtpb_op;

const char* xtpb_last_error(void);
int xtpb_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py "gpu_launches") */
long long xtpb_launch_count(void);
/* of which launches of the contraction engine's TMA-fed instance (cp.async.bulk.tensor + mbarrier pipeline) */
long long xtpb_tma_launch_count(void);
/* ... of which launches that fetch a row-contiguous operand tile with ONE 5-D tensor-map box (contract.cu) */
long long xtpb_tma_single_box_launch_count(void);

/* ---- context: one CUDA device + stream + scratch.  Replaces OpenMP_CUDA / CudaPipeline
 *      (upstream xtp/src/libxtp/openmp_cuda.cc, cudapipeline.cc). ---- */
int xtpb_ctx_create(int device, xtpb_ctx** out);
int xtpb_ctx_destroy(xtpb_ctx* ctx);
int xtpb_ctx_sync(xtpb_ctx* ctx);
/* seconds spent inside cuSOLVER eigh/inverse since the last reset (reported apart from contractions) */
int xtpb_ctx_solver_seconds(xtpb_ctx* ctx, double* seconds, int reset);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink.  The reference has no distributed layer on this path (one
 *      process, an OpenMP thread per GPU, host reductions: upstream xtp/src/libxtp/openmp_cuda.cc); here rank 0 obtains
 *      a 128-byte id, the host program distributes it (torch.distributed, MPI, a file), and every rank joins BEFORE
 *      creating its TCMatrix.  From then on every rank makes the same sequence of calls with the same host inputs
 *      (the calls are collective); the tensor is distributed over its second index, the BSE operator over the
 *      auxiliary index, and all results returned to the host are complete and identical on every rank. ---- */
int xtpb_comm_unique_id(char* id_128);
int xtpb_ctx_comm_init(xtpb_ctx* ctx, const char* id_128, int rank, int world);
int xtpb_ctx_comm_info(xtpb_ctx* ctx, int* rank, int* world);

/* Host seconds this process has spent inside cudaMalloc / cudaFree on behalf of the library, the number of device
 * allocations, and -- with XTPB_ALLOC_CACHE=1 in the environment when the library is loaded: released blocks are kept
 * in an exact-size cache instead of being returned to the driver -- cache hits and bytes currently cached.
 * Any pointer may be NULL; reset != 0 zeroes the counters. */
int xtpb_alloc_stats(double* seconds, long long* calls, long long* cache_hits, double* cached_bytes, int reset);
/* Host seconds spent, since the last reset of xtpb_alloc_stats, waiting for outstanding GPU work before a released block
 * went into the cache (the device-wide synchronisation cudaFree would have implied): GPU time, not allocator time. */
int xtpb_alloc_free_wait_seconds(double* seconds);

/* pinned (page-locked) host memory for the caller's AO-integral and result buffers: H2D/D2H copies from it run
 * asynchronously at PCIe rate and overlap the contractions (xtpb_tc_fill_block*). */
int xtpb_host_alloc(unsigned long long bytes, void** out);
int xtpb_host_free(void* p);

/* ---- kernel timing for roofline reports: when enabled, every launch of the contraction engine and of the fused
 *      Sigma_c kernels is bracketed by a CUDA-event pair on the launching stream.  Tags (xtpb_profile_get):
 *      0 other, 1 Fill3cMO, 2 aux rotation, 3 epsilon, 4 Sigma_x, 5 Sigma_c off-diagonal, 6 BSE matmul,
 *      7 Davidson projections, 8 small dense (PPM), 9 Sigma_c PPM grid kernel (work = pole evaluations),
 *      10 Sigma_c PPM pair kernel (work = bytes), 11 cuSOLVER, 12 AO unpack (work = bytes), 13 CDA, 14 exact,
 *      15 NCCL collectives (work = payload bytes).
 *      `work` is the summed algorithmic flop count (2MNK; MNK for lower-triangular outputs) unless noted. ---- */
int xtpb_profile_enable(int on);
int xtpb_profile_reset(void);
int xtpb_profile_get(int tag, double* ms, double* work, long long* launches);

/* ---- TCMatrix_gwbse (upstream xtp/include/votca/xtp/threecenter.h, xtp/src/libxtp/threecenter_gwbse.cc) ---- */
/* TCMatrix_gwbse::Initialize(basissize, mmin, mmax, nmin, nmax) */
int xtpb_tc_create(xtpb_ctx* ctx, xtpb_index auxsize, xtpb_index mmin, xtpb_index mmax, xtpb_index nmin,
                   xtpb_index nmax, xtpb_tc** out);
int xtpb_tc_destroy(xtpb_tc* tc);
/* accessors auxsize()/msize()/nsize() */
int xtpb_tc_sizes(const xtpb_tc* tc, xtpb_index* auxsize, xtpb_index* msize, xtpb_index* nsize);
/* whole tensor in the reference's host layout: mtotal slabs, each ntotal x auxsize column-major */
int xtpb_tc_set_raw(xtpb_tc* tc, const double* M_host);
/* the same from a device pointer (same layout, contiguous): synthetic tensors drawn on the GPU (BASELINE configs[4]) */
int xtpb_tc_set_raw_dev(xtpb_tc* tc, const double* M_dev);
/* Read-only view of the resident tensor for callers that keep working on the device (and for parity checks at sizes
 * whose tensor does not fit host memory comfortably): element M[m](n, P) of the reference's matrix_[m] lives at
 * M_dev[m*slab_stride + P*ld_n + n], n < n_local (this rank's share of the second index; all of it on one GPU).
 * Pending aux rotations are applied first.  The pointer is invalidated by xtpb_tc_destroy only. */
int xtpb_tc_device_view(xtpb_tc* tc, const double** M_dev, xtpb_index* ld_n, xtpb_index* slab_stride,
                        xtpb_index* n_local);
/* TCMatrix_gwbse::operator[](m): slab m (ntotal x auxsize, column-major, ld = ntotal) */
int xtpb_tc_get_slab(xtpb_tc* tc, xtpb_index m, double* slab_host);
/* TCMatrix_gwbse::Fill3cMO, split so the caller's integral loop can stream aux shells:
 *   fill_begin : dft_orbitals (n_basis x >= max(mmax,nmax)+1, column-major, ld = ldc)
 *   fill_block : nP consecutive aux functions starting at P0; ao3c holds nP symmetric
 *                n_basis x n_basis slices (column-major, ld = ld_ao, slice stride = ld_ao*n_basis).
 *                Replaces OpenMP_CUDA::setOperators + MultiplyLeftRight.
 *   the _dev variant takes a device pointer (inputs already resident in HBM). */
int xtpb_tc_fill_begin(xtpb_tc* tc, xtpb_index n_basis, const double* C_host, xtpb_index ldc);
int xtpb_tc_fill_block(xtpb_tc* tc, xtpb_index P0, xtpb_index nP, const double* ao3c_host, xtpb_index ld_ao);
int xtpb_tc_fill_block_dev(xtpb_tc* tc, xtpb_index P0, xtpb_index nP, const double* ao3c_dev, xtpb_index ld_ao);
/* the same with packed lower triangles (row mu holds nu = 0..mu; slice stride n_basis(n_basis+1)/2): half the
 * PCIe / HBM bytes of the full symmetric slices */
int xtpb_tc_fill_block_packed(xtpb_tc* tc, xtpb_index P0, xtpb_index nP, const double* ao3c_packed_host);
int xtpb_tc_fill_block_packed_dev(xtpb_tc* tc, xtpb_index P0, xtpb_index nP, const double* ao3c_packed_dev);
/* collective Fill3cMO for world > 1 (also valid for world == 1): this rank passes the packed slices of ITS aux range
 * [P0, P0+nP) = xtpb_tc_local_aux_range (naux*rank/world .. naux*(rank+1)/world), host (pinned) or device pointer.
 * Half-transformed blocks are all-gathered over NVLink while the next block is being contracted. */
int xtpb_tc_local_aux_range(xtpb_tc* tc, xtpb_index* P0, xtpb_index* nP);
int xtpb_tc_fill_sharded_packed(xtpb_tc* tc, const double* ao3c_packed_local, int on_device);
/* TCMatrix_gwbse::MultiplyRightWithAuxMatrix(matrix): matrix is auxsize x auxsize */
int xtpb_tc_multiply_right_with_aux_matrix(xtpb_tc* tc, const double* A_host, xtpb_index lda);
/* second half of TCMatrix_gwbse::Fill: AOCoulomb::Pseudo_InvSqrt_GWBSE(auxoverlap, etol) followed by
 * MultiplyRightWithAuxMatrix.  S_host may be NULL (orthonormal aux basis). */
int xtpb_tc_apply_coulomb_metric(xtpb_tc* tc, const double* V_host, xtpb_index ldv, const double* S_host,
                                 xtpb_index lds, double etol, xtpb_index* removed_functions);
/* Optional hint, given BEFORE the xtpb_tc_fill_* calls: the matrices xtpb_tc_apply_coulomb_metric will receive (the two
 * halves of TCMatrix_gwbse::Fill have independent inputs).  The same pointers must then be passed to
 * xtpb_tc_apply_coulomb_metric; the host matrices must stay valid until it returns.  With XTPB_METRIC_CHOLESKY=0 (or
 * XTPB_LAZY_METRIC=0) the first eigendecomposition of the metric step is started on a helper thread / stream at once,
 * underneath the fill; otherwise nothing is started (see xtpb_tc_metric_path_info). */
int xtpb_tc_coulomb_metric_begin(xtpb_tc* tc, const double* V_host, xtpb_index ldv, const double* S_host,
                                 xtpb_index lds);
/* Optional hint, given AFTER xtpb_tc_fill_begin and BEFORE the xtpb_tc_fill_block* calls of a single-rank fill: the
 * caller is going to run Sigma_PPM::PrepareScreening (xtpb_gw_prepare_screening / xtpb_gw_calculate_gw_perturbation with
 * the plasmon-pole model) with these RPA input energies (rpamin .. rpamax, as handed to the GW object after the scissor
 * shift), this HOMO index and this eta.  The two epsilon matrices of the plasmon-pole model (w = 0 on the real axis,
 * w = 0.5 Ha on the imaginary axis) are then accumulated panel by panel while the aux blocks arrive (ascending,
 * contiguous from 0), i.e. underneath the host-to-device transfers of a PCIe-bound fill; PrepareScreening picks the
 * result up when energies, eta and tensor contents match exactly and recomputes otherwise.  Ignored with several
 * ranks.  (The reference's GWBSE::Evaluate knows these energies before it calls TCMatrix_gwbse::Fill.) */
int xtpb_tc_ppm_prefetch_begin(xtpb_tc* tc, const double* rpa_energies_host, xtpb_index homo, double eta);
/* complete = 1 once both matrices are accumulated for all aux functions; aux_functions_done: rows finished so far;
 * matrices_used: how many prefetched matrices PrepareScreening has picked up on this tensor.  Pointers may be NULL. */
int xtpb_tc_ppm_prefetch_info(xtpb_tc* tc, int* complete, xtpb_index* aux_functions_done, xtpb_index* matrices_used);
/* How xtpb_tc_apply_coulomb_metric obtained its factor R (R R^T = V^-1) so far on this tensor.  eigensolver: the
 * reference's construction, S^-1/2 (S^-1/2 V S^-1/2)^-1/2 with eigenvalues below etol dropped.  cholesky: when two
 * Cholesky factorisations prove that S - etol and V - etol S are positive definite (no function would be removed),
 * R = U^-1 from V = U^T U: epsilon, the plasmon-pole-rotated tensor, Sigma and the BSE are the same for every such R,
 * and if the metric-rotated tensor itself is read before the next rotation the symmetric factor is built then (counted
 * as an eigensolver call).  Either pointer may be NULL. */
int xtpb_tc_metric_path_info(xtpb_tc* tc, xtpb_index* cholesky_calls, xtpb_index* eigensolver_calls);

/* ---- RPA (upstream xtp/src/libxtp/gwbse/rpa.cc) ---- */
/* RPA::calculate_epsilon_i / calculate_epsilon_r for n_omega frequencies in one call.
 * energies_host: RPA input energies, length rpamax-rpamin+1.  eps_host: n_omega matrices auxsize x auxsize. */
int xtpb_rpa_epsilon(xtpb_tc* tc, const double* energies_host, xtpb_index homo, xtpb_index rpamin, xtpb_index rpamax,
                     double eta, const double* omegas_host, int n_omega, int imaginary_axis, double* eps_host);

/* ---- Sigma_base / Sigma_PPM / Sigma_Exact / Sigma_CDA / GW
 *      (upstream xtp/src/libxtp/gwbse/sigma_base.cc, sigma_ppm.cc, sigma_exact.cc, sigma_cda.cc, ppm.cc, gw.cc) ---- */
enum { XTPB_SIGMA_PPM = 0, XTPB_SIGMA_EXACT = 1, XTPB_SIGMA_CDA = 2 };
enum { XTPB_QP_GRID = 0, XTPB_QP_FIXEDPOINT = 1 };
enum { XTPB_QUAD_LEGENDRE = 0, XTPB_QUAD_LAGUERRE = 1, XTPB_QUAD_HERMITE = 2 };

typedef struct xtpb_gw_options {   /* GW::options + Sigma_base::options, upstream gw.h / sigma_base.h */
  xtpb_index homo, qpmin, qpmax, rpamin, rpamax;
  double eta;                 /* 1e-3 */
  double g_sc_limit;          /* 1e-5  QP equation convergence [Ha] */
  xtpb_index g_sc_max_iterations; /* 100 */
  double gw_sc_limit;         /* 1e-5  evGW convergence */
  xtpb_index gw_sc_max_iterations; /* 1 = G0W0 */
  double shift;               /* scissor shift */
  double ScaHFX;
  int sigma_integration;      /* XTPB_SIGMA_* */
  xtpb_index reset_3c;        /* 5 */
  int qp_solver;              /* XTPB_QP_* */
  xtpb_index qp_grid_steps;   /* 1001 */
  double qp_grid_spacing;     /* 0.01 Ha */
  xtpb_index gw_mixing_order; /* 0 = plain */
  double gw_mixing_alpha;     /* 0.7 */
  int quadrature_scheme;      /* XTPB_QUAD_* (CDA) */
  xtpb_index order;           /* 12 (CDA) */
  double alpha;               /* 1e-3 (CDA tail) */
} xtpb_gw_options;
void xtpb_gw_options_default(xtpb_gw_options* opt);

/* GaussianQuadrature::configure + ScaledPoint/ScaledWeight (upstream gwbse/gaussian_quadrature.cc): the
 * imaginary-axis nodes of Sigma_CDA mapped to (0, inf).  points/weights need room for `order` values (legendre,
 * laguerre) or `order` positive nodes of a 2*order rule (hermite).  Host-only, needs no device. */
int xtpb_gaussian_quadrature(int scheme, xtpb_index order, double* points, double* weights, xtpb_index* count);
/* GW::GW(log, Mmn, vxc, dft_energies) + GW::configure(opt).  vxc is qptotal x qptotal. */
int xtpb_gw_create(xtpb_ctx* ctx, xtpb_tc* tc, const xtpb_gw_options* opt, const double* vxc_host, xtpb_index ldv,
                   const double* dft_energies_host, xtpb_index n_energies, xtpb_gw** out);
int xtpb_gw_destroy(xtpb_gw* gw);
/* Sigma_base::CalcExchangeMatrix (qptotal x qptotal, unscaled by ScaHFX) */
int xtpb_gw_sigma_exchange(xtpb_gw* gw, double* sigma_x_host);
/* RPA::setRPAInputEnergies / getRPAInputEnergies (length rpatotal) */
int xtpb_gw_set_rpa_input_energies(xtpb_gw* gw, const double* e_host);
int xtpb_gw_get_rpa_input_energies(xtpb_gw* gw, double* e_host);
/* Sigma_*::PrepareScreening() */
int xtpb_gw_prepare_screening(xtpb_gw* gw);
/* PPM::getPpm_weight / getPpm_freq (length auxsize) after PrepareScreening with XTPB_SIGMA_PPM */
int xtpb_gw_get_ppm(xtpb_gw* gw, double* weight_host, double* freq_host);
/* Sigma_*::CalcCorrelationDiagElement and ...DiagElementDerivative for n (level, frequency) pairs in one
 * call; levels are gw-level indices (0 = qpmin).  derivs_host may be NULL. */
int xtpb_gw_sigma_c_diag_elements(xtpb_gw* gw, xtpb_index n, const xtpb_index* levels_host,
                                  const double* frequencies_host, double* values_host, double* derivs_host);
/* Sigma_base::CalcCorrelationDiag(frequencies): one frequency per gw level */
int xtpb_gw_sigma_c_diag(xtpb_gw* gw, const double* frequencies_host, double* values_host);
/* Sigma_c of every gw level on its QP grid (the scan GW::SolveQP_Grid runs, and what GW::PlotSigma tabulates):
 * values_host[level*qp_grid_steps + j] = Sigma_c(level, center[level] + (j - (steps-1)/2) * qp_grid_spacing) */
int xtpb_gw_sigma_c_grid(xtpb_gw* gw, const double* center_frequencies_host, double* values_host);
/* How the last grid scan (xtpb_gw_sigma_c_grid, or the one inside GW::SolveQP_Grid) was evaluated on this rank:
 * compressed = 1 when far poles went through Chebyshev moments (0: pole-by-pole kernel, e.g. XTPB_SIGMA_GRID=direct,
 * unsorted RPA energies or a non-PPM Sigma), the bins of its plan, the (pole, frequency) evaluations actually
 * performed one by one and the number the plain double sum needs.  Any pointer may be NULL. */
int xtpb_gw_grid_scan_info(xtpb_gw* gw, int* compressed, xtpb_index* n_bins, double* direct_evaluations,
                           double* equivalent_evaluations);
/* How single (level, frequency) values of the plasmon-pole Sigma_c without derivatives (the bisection rounds and the
 * final Sigma_c of GW::SolveQP, xtpb_gw_sigma_c_diag_elements with derivatives == NULL) have been evaluated since the
 * handle was created: calls served from the moments of the last compressed grid scan (no slab traffic; valid while
 * the tensor, the RPA energies and the PPM parameters are unchanged) and calls that streamed the slabs
 * (XTPB_SIGMA_POINTS=direct forces the latter).  Either pointer may be NULL. */
int xtpb_gw_point_eval_info(xtpb_gw* gw, xtpb_index* compressed_calls, xtpb_index* direct_calls);
/* Host-side plan of the compressed grid scan behind xtpb_gw_sigma_c_grid / GW::SolveQP_Grid (needs no device; exposed
 * so that the bin/near-range logic can be tested on a CPU box).  The pole axis [zmin, zmax] is cut into n_bins bins
 * (edges: n_bins + 1 ascending values); near_ranges[(level*n_chunks + chunk)*4 + {0,1}] is the inclusive range of bins
 * whose poles are summed one by one for the chunk of C = xtpb_ppm_grid_chunk() consecutive grid points starting at
 * grid_start[level] + C*chunk*spacing (lo > hi: none); every other bin enters through its Chebyshev moments.
 * Entries {2,3} are the inclusive sub-range of INNER bins, all of whose poles lie inside the damping window of every
 * grid point of the chunk: there the damped kernel is smooth in the pole position and the bin is replaced by 16
 * equivalent poles at its Chebyshev nodes (none: entry 2 = entry 1 + 1, entry 3 = entry 1).
 * usable = 0: the scan falls back to the pole-by-pole kernel.  edges / near_ranges may be NULL to query the sizes
 * (near_ranges needs 4*n_levels*n_chunks ints, n_chunks = ceil(steps/C)). */
int xtpb_ppm_grid_chunk(void);      /* consecutive grid points per chunk of the plan below */
int xtpb_ppm_grid_plan(xtpb_index n_levels, const double* grid_start, double spacing, xtpb_index steps, double zmin,
                       double zmax, xtpb_index edges_capacity, double* edges, xtpb_index* n_bins, int* near_ranges,
                       xtpb_index* n_chunks, int* usable);
/* GW::PlotSigma(filename, steps, spacing, states): for every listed gw level (0 = qpmin) the correlation self-energy
 * on `steps` frequencies centred on the level's RPA input energy, as the steps x (2 n_states) column-major table
 * upstream writes to the file: column 2i = frequency, column 2i+1 = Sigma_c(w) + e_KS + Sigma_x - Vxc (the curve
 * whose intersection with w is the quasiparticle energy).  Needs PrepareScreening and Sigma_x (CalculateGWPerturbation). */
int xtpb_gw_plot_sigma(xtpb_gw* gw, xtpb_index steps, double spacing, xtpb_index n_states,
                       const xtpb_index* states_host, double* table_host);
/* Sigma_base::CalcCorrelationOffDiag(frequencies): qptotal x qptotal, zero diagonal */
int xtpb_gw_sigma_c_offdiag(xtpb_gw* gw, const double* frequencies_host, double* sigma_c_host);
/* GW::CalculateGWPerturbation / CalculateHQP / getGWAResults / getHQP / DiagonalizeQPHamiltonian */
int xtpb_gw_calculate_gw_perturbation(xtpb_gw* gw);
int xtpb_gw_calculate_hqp(xtpb_gw* gw);
int xtpb_gw_get_gwa_results(xtpb_gw* gw, double* qp_energies_host);
int xtpb_gw_get_hqp(xtpb_gw* gw, double* hqp_host);
int xtpb_gw_diagonalize_qp_hamiltonian(xtpb_gw* gw, double* eigenvalues_host, double* eigenvectors_host);
/* number of gw levels whose QP equation did not converge in the last SolveQP */
int xtpb_gw_unconverged_levels(xtpb_gw* gw, xtpb_index* count);

/* ---- BSE / BSE_OPERATOR (upstream xtp/src/libxtp/gwbse/bse.cc, bse_operator.{h,cc}) ---- */
typedef struct xtpb_bse_options {  /* BSE::options, upstream bse.h */
  xtpb_index homo, rpamin, rpamax, qpmin, qpmax, vmin, cmax;
  xtpb_index nmax;
  int use_Hqp_offdiag;        /* 1 */
} xtpb_bse_options;

/* BSE::configure(opt, RPAInputEnergies, Hqp_in): AdjustHqpSize + SetupDirectInteractionOperator
 * (epsilon(0) at the given energies, eigen-decomposition, aux rotation of the BSE windows).
 * The rotated windows are kept in BSE-owned device buffers; unlike the reference the
 * TCMatrix itself is left untouched unless rotate_full_tc != 0. */
int xtpb_bse_create(xtpb_ctx* ctx, xtpb_tc* tc, const xtpb_bse_options* opt, const double* rpa_input_energies_host,
                    const double* hqp_host, xtpb_index ldh, int rotate_full_tc, xtpb_bse** out);
int xtpb_bse_destroy(xtpb_bse* bse);
int xtpb_bse_get_epsilon_0_inv(xtpb_bse* bse, double* eps_inv_host);
/* eps0_reused = 1 when SetupDirectInteractionOperator found the tensor already in the eigenbasis of epsilon(0) for
 * exactly these RPA input energies (G0W0 after Sigma_PPM::PrepareScreening rotated it there) and read the eigenvalues
 * instead of recomputing epsilon(0), its eigensolver and the rotation.  XTPB_BSE_REUSE_EPS0=0 disables the shortcut. */
int xtpb_bse_screening_info(xtpb_bse* bse, int* eps0_reused);
/* BSE_OPERATOR<cqp,cx,cd,cd2>(epsilon_0_inv, Mmn, Hqp) + configure(BSEOperator_Options).
 * Typedefs upstream: SingletOperator_TDA <1,2,1,0>, TripletOperator_TDA <1,0,1,0>, SingletOperator_BTDA_B <0,2,0,1>,
 * TripletOperator_BTDA_B <0,0,0,1>, HxOperator <0,1,0,0>, HdOperator <0,0,1,0>, Hd2Operator <0,0,0,1>, HqpOperator <1,0,0,0>. */
int xtpb_bse_operator_create(xtpb_bse* bse, int cqp, int cx, int cd, int cd2, xtpb_op** out);
/* same operator built from explicit pieces, without an epsilon(0) rotation: eps_inv (auxsize),
 * Hqp ((vtotal+ctotal)^2).  Used by operator-level tests, mirrors the reference constructor. */
int xtpb_bse_operator_create_raw(xtpb_ctx* ctx, xtpb_tc* tc, xtpb_index homo, xtpb_index rpamin, xtpb_index vmin,
                                 xtpb_index cmax, const double* eps_inv_host, const double* hqp_host, xtpb_index ldh,
                                 int cqp, int cx, int cd, int cd2, xtpb_op** out);
/* a dense symmetric matrix as a MatrixFreeOperator (tests of the Davidson solver) */
int xtpb_dense_operator_create(xtpb_ctx* ctx, const double* A_host, xtpb_index n, xtpb_index lda, xtpb_op** out);
int xtpb_op_destroy(xtpb_op* op);
/* MatrixFreeOperator::rows() */
int xtpb_op_size(xtpb_op* op, xtpb_index* size);
/* MatrixFreeOperator::matmul(X): X, Y are size x k */
int xtpb_op_matmul(xtpb_op* op, const double* X_host, xtpb_index ldx, xtpb_index k, double* Y_host, xtpb_index ldy);
/* MatrixFreeOperator::diagonal() */
int xtpb_op_diagonal(xtpb_op* op, double* diag_host);
/* MatrixFreeOperator::get_full_matrix() (size x size; small problems only) */
int xtpb_op_get_full_matrix(xtpb_op* op, double* H_host, xtpb_index ldh);

/* ---- DavidsonSolver (upstream xtp/src/libxtp/davidsonsolver.cc) ---- */
enum { XTPB_DAVIDSON_DPR = 0, XTPB_DAVIDSON_OLSEN = 1 };
enum { XTPB_UPDATE_MIN = 0, XTPB_UPDATE_SAFE = 1, XTPB_UPDATE_MAX = 2 };
typedef struct xtpb_davidson_options {
  double tolerance;           /* loose 1e-3, normal 1e-4, strict 1e-5, lapack 1e-9 */
  int correction;             /* XTPB_DAVIDSON_* */
  int size_update;            /* XTPB_UPDATE_* */
  xtpb_index iter_max;        /* 50 */
  xtpb_index max_search_space;/* 0 -> 5*neigen; BSE uses 10*nmax */
  xtpb_index size_initial_guess; /* 0 -> 2*neigen */
} xtpb_davidson_options;
void xtpb_davidson_options_default(xtpb_davidson_options* opt);
/* DavidsonSolver::solve(A, neigen, size_initial_guess) for symmetric operators.
 * eigenvalues_host: neigen; eigenvectors_host: size x neigen (ld = ldv).
 * info: 0 = Eigen::Success, 1 = Eigen::NoConvergence.  iterations: num_iterations(). */
int xtpb_davidson_solve(xtpb_op* op, xtpb_index neigen, const xtpb_davidson_options* opt, double* eigenvalues_host,
                        double* eigenvectors_host, xtpb_index ldv, int* info, xtpb_index* iterations);

/* Anderson::UpdateInput / UpdateOutput / MixHistory (upstream xtp/src/libxtp/anderson_mixing.cc), the mixer
 * GW::CalculateGWPerturbation applies to the evGW iterates when gw_mixing_order > 0 (1 = linear mixing): feeds
 * n_history (input, output) pairs of length n, oldest first (row h at inputs_host + h*n), through a mixer of the
 * given order and returns the next guess.  Host-only, needs no device. */
int xtpb_anderson_mix(xtpb_index order, double alpha, xtpb_index n, xtpb_index n_history, const double* inputs_host,
                      const double* outputs_host, double* mixed_host);
/* The dense symmetric eigensolver the Davidson solvers apply to their projected matrices (upstream:
 * Eigen::SelfAdjointEigenSolver inside DavidsonSolver::getRitz, davidsonsolver.cc): Householder tridiagonalisation +
 * implicit QL on the host.  A_host (n x n, ld = lda, lower triangle read) is overwritten by the eigenvectors,
 * w_host receives the ascending eigenvalues.  Host-only, needs no device. */
int xtpb_host_eigh(xtpb_index n, double* A_host, xtpb_index lda, double* w_host);

/* ---- full BSE, transition dipoles, oscillator strengths ---- */
/* Full (non-TDA) BSE: BSE::Solve_singlets / Solve_triplets with useTDA = false (upstream bse.cc
 * Solve_nonhermitian_Davidson over HamiltonianOperator<A,B>, A = *_TDA operator, B = *_BTDA_B operator).
 * energies: nmax; X, Y: size x nmax (ld), normalised X^T X - Y^T Y = 1 (upstream BSE_*_coefficients / _AR).
 * xtpb_davidson_options as for the TDA solver (tolerance on both residuals, iter_max, max_search_space). */
int xtpb_bse_solve_btda(xtpb_bse* bse, int singlet, const xtpb_davidson_options* opt, double* energies_host,
                        double* X_host, double* Y_host, xtpb_index ld, int* info, xtpb_index* iterations);
/* BSE::Perturbative_DynamicalScreening(type, orb): first-order correction of BSE energies for the frequency
 * dependence of the screening in the direct term.  Per state s: E <- E_static + <s|Hd^(w = E)|s> - <s|Hd^(0)|s> with
 * Hd^ = HdOperator built from epsilon(w) on the real axis (SetupDirectInteractionOperator at w), repeated until
 * |dE| < dyn_tolerance or max_dyn_iter (upstream defaults 10 and 1e-5 Ha).  X (and Y for the full BSE, else NULL):
 * size x n_states eigenvectors as returned by the solvers.  iterations_host may be NULL. */
int xtpb_bse_perturbative_dynamical_screening(xtpb_bse* bse, xtpb_index n_states, const double* energies_static_host,
                                              const double* X_host, const double* Y_host, xtpb_index ld,
                                              xtpb_index max_dyn_iter, double dyn_tolerance,
                                              double* energies_dynamic_host, xtpb_index* iterations_host);
/* BSE::CalcCoupledTransition_Dipoles (+ Orbitals::CalcFreeTransition_Dips): d_s = -sqrt(2) sum_vc (X+Y)_vc,s <v|r|c>
 * from the three AO dipole matrices (3 x n_basis x n_basis, symmetric) and the MO coefficients.  Y may be NULL (TDA).
 * dipoles_host: n_states x 3, state-major. */
int xtpb_bse_transition_dipoles(xtpb_bse* bse, xtpb_index n_basis, const double* C_host, xtpb_index ldc,
                                const double* ao_dipoles_host, xtpb_index n_states, const double* X_host,
                                const double* Y_host, xtpb_index ld, double* dipoles_host);
/* Orbitals::Oscillatorstrengths: f_s = 2/3 E_s |d_s|^2 (host arithmetic, no device needed) */
int xtpb_oscillator_strengths(xtpb_index n_states, const double* energies_host, const double* dipoles_host,
                              double* strengths_host);

/* ---- GWBSE::Initialize level-range logic (upstream xtp/src/libxtp/gwbse/gwbse.cc) ----
 * The `ranges` option of the dftgwbse calculator: which levels enter the RPA, the QP equation and the BSE.
 *   default : rpamax = n_levels-1, qpmin = 0, qpmax = 2 homo + 1, bse vmin = 0, cmax = 2 homo + 1
 *   factor  : rpamax = int(f_rpamax n_levels) - 1;  qpmin = n_occ - int(f_qpmin n_occ) - 1;
 *             qpmax = n_occ + int(f_qpmax n_occ) - 1;  vmin, cmax likewise from f_bsemin, f_bsemax
 *   explicit: the five values are level indices
 *   full    : everything up to n_levels-1
 * followed by the clamps (upper bounds to n_levels-1, lower bounds to 0 and to at most homo / at least homo+1) and
 * rpamin = n_core_ignored (the `ignore_corelevels` option; 0 = keep all).  Host arithmetic, needs no device. */
enum { XTPB_RANGES_DEFAULT = 0, XTPB_RANGES_FACTOR = 1, XTPB_RANGES_EXPLICIT = 2, XTPB_RANGES_FULL = 3 };
typedef struct xtpb_gwbse_range_options {
  int mode;                         /* XTPB_RANGES_* */
  xtpb_index n_levels;              /* Orbitals::getBasisSetSize() */
  xtpb_index n_occ;                 /* Orbitals::getNumberOfAlphaElectrons(); homo = n_occ - 1 */
  xtpb_index n_core_ignored;        /* rpamin */
  double rpamax, qpmin, qpmax, bsemin, bsemax;   /* factors (factor mode) or level indices (explicit mode) */
} xtpb_gwbse_range_options;
typedef struct xtpb_gwbse_ranges {
  xtpb_index homo, rpamin, rpamax, qpmin, qpmax, vmin, cmax;
  xtpb_index qptotal, rpatotal, bse_vtotal, bse_ctotal, bse_size;
} xtpb_gwbse_ranges;
int xtpb_gwbse_level_ranges(const xtpb_gwbse_range_options* opt, xtpb_gwbse_ranges* out);

/* ---- engine-level hook used by the parity tests of the contraction kernel ----
 * C = alpha * sum_{outer,k} A(row,outer,k) d(outer,k) B(col,outer,k) + beta*C with every stride explicit
 * (element units, host buffers of the given lengths are copied to the device and back). */
typedef struct xtpb_contract_desc {
  xtpb_index M, N, K, n_outer, n_batch;
  xtpb_index a_row, a_k, a_outer, a_batch, a_len;
  xtpb_index b_row, b_k, b_outer, b_batch, b_len;
  xtpb_index c_row, c_col, c_batch, c_len, c_col_inner, c_col_outer;
  xtpb_index d_outer, d_batch, d_len;   /* d_len = 0: no weights */
  double alpha, beta;
  int lower, force_cfg, force_splits;   /* force_cfg: -1 auto; 0..2 tile; 4..7 cp.async instance; 8..10 TMA instance */
} xtpb_contract_desc;
int xtpb_contract_host(xtpb_ctx* ctx, const xtpb_contract_desc* desc, const double* A_host, const double* B_host,
                       const double* d_host, double* C_host);
/* times `reps` launches of the same contraction on device buffers filled with pseudo-random data; returns
 * the mean milliseconds per launch (CUDA events on the library stream). */
int xtpb_contract_bench(xtpb_ctx* ctx, const xtpb_contract_desc* desc, int reps, double* ms_per_launch);
/* The launcher's plan for a shape on a device with n_sms SMs -- host arithmetic only, needs no device: tile_cfg
 * (0: 128x128, 1: 128x64, 2: 128x32 CTA tiles) and the split-K factor (1 = none), honouring desc->force_cfg /
 * force_splits.  Exposed so that the heuristics (small-grid split, tail-balancing split of long contractions) can be
 * tested on a CPU box. */
int xtpb_contract_plan(const xtpb_contract_desc* desc, int n_sms, int* tile_cfg, int* split_k);

#ifdef __cplusplus
}
#endif
#endif /* XTPB200_H */
