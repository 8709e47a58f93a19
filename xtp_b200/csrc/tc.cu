// TCMatrix_gwbse on the device: the RI tensor M[m][P][n] stays resident in HBM for the whole GW-BSE step.
// Upstream: xtp/src/libxtp/threecenter_gwbse.cc (Initialize, Fill3cMO, MultiplyRightWithAuxMatrix, operator[]).
#include <algorithm>

#include "internal.h"

namespace xtpb {

TCMatrix::TCMatrix(Context* c, long long auxsize, long long mmin_, long long mmax_, long long nmin_, long long nmax_)
    : ctx(c), naux(auxsize), mmin(mmin_), mmax(mmax_), nmin(nmin_), nmax(nmax_) {
  XTPB_REQUIRE(auxsize > 0 && mmax_ >= mmin_ && nmax_ >= nmin_ && mmin_ >= 0 && nmin_ >= 0, "bad TCMatrix ranges");
  mtotal = mmax - mmin + 1;
  ntotal_glob = nmax - nmin + 1;
  rank = c->rank;
  world = c->world;
  XTPB_REQUIRE(ntotal_glob >= world, "fewer second-index levels than ranks");
  ntotal = nloc_below(ntotal_glob);
  ldn = round_up(ntotal, 2);
  slab = naux * ldn;
  M.alloc((size_t)(mtotal * slab));
  M.zero(ctx->stream);
}

TCMatrix::~TCMatrix() {
  if (prefetch.th.joinable()) prefetch.th.join();
  for (int b = 0; b < 2; ++b) {
    if (ev_copied[b]) cudaEventDestroy(ev_copied[b]);
    if (ev_consumed[b]) cudaEventDestroy(ev_consumed[b]);
  }
  if (copy_stream) cudaStreamDestroy(copy_stream);
}

void TCMatrix::set_raw(const double* host) {
  pending = false;
  metric_src = MetricSources{};
  eps0.valid = false;
  ++generation;
  ++content_gen;
  if (world == 1) {
    ctx->h2d_2d(M.p, ldn, host, ntotal, ntotal, mtotal * naux);
    ctx->sync();
    return;
  }
  // every rank is handed the whole tensor and keeps its cyclic share of the columns
  DBuf full((size_t)(naux * ntotal_glob));
  for (long long m = 0; m < mtotal; ++m) {
    ctx->h2d(full.p, host + m * naux * ntotal_glob, (size_t)(naux * ntotal_glob));
    k_cols_full_to_local(slab_ptr(m), ldn, full.p, ntotal_glob, naux, ntotal, rank, world, ctx->stream);
  }
  ctx->sync();
}

void TCMatrix::get_slab(long long m, double* host) {
  XTPB_REQUIRE(m >= 0 && m < mtotal, "slab index out of range");
  flush();
  if (world == 1) {
    ctx->d2h_2d(host, ntotal, slab_ptr(m), ldn, ntotal, naux);
    return;
  }
  // collective: local columns scattered into a zeroed full slab, summed over ranks
  DBuf full((size_t)(naux * ntotal_glob));
  full.zero(ctx->stream);
  k_cols_local_to_full(full.p, ntotal_glob, slab_ptr(m), ldn, naux, ntotal, rank, world, ctx->stream);
  ctx->allreduce_sum(full.p, (size_t)(naux * ntotal_glob));
  ctx->d2h(host, full.p, (size_t)(naux * ntotal_glob));
}

const double* TCMatrix::local_energies(const double* e_glob_dev, DBuf& tmp) {
  if (world == 1) return e_glob_dev;
  tmp.ensure((size_t)ntotal);
  k_gather_strided(tmp.p, e_glob_dev, rank, world, ntotal, ctx->stream);
  return tmp.p;
}

void TCMatrix::fill_begin(long long nb, const double* C_host, long long ldc_host) {
  XTPB_REQUIRE(nb > 0 && ldc_host >= nb, "bad MO coefficient matrix");
  pending = false;            // a new fill starts from the un-rotated tensor
  metric_src = MetricSources{};
  ppm_pre.armed = ppm_pre.complete = false;
  eps0.valid = false;
  ++generation;
  ++content_gen;
  n_basis = nb;
  ldc = round_up(nb, 2);
  Cm.alloc((size_t)(ldc * mtotal));
  Cn.alloc((size_t)(ldc * ntotal));
  Cm.zero(ctx->stream);
  Cn.zero(ctx->stream);
  ctx->h2d_2d(Cm.p, ldc, C_host + mmin * ldc_host, ldc_host, nb, mtotal);
  // local second-index columns: host columns nmin + rank, nmin + rank + world, ... (pitch world*ldc_host)
  ctx->h2d_2d(Cn.p, ldc, C_host + (nmin + rank) * ldc_host, ldc_host * world, nb, ntotal);
  ctx->sync();
}

// XTPB_SHARD_DENSE=0: every rank forms the replicated N_aux^3 products (metric / PPM sandwiches, folded rotation) whole
static bool shard_dense_products() {
  static const bool on = [] { const char* e = getenv("XTPB_SHARD_DENSE"); return !(e && e[0] == '0'); }();
  return on;
}

static bool fill_merge_batches() {
  static const bool on = [] { const char* e = getenv("XTPB_FILL_MERGE"); return !(e && e[0] == '0'); }();
  return on;
}

// Fill3cMO for aux functions P0..P0+nP-1 (upstream: per aux function dftn^T * T_P * dftm; here batched over P):
//   W_P  = T_P * C_m            (n_basis x mtotal)      2 n_basis^2 mtotal flops per P
//   M[m][P][:] = C_n^T * W_P    (ntotal x mtotal)       2 ntotal n_basis mtotal flops per P
void TCMatrix::fill_block_dev(long long P0, long long nP, const double* ao, long long ld_ao) {
  XTPB_REQUIRE(n_basis > 0, "fill_begin must be called before fill_block");
  XTPB_REQUIRE(world == 1, "with more than one rank use the collective fill (xtpb_tc_fill_sharded_packed)");
  XTPB_REQUIRE(P0 >= 0 && nP >= 0 && P0 + nP <= naux && ld_ao >= n_basis, "bad aux block");
  ++generation;
  ++content_gen;
  ProfScope prof(PROF_FILL);
  const long long ldw = round_up(n_basis, 2);
  const long long wslice = ldw * mtotal;
  const long long sub_max = std::max<long long>(1, std::min<long long>(128, (1LL << 27) / std::max<long long>(1, wslice)));
  ctx->scratch_a.ensure((size_t)(sub_max * wslice));
  double* W = ctx->scratch_a.p;
  for (long long p = 0; p < nP; p += sub_max) {
    const long long cnt = std::min(sub_max, nP - p);
    GemmParams g{};
    // W(mu, m) = sum_nu T(mu,nu) Cm(nu,m):   A(row mu, k nu) = T[nu + mu*ld]  (T symmetric)
    g.A = GemmOperand{ao + p * ld_ao * n_basis, ld_ao, 1, 0, ld_ao * n_basis};
    g.A.library_owned = 0;        // may be the caller's device memory (xtpb_tc_fill_block_dev)
    g.B = GemmOperand{Cm.p, ldc, 1, 0, 0};
    g.C = W; g.c_sm = 1; g.c_sn = ldw; g.c_batch = wslice;
    g.M = (int)n_basis; g.N = (int)mtotal; g.K = (int)n_basis; g.n_outer = 1; g.n_batch = (int)cnt;
    g.alpha = 1.0; g.beta = 0.0;
    GemmParams h{};
    // out(n, m) = sum_mu Cn(mu,n) W(mu,m) -> M[m][P0+p+b][n]
    h.A = GemmOperand{Cn.p, ldc, 1, 0, 0};
    h.B = GemmOperand{W, ldw, 1, 0, wslice};
    h.C = M.p + (P0 + p) * ldn; h.c_sm = 1; h.c_sn = slab; h.c_batch = ldn;
    h.M = (int)ntotal; h.N = (int)mtotal; h.K = (int)n_basis; h.n_outer = 1; h.n_batch = (int)cnt;
    h.alpha = 1.0; h.beta = 0.0;
    if (fill_merge_batches() && cnt * n_basis < (1LL << 31) && cnt * mtotal < (1LL << 31)) {
      // The slices of a group are equally spaced, so the batch index folds into an operand row index: rows (b, mu) of
      // the first product, columns (b, m) of the second (two-level output maps) -- one tile grid without the padding
      // of n_basis and mtotal to whole tiles in every batch (C60 size: 1860 -> 1920 rows, 360 -> 384 columns).
      g.M = (int)(cnt * n_basis); g.n_batch = 1; g.c_m_inner = (int)n_basis; g.c_sm_outer = wslice;
      h.N = (int)(cnt * mtotal); h.n_batch = 1; h.c_n_inner = (int)mtotal; h.c_sn_outer = ldn;
    }
    contract(g, ctx->ws, ctx->stream);
    contract(h, ctx->ws, ctx->stream);
    ppm_prefetch_advance(P0 + p, cnt);
  }
}

void TCMatrix::ppm_prefetch_begin(const double* e_host, long long n_occ_, double eta_) {
  XTPB_REQUIRE(n_basis > 0, "fill_begin must be called before the PPM prefetch hint");
  ppm_pre.armed = ppm_pre.complete = false;
  static const bool on = [] { const char* e = getenv("XTPB_PPM_PREFETCH"); return !(e && e[0] == '0'); }();
  if (!on || world != 1 || nmin != mmin || n_occ_ <= 0 || n_occ_ >= ntotal || n_occ_ > mtotal) return;   // hint only
  ppm_pre.energies.assign(e_host, e_host + ntotal);
  ppm_pre.n_occ = n_occ_;
  ppm_pre.eta = eta_;
  ppm_pre.filled_upto = ppm_pre.done_upto = 0;
  const int a0 = (int)(n_occ_ & ~1LL), K = (int)(ntotal - a0);
  ppm_pre.E.ensure((size_t)(2 * naux * naux));
  ppm_pre.d.ensure((size_t)(2 * n_occ_ * K + 2));
  DBuf e_dev((size_t)ntotal);
  ctx->h2d(e_dev.p, e_host, (size_t)ntotal);
  double* om_dev = ppm_pre.d.p + 2 * n_occ_ * K;
  const double om[2] = {0.0, 0.5};                  // screening_r, screening_i of the plasmon-pole model [Ha]
  ctx->h2d(om_dev, om, 2);
  k_chi0_weights(ppm_pre.d.p, e_dev.p, e_dev.p, (int)n_occ_, (int)n_occ_, a0, K, om_dev, 1, false, eta_, ctx->stream);
  k_chi0_weights(ppm_pre.d.p + n_occ_ * K, e_dev.p, e_dev.p, (int)n_occ_, (int)n_occ_, a0, K, om_dev + 1, 1, true, eta_,
                 ctx->stream);
  ctx->sync();                                      // e_dev is freed on return
  ppm_pre.armed = true;
}

void TCMatrix::ppm_prefetch_advance(long long P0, long long nP) {
  if (!ppm_pre.armed) return;
  if (P0 != ppm_pre.filled_upto) {                  // out of order: the panels would have holes
    ppm_pre.armed = false;
    return;
  }
  ppm_pre.filled_upto += nP;
  const long long kPanel = 256;
  const long long P1 = ppm_pre.filled_upto == naux
                           ? naux
                           : ppm_pre.done_upto + (ppm_pre.filled_upto - ppm_pre.done_upto) / kPanel * kPanel;
  if (P1 > ppm_pre.done_upto) {
    const long long Pa = ppm_pre.done_upto;
    const int a0 = (int)(ppm_pre.n_occ & ~1LL), K = (int)(ntotal - a0);
    ProfScope prof(PROF_EPSILON);
    GemmParams g{};
    g.A = GemmOperand{M.p + a0 + Pa * ldn, ldn, 1, slab, 0};          // rows P in [Pa, P1), outer = occupied level
    g.B = GemmOperand{M.p + a0, ldn, 1, slab, 0};                      // rows Q in [0, P1)
    g.C = ppm_pre.E.p + Pa; g.c_sm = 1; g.c_sn = naux; g.c_batch = naux * naux;
    g.d = ppm_pre.d.p; g.d_outer = K; g.d_batch = ppm_pre.n_occ * K;
    g.M = (int)(P1 - Pa); g.N = (int)P1; g.K = K; g.n_outer = (int)ppm_pre.n_occ; g.n_batch = 2;
    g.alpha = 1.0; g.beta = 0.0;
    contract(g, ctx->ws, ctx->stream);
    ppm_pre.done_upto = P1;
  }
  if (ppm_pre.done_upto == naux) {
    ppm_pre.armed = false;
    ppm_pre.complete = true;
    ppm_pre.content_gen = content_gen;
  }
}

bool TCMatrix::ppm_prefetch_take(const std::vector<double>& e, long long n_occ_, double eta_, double omega, bool imag,
                                 double* out_dev) {
  if (!ppm_pre.complete || ppm_pre.content_gen != content_gen || n_occ_ != ppm_pre.n_occ || eta_ != ppm_pre.eta ||
      e.size() != ppm_pre.energies.size() || !std::equal(e.begin(), e.end(), ppm_pre.energies.begin()))
    return false;
  int idx = -1;
  if (!imag && omega == 0.0) idx = 0;
  if (imag && omega == 0.5) idx = 1;
  if (idx < 0) return false;
  XTPB_CUDA(cudaMemcpyAsync(out_dev, ppm_pre.E.p + (long long)idx * naux * naux, (size_t)(naux * naux) * 8,
                            cudaMemcpyDeviceToDevice, ctx->stream));
  ++ppm_pre.taken;
  return true;
}

// Host-side Fill3cMO for a range of aux functions: the caller's AO slices (full symmetric n_basis x n_basis with
// leading dimension ld_ao, or packed lower triangles, row mu holding nu = 0..mu) stream through two device staging
// buffers on a dedicated copy stream, so the H2D copy of chunk i+1 overlaps the contractions of chunk i.  With
// pinned host memory (xtpb_host_alloc) the copies run at PCIe rate; pageable memory works but serialises.
void TCMatrix::fill_block_host(long long P0, long long nP, const double* ao_host, long long ld_ao, bool packed) {
  XTPB_REQUIRE(n_basis > 0, "fill_begin must be called before fill_block");
  XTPB_REQUIRE(P0 >= 0 && nP >= 0 && P0 + nP <= naux, "bad aux block");
  if (nP == 0) return;
  const long long ldt = round_up(n_basis, 2);
  const long long full_slice = ldt * n_basis;
  const long long host_slice = packed ? n_basis * (n_basis + 1) / 2 : ld_ao * n_basis;
  const long long dev_slice = packed ? host_slice : full_slice;
  // chunk = what one staging buffer holds (<= 512 MiB of full slices).  With the PPM prefetch armed the contractions come
  // in bursts (a 256-row epsilon panel every few chunks, ~30 ms at C60 size) and the copy stream can only run one
  // chunk ahead of them, so the chunks are made long enough (<= 4 GiB) for a burst to fit underneath one copy
  const long long stage_doubles = ppm_pre.armed ? (1LL << 29) : (1LL << 26);
  const long long sub_max = std::max<long long>(1, std::min<long long>(nP, stage_doubles / full_slice));
  for (int b = 0; b < 2; ++b) stage2[b].ensure((size_t)(sub_max * dev_slice));
  if (packed) unpacked.ensure((size_t)(sub_max * full_slice));
  if (!copy_stream) {
    XTPB_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
      XTPB_CUDA(cudaEventCreateWithFlags(&ev_copied[b], cudaEventDisableTiming));
      XTPB_CUDA(cudaEventCreateWithFlags(&ev_consumed[b], cudaEventDisableTiming));
    }
  }
  // staging buffers may have been (re)allocated on the compute stream's timeline: order the copy stream behind it
  XTPB_CUDA(cudaEventRecord(ev_consumed[0], ctx->stream));
  XTPB_CUDA(cudaEventRecord(ev_consumed[1], ctx->stream));
  long long chunk = 0;
  for (long long p = 0; p < nP; p += sub_max, ++chunk) {
    const int b = (int)(chunk & 1);
    const long long cnt = std::min(sub_max, nP - p);
    XTPB_CUDA(cudaStreamWaitEvent(copy_stream, ev_consumed[b], 0));
    const double* src = ao_host + p * host_slice;
    if (packed || ld_ao == ldt) {
      XTPB_CUDA(cudaMemcpyAsync(stage2[b].p, src, (size_t)(cnt * dev_slice) * 8, cudaMemcpyHostToDevice, copy_stream));
    } else {
      XTPB_CUDA(cudaMemcpy2DAsync(stage2[b].p, ldt * 8, src, ld_ao * 8, n_basis * 8, n_basis * cnt,
                                  cudaMemcpyHostToDevice, copy_stream));
    }
    XTPB_CUDA(cudaEventRecord(ev_copied[b], copy_stream));
    XTPB_CUDA(cudaStreamWaitEvent(ctx->stream, ev_copied[b], 0));
    if (packed) {
      fill_block_packed_dev(P0 + p, cnt, stage2[b].p);
    } else {
      fill_block_dev(P0 + p, cnt, stage2[b].p, ldt);
    }
    XTPB_CUDA(cudaEventRecord(ev_consumed[b], ctx->stream));
  }
  ctx->sync();
}

// packed lower-triangular slices already on the device: unpack a bounded group into full symmetric matrices, contract.
void TCMatrix::fill_block_packed_dev(long long P0, long long nP, const double* packed_dev) {
  XTPB_REQUIRE(n_basis > 0, "fill_begin must be called before fill_block");
  XTPB_REQUIRE(P0 >= 0 && nP >= 0 && P0 + nP <= naux, "bad aux block");
  const long long ldt = round_up(n_basis, 2);
  const long long full_slice = ldt * n_basis, pk_slice = n_basis * (n_basis + 1) / 2;
  // up to 2 GiB of unpacked slices per group: the two contractions of a group are one launch each, and the longer the
  // launch the smaller the share of its last, partly filled wave of tiles (C60 size: 64 slices = 19.5 waves)
  const long long cap = std::max<long long>(1, std::min<long long>(nP, (1LL << 28) / full_slice));
  const long long groups = (nP + cap - 1) / std::max<long long>(cap, 1);
  const long long sub_max = groups > 0 ? (nP + groups - 1) / groups : 1;     // equal groups, none of them short
  unpacked.ensure((size_t)(sub_max * full_slice));
  for (long long p = 0; p < nP; p += sub_max) {
    const long long cnt = std::min(sub_max, nP - p);
    k_unpack_symmetric(unpacked.p, ldt, full_slice, packed_dev + p * pk_slice, pk_slice, (int)n_basis, (int)cnt,
                       ctx->stream);
    fill_block_dev(P0 + p, cnt, unpacked.p, ldt);
  }
}

// Collective Fill3cMO over all ranks (one process per GPU).  Rank r holds the packed AO slices of its aux range;
// per round every rank half-transforms up to `B` of its slices (W_P = T_P C_m, the n_basis^2 m part of the work,
// split over P), the W blocks are all-gathered over NVLink on the communication stream, and every rank finishes
// all gathered aux functions for its own share of the second index (M[m][P][n_loc] = C_nloc^T W_P, split over n).
// The all-gather of round i overlaps the first half of round i+1 and the second half of round i-1.
void TCMatrix::fill_sharded_packed(const double* packed, bool on_device) {
  XTPB_REQUIRE(n_basis > 0, "fill_begin must be called before fill");
  long long lo, hi;
  aux_range(rank, lo, hi);
  ++generation;
  ++content_gen;
  if (world == 1) {
    if (on_device) fill_block_packed_dev(lo, hi - lo, packed);
    else fill_block_host(lo, hi - lo, packed, 0, true);
    ctx->sync();
    return;
  }
  const long long ldt = round_up(n_basis, 2);
  const long long full_slice = ldt * n_basis, pk_slice = n_basis * (n_basis + 1) / 2;
  const long long ldw = ldt, wslice = ldw * mtotal;
  long long maxcnt = 0;
  for (int r = 0; r < world; ++r) {
    long long a, b;
    aux_range(r, a, b);
    maxcnt = std::max(maxcnt, b - a);
  }
  // round size: bounded gather buffer (2 x world x B x wslice doubles <= ~4 GiB) and unpack scratch
  long long B = std::min<long long>(32, std::max<long long>(1, (1LL << 28) / (world * wslice)));
  B = std::min(B, std::max<long long>(1, (1LL << 26) / full_slice));
  B = std::min(B, maxcnt);
  const long long rounds = (maxcnt + B - 1) / B;
  DBuf wsend[2], wall[2];
  for (int b = 0; b < 2; ++b) {
    wsend[b].alloc((size_t)(B * wslice));
    wsend[b].zero(ctx->stream);
    wall[b].alloc((size_t)(world * B * wslice));
  }
  unpacked.ensure((size_t)(B * full_slice));
  // second half: ONE batched launch per round over all world*B gathered slices (a launch per source rank has only
  // ~1.3 waves of tiles at 8 ranks) into a staging tensor [m][slot][n_loc]; the slots are then scattered to their aux
  // index (the canonical aux ranges of the ranks differ in length by one, so the slot -> P map is not one stride)
  ctx->scratch_b.ensure((size_t)(mtotal * world * B * ldn));
  double* stage_out = ctx->scratch_b.p;
  if (!on_device) for (int b = 0; b < 2; ++b) stage2[b].ensure((size_t)(B * pk_slice));
  if (!copy_stream) {
    XTPB_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
      XTPB_CUDA(cudaEventCreateWithFlags(&ev_copied[b], cudaEventDisableTiming));
      XTPB_CUDA(cudaEventCreateWithFlags(&ev_consumed[b], cudaEventDisableTiming));
    }
  }
  cudaEvent_t ev_w[2], ev_g[2], ev_free[2];
  for (int b = 0; b < 2; ++b) {
    XTPB_CUDA(cudaEventCreateWithFlags(&ev_w[b], cudaEventDisableTiming));
    XTPB_CUDA(cudaEventCreateWithFlags(&ev_g[b], cudaEventDisableTiming));
    XTPB_CUDA(cudaEventCreateWithFlags(&ev_free[b], cudaEventDisableTiming));
    XTPB_CUDA(cudaEventRecord(ev_free[b], ctx->stream));
    XTPB_CUDA(cudaEventRecord(ev_consumed[b], ctx->stream));
  }
  cudaStream_t cs = ctx->comm_stream;
  auto second_half = [&](long long i) {      // consumes wall[i&1]
    const int b = (int)(i & 1);
    XTPB_CUDA(cudaStreamWaitEvent(ctx->stream, ev_g[b], 0));
    ProfScope prof(PROF_FILL);
    const long long slots = (long long)world * B;
    GemmParams h{};
    h.A = GemmOperand{Cn.p, ldc, 1, 0, 0};
    h.B = GemmOperand{wall[b].p, ldw, 1, 0, wslice};
    h.C = stage_out; h.c_sm = 1; h.c_sn = slots * ldn; h.c_batch = ldn;
    h.M = (int)ntotal; h.N = (int)mtotal; h.K = (int)n_basis; h.n_outer = 1; h.n_batch = (int)slots;
    h.alpha = 1.0; h.beta = 0.0;
    if (fill_merge_batches() && slots * mtotal < (1LL << 31)) {     // columns (slot, m): see fill_block_dev
      h.N = (int)(slots * mtotal); h.n_batch = 1; h.c_n_inner = (int)mtotal; h.c_sn_outer = ldn;
    }
    contract(h, ctx->ws, ctx->stream);
    XTPB_CUDA(cudaEventRecord(ev_free[b], ctx->stream));      // the gathered blocks are consumed
    k_scatter_fill_slots(M.p, ldn, slab, stage_out, (int)mtotal, (int)ntotal, (int)naux, world, (int)B, (int)i,
                         ctx->stream);
  };
  for (long long i = 0; i < rounds; ++i) {
    const int b = (int)(i & 1);
    const long long p0 = lo + i * B;
    const long long cnt = std::max<long long>(0, std::min(B, hi - p0));
    if (cnt > 0) {
      const double* src = packed + (p0 - lo) * pk_slice;
      if (!on_device) {
        XTPB_CUDA(cudaStreamWaitEvent(copy_stream, ev_consumed[b], 0));
        XTPB_CUDA(cudaMemcpyAsync(stage2[b].p, src, (size_t)(cnt * pk_slice) * 8, cudaMemcpyHostToDevice, copy_stream));
        XTPB_CUDA(cudaEventRecord(ev_copied[b], copy_stream));
        XTPB_CUDA(cudaStreamWaitEvent(ctx->stream, ev_copied[b], 0));
        src = stage2[b].p;
      }
      k_unpack_symmetric(unpacked.p, ldt, full_slice, src, pk_slice, (int)n_basis, (int)cnt, ctx->stream);
      if (!on_device) XTPB_CUDA(cudaEventRecord(ev_consumed[b], ctx->stream));
      ProfScope prof(PROF_FILL);
      GemmParams g{};
      g.A = GemmOperand{unpacked.p, ldt, 1, 0, full_slice};
      g.B = GemmOperand{Cm.p, ldc, 1, 0, 0};
      g.C = wsend[b].p; g.c_sm = 1; g.c_sn = ldw; g.c_batch = wslice;
      g.M = (int)n_basis; g.N = (int)mtotal; g.K = (int)n_basis; g.n_outer = 1; g.n_batch = (int)cnt;
      g.alpha = 1.0; g.beta = 0.0;
      if (fill_merge_batches() && cnt * n_basis < (1LL << 31)) {     // rows (b, mu): see fill_block_dev
        g.M = (int)(cnt * n_basis); g.n_batch = 1; g.c_m_inner = (int)n_basis; g.c_sm_outer = wslice;
      }
      contract(g, ctx->ws, ctx->stream);
    }
    XTPB_CUDA(cudaEventRecord(ev_w[b], ctx->stream));
    XTPB_CUDA(cudaStreamWaitEvent(cs, ev_w[b], 0));
    XTPB_CUDA(cudaStreamWaitEvent(cs, ev_free[b], 0));
    ctx->allgather(wsend[b].p, wall[b].p, (size_t)(B * wslice), cs);
    XTPB_CUDA(cudaEventRecord(ev_g[b], cs));
    if (i > 0) second_half(i - 1);
  }
  second_half(rounds - 1);
  ctx->sync();
  XTPB_CUDA(cudaStreamSynchronize(cs));
  for (int b = 0; b < 2; ++b) {
    cudaEventDestroy(ev_w[b]);
    cudaEventDestroy(ev_g[b]);
    cudaEventDestroy(ev_free[b]);
  }
}

// dst[i][Q][j] = sum_P M[m0+i][P][n0+j] R[P,Q]; A rows-contiguous (j), B = R K-contiguous (column Q of R).
void TCMatrix::rotate_window(double* dst, long long dst_ld, long long dst_slab, int m0, int mcnt, int n0, int ncnt,
                             const double* R_dev, long long ldr) {
  flush();
  ProfScope prof(PROF_ROTATE);
  GemmParams g{};
  g.A = GemmOperand{M.p + (long long)m0 * slab + n0, 1, ldn, 0, slab};
  g.B = GemmOperand{R_dev, ldr, 1, 0, 0};
  g.C = dst; g.c_sm = 1; g.c_sn = dst_ld; g.c_batch = dst_slab;
  g.M = ncnt; g.N = (int)naux; g.K = (int)naux; g.n_outer = 1;
  g.alpha = 1.0; g.beta = 0.0;
  for (int b = 0; b < mcnt; b += 16384) {      // gridDim.z limit
    GemmParams gg = g;
    gg.n_batch = std::min(16384, mcnt - b);
    gg.A.p += (long long)b * slab;
    gg.C += (long long)b * dst_slab;
    contract(gg, ctx->ws, ctx->stream);
  }
}

// Coulomb-metric prefetch: A <- eigenvectors, w <- eigenvalues of the n_aux x n_aux host matrix, on a helper thread
// with its own high-priority stream and cuSOLVER handle, so that the dense eigensolver (latency-bound, ~0.2 s at
// N_aux 5500, replicated on every rank) runs underneath Fill3cMO instead of after it.
void TCMatrix::metric_prefetch_begin(const double* X_host, long long ldx, bool of_overlap) {
  XTPB_REQUIRE(X_host && ldx >= naux, "bad matrix for the Coulomb-metric prefetch");
  if (prefetch.th.joinable()) prefetch.th.join();
  prefetch.err = nullptr;
  prefetch.active = true;
  prefetch.of_overlap = of_overlap;
  prefetch.src = X_host;
  prefetch.U.ensure((size_t)(naux * naux));
  prefetch.w.ensure((size_t)naux);
  prefetch.lam.assign((size_t)naux, 0.0);
  Context* c = ctx;
  c->side_init();
  c->eigh_async_join();       // the helper stream / handle / workspace are shared with Context::eigh_async
  int lwork = 0;
  if (cusolverDnDsyevd_bufferSize(c->side_solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)naux,
                                  prefetch.U.p, (int)naux, prefetch.w.p, &lwork) != CUSOLVER_STATUS_SUCCESS)
    throw Error("xtpb: cusolverDnDsyevd_bufferSize failed");
  c->side_work.ensure((size_t)lwork);
  c->sync();      // buffers above may come out of the block cache: nothing on the main stream still uses them
  const long long na = naux;
  prefetch.th = std::thread([this, c, X_host, ldx, na, lwork] {
    try {
      XTPB_CUDA(cudaSetDevice(c->device));
      XTPB_CUDA(cudaMemcpy2DAsync(prefetch.U.p, na * 8, X_host, ldx * 8, na * 8, na, cudaMemcpyHostToDevice,
                                  c->side_stream));
      const cusolverStatus_t st =
          cusolverDnDsyevd(c->side_solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)na, prefetch.U.p,
                           (int)na, prefetch.w.p, c->side_work.p, lwork, c->side_info);
      if (st != CUSOLVER_STATUS_SUCCESS) throw Error("xtpb: cusolverDnDsyevd failed in the Coulomb-metric prefetch");
      int info = 0;
      XTPB_CUDA(cudaMemcpyAsync(prefetch.lam.data(), prefetch.w.p, (size_t)na * 8, cudaMemcpyDeviceToHost,
                                c->side_stream));
      XTPB_CUDA(cudaMemcpyAsync(&info, c->side_info, sizeof(int), cudaMemcpyDeviceToHost, c->side_stream));
      XTPB_CUDA(cudaStreamSynchronize(c->side_stream));
      if (info != 0) throw Error("xtpb: cuSOLVER devInfo " + std::to_string(info) + " in the Coulomb-metric prefetch");
    } catch (...) {
      prefetch.err = std::current_exception();
    }
  });
}

bool TCMatrix::metric_prefetch_join() {
  if (!prefetch.active) return false;
  if (prefetch.th.joinable()) prefetch.th.join();
  prefetch.active = false;
  if (prefetch.err) std::rethrow_exception(prefetch.err);
  return true;
}

// MultiplyRightWithAuxMatrix: out of place through a bounded scratch of `chunk` slabs, copied back.
void TCMatrix::set_pending(const double* R_dev, long long ldr) {
  flush();
  eps0.valid = false;
  ++generation;
  static const bool lazy = [] { const char* e = getenv("XTPB_LAZY_METRIC"); return !(e && e[0] == '0'); }();
  if (!lazy) {
    rotate(R_dev, ldr);
    return;
  }
  pendR.ensure((size_t)(naux * naux));
  k_copy_2d(pendR.p, naux, R_dev, ldr, (int)naux, naux, ctx->stream);
  pending = true;
}

void TCMatrix::flush() {
  if (!pending) return;
  if (metric_src.cholesky) {
    // somebody needs the metric-rotated tensor itself: build the reference's symmetric factor after all, from the
    // retained matrices (the Cholesky tests have shown that it removes no function)
    DBuf R;
    metric_factor_eig(metric_src.V.p, metric_src.has_S ? metric_src.S.p : nullptr, metric_src.etol, false, R);
    XTPB_CUDA(cudaMemcpyAsync(pendR.p, R.p, (size_t)(naux * naux) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->sync();
    metric_src = MetricSources{};
  }
  pending = false;
  rotate(pendR.p, naux);
}

// the Cholesky path needs the deferred rotation (the factor must stay pending until the PPM eigenvectors absorb it)
static bool metric_cholesky_enabled() {
  static const bool on = [] {
    const char* c = getenv("XTPB_METRIC_CHOLESKY");
    const char* l = getenv("XTPB_LAZY_METRIC");
    return !(c && c[0] == '0') && !(l && l[0] == '0');
  }();
  return on;
}

void TCMatrix::metric_hint(const double* V_host, long long ldv, const double* S_host, long long lds) {
  hint.given = true;
  hint.V = V_host;
  hint.S = S_host;
  // eigen path only: start its first eigendecomposition underneath the fill.  With the Cholesky path the decision
  // needs etol (known at apply time) and ~0.04 s of factorisations, so nothing is started here.
  if (!metric_cholesky_enabled()) {
    if (S_host) metric_prefetch_begin(S_host, lds, true);
    else metric_prefetch_begin(V_host, ldv, false);
  }
}

// f(X) = U diag(1/sqrt(lambda) | 0) U^T for symmetric X, eigenvalues < etol dropped; the reference's construction
long long TCMatrix::metric_factor_eig(double* A, double* S, double etol, bool prefetched, DBuf& R_out) {
  const long long na = naux;
  long long removed = 0;
  DBuf B((size_t)(na * na)), Cc((size_t)(na * na)), w((size_t)na), Ssqrt, Vm1((size_t)(na * na));
  std::vector<double> lam((size_t)na), sc((size_t)na);
  auto mm = [&](const double* X, const double* Yp, double* Z) {
    // Z = X * Y for symmetric X (read through its K-contiguous view), Y column-major
    GemmParams g{};
    g.A = op_k_contig(X, na);
    g.B = op_k_contig(Yp, na);
    g.C = Z; g.c_sm = 1; g.c_sn = na;
    g.M = g.N = g.K = (int)na; g.n_outer = 1; g.n_batch = 1; g.alpha = 1.0;
    contract(g, ctx->ws, ctx->stream);
  };
  auto inv_sqrt = [&](double* X, double* out) {
    if (prefetched) {                            // eigenvectors / eigenvalues are already there
      prefetched = false;
      XTPB_CUDA(cudaMemcpyAsync(X, prefetch.U.p, (size_t)(na * na) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      lam = prefetch.lam;
    } else {
      ctx->eigh((int)na, X, na, w.p);            // X <- U
      ctx->d2h(lam.data(), w.p, (size_t)na);
    }
    for (long long i = 0; i < na; ++i) {
      if (lam[i] < etol) { ++removed; sc[i] = 0.0; } else sc[i] = 1.0 / std::sqrt(lam[i]);
    }
    ctx->h2d(w.p, sc.data(), (size_t)na);
    XTPB_CUDA(cudaMemcpyAsync(B.p, X, (size_t)(na * na) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    k_scale_columns(B.p, (int)na, (int)na, na, w.p, ctx->stream);   // B = U diag(s)
    // out = B U^T : out(i,j) = sum_k B(i,k) U(j,k)  -> A rows-contig (B), B-operand rows-contig (U)
    GemmParams g{};
    g.A = op_rows_contig(B.p, na);
    g.B = op_rows_contig(X, na);
    g.C = out; g.c_sm = 1; g.c_sn = na;
    g.M = g.N = g.K = (int)na; g.n_outer = 1; g.n_batch = 1; g.alpha = 1.0;
    contract(g, ctx->ws, ctx->stream);
  };
  if (S) {
    Ssqrt.alloc((size_t)(na * na));
    inv_sqrt(S, Ssqrt.p);
    // ortho = Ssqrt V Ssqrt  (all symmetric)
    mm(Ssqrt.p, A, Cc.p);
    // Cc is not symmetric: (Cc * Ssqrt)(i,j) = sum_k Cc(i,k) Ssqrt(k,j) -> rows-contiguous view of Cc
    GemmParams g{};
    g.A = op_rows_contig(Cc.p, na);
    g.B = op_k_contig(Ssqrt.p, na);
    g.C = A; g.c_sm = 1; g.c_sn = na;
    g.M = g.N = g.K = (int)na; g.n_outer = 1; g.n_batch = 1; g.alpha = 1.0;
    contract(g, ctx->ws, ctx->stream);
  }
  inv_sqrt(A, Vm1.p);
  if (S) {
    // R = S^-1/2 (S^-1/2 V S^-1/2)^-1/2  ("((S-1/2 V S-1/2)-1/2 S-1/2)T" in upstream's words): R R^T = V^-1
    mm(Ssqrt.p, Vm1.p, Cc.p);      // Cc = Ssqrt * Vm1
    R_out = std::move(Cc);
  } else {
    R_out = std::move(Vm1);
  }
  ctx->sync();
  ++metric_eig_count;
  return removed;
}

long long TCMatrix::apply_coulomb_metric(const double* V_host, long long ldv, const double* S_host, long long lds,
                                         double etol) {
  const long long na = naux;
  if (hint.given) {
    const bool same = hint.V == V_host && hint.S == S_host;
    hint = MetricHint{};
    if (!same) {
      metric_prefetch_join();
      throw Error("xtpb: xtpb_tc_coulomb_metric_begin was given different matrices than xtpb_tc_apply_coulomb_metric");
    }
  }
  const bool prefetched = metric_prefetch_join();
  flush();                      // an earlier pending factor reaches the tensor in its reference (symmetric) form
  metric_src = MetricSources{};
  DBuf A((size_t)(na * na)), S;
  // a prefetched decomposition already holds the eigenvectors of its matrix on the device
  if (S_host || !prefetched) ctx->h2d_2d(A.p, na, V_host, ldv, na, na);
  if (S_host) {
    S.alloc((size_t)(na * na));
    if (!prefetched) ctx->h2d_2d(S.p, na, S_host, lds, na, na);
  }
  if (metric_cholesky_enabled() && !prefetched && etol > 0.0) {
    // would the reference remove a function?  S - etol > 0 and V - etol S > 0 (V - etol without an overlap), decided
    // by Cholesky factorisations of copies
    DBuf T((size_t)(na * na));
    bool none_removed = true;
    if (S_host) {
      XTPB_CUDA(cudaMemcpyAsync(T.p, S.p, (size_t)(na * na) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      k_add_diagonal(T.p, (int)na, na, -etol, ctx->stream);
      none_removed = ctx->cholesky((int)na, T.p, na, false);
    }
    if (none_removed) {
      XTPB_CUDA(cudaMemcpyAsync(T.p, A.p, (size_t)(na * na) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      if (S_host) k_axpby(T.p, S.p, na * na, -etol, 1.0, ctx->stream);
      else k_add_diagonal(T.p, (int)na, na, -etol, ctx->stream);
      none_removed = ctx->cholesky((int)na, T.p, na, false);
    }
    if (none_removed) {
      // V = U^T U (upper factor), R = U^-1: R R^T = (U^T U)^-1 = V^-1
      XTPB_CUDA(cudaMemcpyAsync(T.p, A.p, (size_t)(na * na) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      if (ctx->cholesky((int)na, T.p, na, true)) {
        ctx->tri_inverse((int)na, T.p, na, true);
        k_zero_strict_lower(T.p, (int)na, na, ctx->stream);
        set_pending(T.p, na);
        metric_src.cholesky = true;
        metric_src.has_S = S_host != nullptr;
        metric_src.etol = etol;
        metric_src.V = std::move(A);
        if (S_host) metric_src.S = std::move(S);
        ctx->sync();
        ++metric_cholesky_count;
        return 0;
      }
    }
  }
  DBuf R;
  const long long removed = metric_factor_eig(A.p, S_host ? S.p : nullptr, etol, prefetched, R);
  set_pending(R.p, na);       // deferred: folded into the next full rotation (see TCMatrix::set_pending)
  ctx->sync();
  return removed;
}

void TCMatrix::rotate(const double* R_dev, long long ldr, bool covariant) {
  if (pending && metric_src.cholesky && !covariant) flush();
  ++generation;
  ++content_gen;
  eps0.valid = false;         // the caller re-validates when R is an eps(0) eigenbasis (GW::prepare_ppm)
  DBuf folded;
  if (pending) {      // M <- M (Rp R): fold the deferred factor into this rotation
    pending = false;
    metric_src = MetricSources{};   // the folded product is the same for every Rp with Rp Rp^T = V^-1 (see internal.h)
    folded.alloc((size_t)(naux * naux));
    ProfScope prof(PROF_DENSE_AUX);
    // every rank holds Rp and R: with several ranks each forms a column block of the product, exchanged by one
    // grouped broadcast (see congruence_sym)
    const bool split = world > 1 && shard_dense_products() && naux >= 64LL * world;
    const long long j0 = split ? naux * rank / world : 0, j1 = split ? naux * (rank + 1) / world : naux;
    GemmParams g{};
    g.A = op_rows_contig(pendR.p, naux);
    g.B = op_k_contig(R_dev + j0 * ldr, ldr);
    g.C = folded.p + j0 * naux; g.c_sm = 1; g.c_sn = naux;
    g.M = g.K = (int)naux; g.N = (int)(j1 - j0); g.n_outer = 1; g.n_batch = 1; g.alpha = 1.0;
    contract(g, ctx->ws, ctx->stream);
    if (split) {
      ctx->group_start();
      for (int r = 0; r < world; ++r) {
        const long long a = naux * r / world, b = naux * (r + 1) / world;
        ctx->bcast(folded.p + a * naux, (size_t)((b - a) * naux), r);
      }
      ctx->group_end();
    }
    R_dev = folded.p;
    ldr = naux;
  }
  const long long budget = 1LL << 29;   // doubles (4 GiB)
  const long long chunk = std::max<long long>(1, std::min<long long>(mtotal, budget / slab));
  ctx->scratch_b.ensure((size_t)(chunk * slab));
  for (long long m = 0; m < mtotal; m += chunk) {
    const long long cnt = std::min(chunk, mtotal - m);
    rotate_window(ctx->scratch_b.p, ldn, slab, (int)m, (int)cnt, 0, (int)ntotal, R_dev, ldr);
    // padding column (ldn > ntotal) of the scratch is never written; copy only the payload rows
    k_copy_2d(slab_ptr(m), ldn, ctx->scratch_b.p, ldn, (int)ntotal, cnt * naux, ctx->stream);
  }
  if (folded.p) ctx->sync();      // folded is freed on return
}

// E <- R^T E R for a symmetric E (full storage on entry) and a general R, both n x n with ld = n; T: n x n scratch.
// One rank: T = E R, then the lower triangle of R^T T (3 n^3 flops; the upper triangle of the result is NOT written).
// Several ranks hold the same E and R: rank r forms the column block [n r / world, n (r+1) / world) of T and of the
// full result (4 n^3 / world flops instead of 3 n^3 on every rank) and the blocks are exchanged by one grouped
// broadcast (n^2 doubles in total); the result is then complete (both triangles) and identical on every rank.
void congruence_sym(Context* ctx, double* E, const double* R, double* T, long long n) {
  const int world = ctx->world;
  const bool split = world > 1 && shard_dense_products() && n >= 64LL * world;
  const long long j0 = split ? n * ctx->rank / world : 0, j1 = split ? n * (ctx->rank + 1) / world : n;
  GemmParams g{};
  g.A = op_k_contig(E, n);                              // E symmetric: E(i,k) = E[k + i n]
  g.B = op_k_contig(R + j0 * n, n);                     // columns j0..j1 of R
  g.C = T + j0 * n; g.c_sm = 1; g.c_sn = n;
  g.M = (int)n; g.N = (int)(j1 - j0); g.K = (int)n; g.n_outer = 1; g.n_batch = 1; g.alpha = 1.0;
  contract(g, ctx->ws, ctx->stream);
  GemmParams h{};
  h.A = op_k_contig(R, n);                              // (R^T)(i,k) = R[k + i n]
  h.B = op_k_contig(T + j0 * n, n);
  h.C = E + j0 * n; h.c_sm = 1; h.c_sn = n;
  h.M = (int)n; h.N = (int)(j1 - j0); h.K = (int)n; h.n_outer = 1; h.n_batch = 1; h.alpha = 1.0;
  h.lower = split ? 0 : 1;
  contract(h, ctx->ws, ctx->stream);
  if (split) {
    ctx->group_start();
    for (int r = 0; r < world; ++r) {
      const long long a = n * r / world, b = n * (r + 1) / world;
      ctx->bcast(E + a * n, (size_t)((b - a) * n), r);
    }
    ctx->group_end();
  }
}

// eps(w) = 1 + sum_{m occ} A_m^T diag(d_m(w)) A_m with A_m = M[m](unocc, :)   (upstream RPA::calculate_epsilon)
// One lower-triangular SYRK-style launch per call; the frequency index is the batch dimension.
void rpa_epsilon_dev(TCMatrix& tc, const double* energies_dev, long long n_occ, double eta, const double* omegas_host,
                     int n_omega, bool imag, double, double* out_dev, int owner_shift,
                     const std::vector<double>* energies_host) {
  Context* ctx = tc.ctx;
  ProfScope prof(PROF_EPSILON);
  XTPB_REQUIRE(n_occ > 0 && n_occ < tc.ntotal_glob && n_occ <= tc.mtotal, "RPA needs occupied and unoccupied levels");
  // second index: the local (cyclic) share of the unoccupied levels; first index: all occupied levels
  const long long n_occ_loc = tc.nloc_below(n_occ);
  const int a0 = (int)(n_occ_loc & ~1LL);          // 16-byte aligned start of the contraction range
  const int K = (int)(tc.ntotal - a0);
  DBuf e_loc_buf;
  const double* e_loc = tc.local_energies(energies_dev, e_loc_buf);
  DBuf d((size_t)n_omega * n_occ * std::max(K, 1) + n_omega);
  double* om_dev = d.p + (size_t)n_omega * n_occ * std::max(K, 1);
  ctx->h2d(om_dev, omegas_host, n_omega);
  const size_t out_count = (size_t)n_omega * tc.naux * tc.naux;
  // accumulated underneath Fill3cMO already (TCMatrix::PpmPrefetch)?  Then only the finishing steps remain.
  const bool prefetched = energies_host && n_omega == 1 && owner_shift < 0 && ctx->world == 1 &&
                          tc.ppm_prefetch_take(*energies_host, n_occ, eta, omegas_host[0], imag, out_dev);
  if (!prefetched && K > 0) {
    k_chi0_weights(d.p, energies_dev, e_loc, (int)n_occ, (int)n_occ_loc, a0, K, om_dev, n_omega, imag, eta,
                   ctx->stream);
    GemmParams g{};
    g.A = GemmOperand{tc.M.p + a0, tc.ldn, 1, tc.slab, 0};
    g.B = g.A;
    g.C = out_dev; g.c_sm = 1; g.c_sn = tc.naux; g.c_batch = tc.naux * tc.naux;
    g.d = d.p; g.d_outer = K; g.d_batch = (long long)n_occ * K;
    g.M = (int)tc.naux; g.N = (int)tc.naux; g.K = K; g.n_outer = (int)n_occ; g.n_batch = n_omega;
    g.alpha = 1.0; g.beta = 0.0; g.lower = 1;
    contract(g, ctx->ws, ctx->stream);
  } else if (!prefetched) {
    XTPB_CUDA(cudaMemsetAsync(out_dev, 0, out_count * 8, ctx->stream));
  }
  const bool sharded = owner_shift >= 0 && ctx->world > 1;
  auto mine = [&](int w) { return !sharded || (w + owner_shift) % ctx->world == ctx->rank; };
  if (sharded) {                                    // each frequency's matrix is summed onto its owner only
    XTPB_REQUIRE(!tc.pending, "frequency-sharded epsilon needs a flushed tensor");
    ctx->group_start();
    for (int w = 0; w < n_omega; ++w)
      ctx->reduce_sum(out_dev + (long long)w * tc.naux * tc.naux, (size_t)(tc.naux * tc.naux),
                      (w + owner_shift) % ctx->world);
    ctx->group_end();
  } else {
    ctx->allreduce_sum(out_dev, out_count);         // partial sums over the local unoccupied levels
  }
  if (tc.pending) {
    // the tensor still lacks the deferred aux rotation Rp: eps = 1 + Rp^T E Rp with E from the un-rotated tensor
    const long long na = tc.naux;
    DBuf T((size_t)(na * na));
    for (int w = 0; w < n_omega; ++w) {
      double* E = out_dev + (long long)w * na * na;
      symmetrize_from_lower(E, (int)na, na, 0.0, ctx->stream);
      congruence_sym(ctx, E, tc.pendR.p, T.p, na);
    }
    ctx->sync();   // T is freed at scope exit
  }
  for (int w = 0; w < n_omega; ++w)
    if (mine(w))
      symmetrize_from_lower(out_dev + (long long)w * tc.naux * tc.naux, (int)tc.naux, tc.naux, 1.0, ctx->stream);
  ctx->sync();   // d is freed on return
}

}  // namespace xtpb
