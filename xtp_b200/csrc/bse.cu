// BSE setup, the matrix-free BSE Hamiltonian, and the block Davidson solver.
// Upstream: xtp/src/libxtp/gwbse/bse.cc, bse_operator.{h,cc}, xtp/src/libxtp/davidsonsolver.cc.
//
// Unlike the reference, which rebuilds every row block of H on every matmul (2 v^2 c^2 N_aux flops per call,
// independent of the number of trial vectors), the products are factorised through the RI tensors:
//   Hx  X : T = Mvc^T X ; Y += cx Mvc T                                      4 vc N_aux k flops
//   Hd  X : U[k][c1][P][v2] = sum_c2 Mcc[c1][P][c2] X[k][v2,c2]              2 N_aux c^2 v k
//           Y[k][v1,c1]    -= sum_{P,v2} eps_inv[P] Mvv[v1][P][v2] U[...]    2 N_aux v^2 c k
//   Hqp X : two small products with the QP Hamiltonian blocks.
// Two strategies (BseOperator): "factorised" applies the products above on every call and never holds H;
// "dense" (XTPB_BSE_MODE=dense, or auto when it fits XTPB_BSE_DENSE_MAX_GB) builds the screened direct term + Hqp
// once as a (vc)^2 matrix in HBM and streams it per call, the exchange term staying factorised.  On a single rank the
// direct term is contracted for the occupied pairs v1 <= v2 only (Hd is symmetric under (v1,c1) <-> (v2,c2)) and
// mirrored by a scatter kernel: v (v+1) c^2 N_aux flops instead of 2 v^2 c^2 N_aux.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <numeric>

#include "internal.h"

namespace xtpb {

namespace {
constexpr double kRpaEtaDefault = 1e-3;   // RPA object's eta when BSE builds its own screening (oracle: RPA.eta)
constexpr int kHostEighMax = 768;         // projected eigenproblems up to this size are solved on the host
}

// ------------------------------------------------------------------ BSE::configure
BSE::BSE(Context* c, TCMatrix* t, const xtpb_bse_options& o, const double* rpa_e, const double* hqp_in, long long ldh,
         bool)
    : ctx(c), tc(t), opt(o) {
  XTPB_REQUIRE(o.rpamin == t->nmin && o.rpamax == t->nmax && o.rpamin == t->mmin, "TCMatrix ranges do not match rpamin/rpamax");
  XTPB_REQUIRE(o.vmin >= o.rpamin && o.vmin <= o.homo && o.cmax > o.homo && o.cmax <= t->mmax, "BSE window outside the TCMatrix");
  vt = o.homo - o.vmin + 1;
  ct = o.cmax - o.homo;
  size = vt * ct;
  naux = t->naux;
  // BSE::AdjustHqpSize
  const long long hsize = vt + ct, gwsize = o.qpmax - o.qpmin + 1, off = o.vmin - o.rpamin;
  std::vector<double> H((size_t)(hsize * hsize), 0.0);
  auto Hq = [&](long long i, long long j) { return hqp_in[i + j * ldh]; };
  if (o.vmin >= o.qpmin) {
    const long long start = o.vmin - o.qpmin;
    if (o.cmax <= o.qpmax) {
      for (long long j = 0; j < hsize; ++j)
        for (long long i = 0; i < hsize; ++i) H[i + j * hsize] = Hq(start + i, start + j);
    } else {
      const long long virtoffset = gwsize - start, extra = o.cmax - o.qpmax;
      for (long long j = 0; j < virtoffset; ++j)
        for (long long i = 0; i < virtoffset; ++i) H[i + j * hsize] = Hq(start + i, start + j);
      for (long long i = 0; i < extra; ++i) {
        const long long d = hsize - extra + i;
        H[d + d * hsize] = rpa_e[off + virtoffset + i];
      }
    }
  } else {
    const long long occ_extra = o.qpmin - o.vmin;
    for (long long i = 0; i < occ_extra; ++i) H[i + i * hsize] = rpa_e[off + i];
    // upstream copies the whole gwsize block (an Eigen assertion when cmax < qpmax); only the part of the QP window
    // inside the BSE window can be meant
    const long long cnt = std::min(gwsize, hsize - occ_extra);
    for (long long j = 0; j < cnt; ++j)
      for (long long i = 0; i < cnt; ++i) H[(occ_extra + i) + (occ_extra + j) * hsize] = Hq(i, j);
    if (o.cmax > o.qpmax) {
      const long long virtoffset = occ_extra + gwsize, extra = o.cmax - o.qpmax;
      for (long long i = 0; i < extra; ++i) {
        const long long d = hsize - extra + i;
        H[d + d * hsize] = rpa_e[off + virtoffset + i];
      }
    }
  }
  if (!o.use_Hqp_offdiag)
    for (long long j = 0; j < hsize; ++j)
      for (long long i = 0; i < hsize; ++i)
        if (i != j) H[i + j * hsize] = 0.0;
  hqp = std::move(H);
}

// BSE::SetupDirectInteractionOperator: eps(0) at the given energies -> eigenvectors U (returned, device) and
// eps_inv = 1/lambda (lambda > 1e-8).
std::unique_ptr<DBuf> bse_setup_screening(BSE& b, const double* rpa_e, double omega) {
  TCMatrix* tc = b.tc;
  Context* ctx = b.ctx;
  const long long na = tc->naux, rpatotal = tc->ntotal_glob;
  const long long n_occ = b.opt.homo - b.opt.rpamin + 1;
  // G0W0 after Sigma_PPM: the tensor is already in the eigenbasis of eps(0) at exactly these energies (upstream
  // recomputes eps(0) from the rotated tensor and diagonalises a matrix that is diag(lambda) up to rounding).  Read the
  // eigenvalues instead: no epsilon contraction, no N_aux^3 eigensolver, no window rotation (U = 1).
  // XTPB_BSE_REUSE_EPS0=0 forces the full recomputation.
  {
    const char* env = getenv("XTPB_BSE_REUSE_EPS0");
    const TCMatrix::Eps0Basis& z = tc->eps0;
    if (omega == 0.0 && !(env && env[0] == '0') && z.valid && !tc->pending && z.eta == kRpaEtaDefault && z.n_occ == n_occ &&
        (long long)z.energies.size() == rpatotal && std::equal(z.energies.begin(), z.energies.end(), rpa_e)) {
      b.eps_inv.resize((size_t)na);
      for (long long i = 0; i < na; ++i) b.eps_inv[i] = z.lambda[i] > 1e-8 ? 1.0 / z.lambda[i] : 0.0;
      b.eps0_reused = true;
      return nullptr;
    }
  }
  // the eigenvectors below are tied to the aux basis epsilon is formed in, and the windows are cut from the flushed
  // tensor: flush first, so that both see the same (reference, symmetric) metric factor
  tc->flush();
  DBuf e_dev((size_t)rpatotal), lam((size_t)na);
  ctx->h2d(e_dev.p, rpa_e, (size_t)rpatotal);
  auto U = std::make_unique<DBuf>((size_t)(na * na));
  const double w0 = omega;
  rpa_epsilon_dev(*tc, e_dev.p, n_occ, kRpaEtaDefault, &w0, 1, false, 0.0, U->p);
  ctx->eigh((int)na, U->p, na, lam.p);
  std::vector<double> lambda((size_t)na);
  ctx->d2h(lambda.data(), lam.p, (size_t)na);
  b.eps_inv.resize((size_t)na);
  for (long long i = 0; i < na; ++i) b.eps_inv[i] = lambda[i] > 1e-8 ? 1.0 / lambda[i] : 0.0;
  return U;
}

// BSE::Perturbative_DynamicalScreening (bse.cc): first-order correction of the BSE energies for the frequency dependence
// of the screening in the direct term.  With Hd(w) built from eps(w) (real axis) instead of eps(0),
//   E_dyn,s = E_static,s + <s|Hd^(w = E_dyn,s)|s> - <s|Hd^(0)|s>,       Hd^ = HdOperator = -Hd,
// iterated per state until |dE| < dyn_tolerance or max_dyn_iter.  Without the Tamm-Dancoff approximation the
// expectation value is X^T Hd^ X + Y^T Hd^ Y + 2 X^T Hd2^ Y (the energy functional X^T A X + Y^T A Y + 2 X^T B Y).
// Every evaluation is SetupDirectInteractionOperator at the new frequency (eps(w), eigensolver, window rotation)
// followed by one factorised operator application to the state's vector, as upstream does per excitation.
void bse_dynamical_screening(BSE& b, const double* R_static, const double* rpa_e, long long n_states,
                             const double* e_static, const double* X_host, const double* Y_host, long long ld,
                             long long max_iter, double tol, double* e_dyn, long long* iters) {
  Context* ctx = b.ctx;
  const long long size = b.size, lds = round_up(size, 2);
  XTPB_REQUIRE(n_states >= 1 && ld >= size && max_iter >= 1 && tol > 0.0, "bad dynamical-screening arguments");
  DBuf X((size_t)(lds * n_states)), Yv, HX((size_t)(lds * n_states)), HY, H2Y, dots((size_t)(3 * n_states));
  X.zero(ctx->stream);
  ctx->h2d_2d(X.p, lds, X_host, ld, size, n_states);
  if (Y_host) {
    Yv.alloc((size_t)(lds * n_states));
    HY.alloc((size_t)(lds * n_states));
    H2Y.alloc((size_t)(lds * n_states));
    Yv.zero(ctx->stream);
    ctx->h2d_2d(Yv.p, lds, Y_host, ld, size, n_states);
  }
  const std::vector<double> eps_static = b.eps_inv;
  // <s|Hd^|s> for states [s0, s0 + cnt) with the screening (eps_inv, R)
  auto expectation = [&](const std::vector<double>& eps_inv, const double* R, long long s0, long long cnt,
                         std::vector<double>& out) {
    BseOperator hd(ctx, b.tc, b.opt.homo, b.opt.rpamin, b.opt.vmin, b.opt.cmax, eps_inv.data(), b.hqp.data(),
                   b.vt + b.ct, 0, 0, 1, 0, R, /*force_factorised=*/true);
    hd.matmul_dev(X.p + s0 * lds, lds, (int)cnt, HX.p, lds);
    k_column_dots(dots.p, X.p + s0 * lds, lds, HX.p, lds, size, (int)cnt, ctx->stream);
    if (Y_host) {
      hd.matmul_dev(Yv.p + s0 * lds, lds, (int)cnt, HY.p, lds);
      k_column_dots(dots.p + n_states, Yv.p + s0 * lds, lds, HY.p, lds, size, (int)cnt, ctx->stream);
      BseOperator hd2(ctx, b.tc, b.opt.homo, b.opt.rpamin, b.opt.vmin, b.opt.cmax, eps_inv.data(), b.hqp.data(),
                      b.vt + b.ct, 0, 0, 0, 1, R, true);
      hd2.matmul_dev(Yv.p + s0 * lds, lds, (int)cnt, H2Y.p, lds);
      k_column_dots(dots.p + 2 * n_states, X.p + s0 * lds, lds, H2Y.p, lds, size, (int)cnt, ctx->stream);
    }
    std::vector<double> d((size_t)(3 * n_states), 0.0);
    ctx->d2h(d.data(), dots.p, (size_t)(3 * n_states));
    out.resize((size_t)cnt);
    for (long long i = 0; i < cnt; ++i)
      out[i] = d[i] + (Y_host ? d[n_states + i] + 2.0 * d[2 * n_states + i] : 0.0);
  };
  std::vector<double> stat;
  expectation(eps_static, R_static, 0, n_states, stat);
  for (long long s = 0; s < n_states; ++s) {
    double e = e_static[s];
    long long it = 0;
    for (; it < max_iter; ++it) {
      const double old = e;
      std::unique_ptr<DBuf> U = bse_setup_screening(b, rpa_e, old);      // overwrites b.eps_inv
      std::vector<double> dyn;
      expectation(b.eps_inv, U ? U->p : nullptr, s, 1, dyn);
      e = e_static[s] + dyn[0] - stat[s];
      if (std::fabs(e - old) < tol) { ++it; break; }
    }
    e_dyn[s] = e;
    if (iters) iters[s] = it;
  }
  b.eps_inv = eps_static;
  ctx->sync();
}

// ------------------------------------------------------------------ BSE_OPERATOR
BseOperator::BseOperator(Context* c, TCMatrix* tc, long long homo, long long rpamin, long long vmin, long long cmax,
                         const double* eps_inv_host, const double* hqp_host, long long ldh, int cqp_, int cx_, int cd_,
                         int cd2_, const double* R_dev, bool force_factorised)
    : cqp(cqp_), cx(cx_), cd(cd_), cd2(cd2_) {
  ctx = c;
  tc->flush();
  XTPB_REQUIRE(!(cd != 0 && cd2 != 0), "Hamiltonian cannot contain Hd and Hd2 at the same time");
  XTPB_REQUIRE(vmin >= rpamin && vmin <= homo && cmax > homo && cmax <= tc->mmax && cmax <= tc->nmax,
               "BSE window outside the TCMatrix");
  vt = homo - vmin + 1;
  ct = cmax - homo;
  size = vt * ct;
  // Multi-GPU: the operator is sharded over the auxiliary index (every term of H X is a sum over P); this rank
  // keeps the windows for P in [P0, P0 + naux) and matmul/diagonal all-reduce their partial results.
  const int world = ctx->world;
  const long long naux_glob = tc->naux;
  tc_naux_glob = naux_glob;
  const long long P0 = naux_glob * ctx->rank / world;
  naux = naux_glob * (ctx->rank + 1) / world - P0;
  XTPB_REQUIRE(naux > 0, "fewer auxiliary functions than ranks");
  const int v0 = (int)(vmin - rpamin), c0 = (int)(homo + 1 - rpamin);
  eps_inv_dev.alloc((size_t)naux_glob);
  ctx->h2d(eps_inv_dev.p, eps_inv_host, (size_t)naux_glob);
  const long long hs = vt + ct;
  hqp_dev.alloc((size_t)(hs * hs));
  ctx->h2d_2d(hqp_dev.p, hs, hqp_host, ldh, hs, hs);
  hqp_diag_dev.alloc((size_t)hs);
  k_extract_diagonal(hqp_dev.p, (int)hs, hs, hqp_diag_dev.p, ctx->stream);

  // rotation matrix with eps_inv folded into its columns, for the windows that carry the screening
  DBuf Rs;
  if (R_dev && (cd || cd2)) {
    Rs.alloc((size_t)(naux_glob * naux_glob));
    XTPB_CUDA(cudaMemcpyAsync(Rs.p, R_dev, (size_t)(naux_glob * naux_glob) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    k_scale_columns(Rs.p, (int)naux_glob, (int)naux_glob, naux_glob, eps_inv_dev.p, ctx->stream);
  }
  // dst[i*dst_slab + Ql*dst_ld + j] = (rotated / screened) M[m0+i][Pbeg+Ql][n0+j], Ql < pcnt, j < ncnt (pure strides:
  // [i][P][j] windows for the factorised operator, [P][i][j] "flat" operands for the dense build)
  auto window_into = [&](double* dst, long long dst_ld, long long dst_slab, int m0, int mcnt, int n0, int ncnt,
                         bool screened, long long Pbeg, long long pcnt) {
    if (world == 1) {
      XTPB_REQUIRE(Pbeg == 0 && pcnt == naux_glob, "single rank holds every aux function");
      if (R_dev)
        tc->rotate_window(dst, dst_ld, dst_slab, m0, mcnt, n0, ncnt, screened ? Rs.p : R_dev, naux_glob);
      else
        k_extract_window(dst, dst_ld, dst_slab, tc->M.p, tc->ldn, tc->slab, m0, mcnt, n0, ncnt, (int)naux_glob,
                         screened ? eps_inv_dev.p : nullptr, ctx->stream);
      return;
    }
    // re-shard: the tensor is distributed over the second index, the operator over P (or, dense, over columns).
    // Rotate / extract the local columns of the window for all P, all-gather the (small) window, keep what is asked.
    const long long jl0 = tc->nloc_below(n0), cl = tc->nloc_below(n0 + ncnt) - jl0;
    const long long ldl = round_up((ncnt + world - 1) / world + 1, 2), lslab = naux_glob * ldl;
    DBuf loc((size_t)(lslab * mcnt)), G((size_t)(lslab * mcnt * world));
    loc.zero(ctx->stream);
    if (cl > 0) {
      if (R_dev)
        tc->rotate_window(loc.p, ldl, lslab, m0, mcnt, (int)jl0, (int)cl, screened ? Rs.p : R_dev, naux_glob);
      else
        k_extract_window(loc.p, ldl, lslab, tc->M.p, tc->ldn, tc->slab, m0, mcnt, (int)jl0, (int)cl, (int)naux_glob,
                         screened ? eps_inv_dev.p : nullptr, ctx->stream);
    }
    ctx->allgather(loc.p, G.p, (size_t)(lslab * mcnt));
    k_window_from_gathered(dst, dst_ld, dst_slab, G.p, ldl, mcnt, (int)naux_glob, (int)Pbeg, (int)pcnt, n0, ncnt, world,
                           ctx->stream);
    ctx->sync();   // loc / G are freed on return
  };
  auto window = [&](DBuf& dst, long long& ld, long long& sl, int m0, int mcnt, int n0, int ncnt, bool screened) {
    ld = round_up(ncnt, 2);
    sl = naux * ld;
    dst.alloc((size_t)(sl * mcnt));
    dst.zero(ctx->stream);
    window_into(dst.p, ld, sl, m0, mcnt, n0, ncnt, screened, P0, naux);
  };

  // ---- dense mode: materialise H once (HBM is large: C60's 32 400^2 doubles are 8.4 GB) when it fits the budget.
  // The reference rebuilds H row blocks on every matmul (2 v^2 c^2 N_aux flops per call); the factorised products
  // below cost 2 N_aux k vc (v+c) per call; building H costs 2 v^2 c^2 N_aux (+ half of that for Hx) ONCE with
  // long-K (K = N_aux) contractions, after which a matmul only streams H (HBM-bound).
  {
    const char* env = getenv("XTPB_BSE_DENSE_MAX_GB");
    const double max_gb = env ? atof(env) : 32.0;
    v2lo = vt * ctx->rank / world;
    ns = vt * (ctx->rank + 1) / world - v2lo;
    h_ld = round_up(size, 2);
    const double gb = (double)h_ld * (double)(vt / world + 1) * (double)ct * 8e-9;
    // The exchange term has rank N_aux: Hx X = Mvc (Mvc^T X) costs 4 vc N_aux k flops per call (two HBM-bound passes
    // over the 1.4 GB operand at C60 size) while adding it to the dense H costs 2 (vc)^2 N_aux = 11.5 TFLOP once, as
    // much as Hd itself.  So only the screened direct term (and Hqp) is materialised; XTPB_BSE_HX_DENSE=1 restores the
    // fully dense H.
    const char* hx_env = getenv("XTPB_BSE_HX_DENSE");
    hx_factorised = cx != 0 && !(hx_env && hx_env[0] == '1');
    const bool has_dense_terms = hx_factorised ? (cd || cd2) : (cx || cd || cd2);
    // XTPB_BSE_MODE=factorised never materialises H (the strategy BASELINE.json's north_star describes; mandatory
    // when H does not fit); =dense insists on it when it fits; default: dense when it fits the budget
    const char* mode = getenv("XTPB_BSE_MODE");
    const bool want_factorised = force_factorised || (mode && mode[0] == 'f');
    dense = !want_factorised && has_dense_terms && gb <= max_gb && vt >= world && size <= 60000 &&
            (double)std::max(vt, ct) * (double)std::max(vt, ct) < 2.0e9;
    if (!dense) hx_factorised = false;        // the factorised operator below has its own exchange term
  }
  if (dense) {
    const long long ncols = ns * ct;
    H.alloc((size_t)(h_ld * std::max<long long>(ncols, 1)));
    H.zero(ctx->stream);
    ProfScope prof(PROF_BSE_MATMUL);
    auto flat = [&](DBuf& F, long long& ldF, int m0, int mcnt, int n0, int ncnt, bool screened) {
      ldF = round_up((long long)mcnt * ncnt, 2);
      F.alloc((size_t)(ldF * naux_glob));
      F.zero(ctx->stream);
      window_into(F.p, ldF, ncnt, m0, mcnt, n0, ncnt, screened, 0, naux_glob);     // [P][i][j]
    };
    DBuf Fa, Fb;
    long long ldA = 0, ldB = 0;
    if (cx || cd2) flat(Fvc, ldvcF, v0, (int)vt, c0, (int)ct, false);
    // NOTE (multi-GPU): every flat() is a collective re-shard, so all ranks must ask for the SAME window; the
    // column range a rank owns is then a contiguous row range of the flat operand (first index = v2), which is why
    // the symmetric partners M[v2][P][v1] = M[v1][P][v2] and M[v2][P][c1] = M[c1][P][v2] are used below.
    // single rank: Hd is symmetric under (v1,c1) <-> (v2,c2), so only the pairs v1 <= v2 are contracted (half of the
    // 2 (vc)^2 N_aux flops) into a scratch [(c2,c1)][pair] and scattered into H and its mirror image
    // (XTPB_BSE_SYM=0: contract every (v1, v2)); with several ranks every rank owns a v2 range and contracts all v1
    bool sym_done = false;
    if (cd && ncols > 0 && world == 1 && vt > 1) {
      const char* sym_env = getenv("XTPB_BSE_SYM");
      const long long npairs = vt * (vt + 1) / 2;
      size_t free_b = 0, total_b = 0;
      long long budget = 1LL << 28;                                        // doubles (2 GiB) when the query fails
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)                // else half of what is free, <= 8 GiB
        budget = std::min<long long>((long long)(free_b / 16), 1LL << 30);
      flat(Fa, ldA, c0, (int)ct, c0, (int)ct, false);                      // rows (i = c2, j = c1)
      flat(Fb, ldB, v0, (int)vt, v0, (int)vt, true);                       // rows (i = v2, j = v1)
      long long tchunk = std::min<long long>(npairs, budget / (ct * ct));
      if (tchunk < npairs) tchunk = tchunk / 128 * 128;                    // whole tile columns, 16-byte aligned chunks
      if (!(sym_env && sym_env[0] == '0') && tchunk >= std::min<long long>(npairs, 128)) {
        const long long ldT = round_up(npairs, 2);
        DBuf Ftri((size_t)(ldT * naux_glob)), Tmp((size_t)(tchunk * ct * ct));
        Ftri.zero(ctx->stream);
        k_bse_pack_pairs(Ftri.p, ldT, Fb.p, ldB, (int)vt, (int)naux_glob, ctx->stream);
        for (long long t0 = 0; t0 < npairs; t0 += tchunk) {
          const long long tc_ = std::min(tchunk, npairs - t0);
          GemmParams g{};
          g.A = GemmOperand{Fa.p, 1, ldA, 0, 0};
          g.B = GemmOperand{Ftri.p + t0, 1, ldT, 0, 0};
          g.C = Tmp.p; g.c_sm = 1; g.c_sn = ct * ct;
          g.M = (int)(ct * ct); g.N = (int)tc_; g.K = (int)naux_glob; g.n_outer = 1; g.n_batch = 1;
          g.alpha = -(double)cd; g.beta = 0.0;
          contract(g, ctx->ws, ctx->stream);
          k_bse_scatter_pairs(H.p, h_ld, (int)ct, Tmp.p, t0, tc_, ctx->stream);
        }
        ctx->sync();
        sym_done = true;
        Fa.release();
        Fb.release();
      }
    }
    if (cd && ncols > 0 && !sym_done) {
      // Hd[(v1,c1),(v2,c2)] = sum_P Mcc[c2][P][c1] * (eps_inv_P Mvv[v2][P][v1])
      if (!Fa.p) flat(Fa, ldA, c0, (int)ct, c0, (int)ct, false);           // rows (i = c2, j = c1)
      if (!Fb.p) flat(Fb, ldB, v0, (int)vt, v0, (int)vt, true);            // rows (i = v2, j = v1), all v2
      GemmParams g{};
      g.A = GemmOperand{Fa.p, 1, ldA, 0, 0};
      g.B = GemmOperand{Fb.p + v2lo * vt, 1, ldB, 0, 0};                   // this rank's v2 range
      g.C = H.p;
      g.c_m_inner = (int)ct; g.c_sm = 1; g.c_sm_outer = h_ld;              // row (c2, c1): c1 -> row, c2 -> column part
      g.c_n_inner = (int)vt; g.c_sn = ct; g.c_sn_outer = ct * h_ld;        // col (v2l, v1): v1 -> row part, v2l -> column
      g.M = (int)(ct * ct); g.N = (int)(ns * vt); g.K = (int)naux_glob; g.n_outer = 1; g.n_batch = 1;
      g.alpha = -(double)cd; g.beta = 1.0;
      contract(g, ctx->ws, ctx->stream);
      ctx->sync();
      Fa.release();
      Fb.release();
    }
    if (cd2 && ncols > 0) {
      // Hd2[(v1,c1),(v2,c2)] = sum_P Mvc[v1][P][c2] * (eps_inv_P Mvc[v2][P][c1])
      flat(Fb, ldB, v0, (int)vt, c0, (int)ct, true);                       // rows (i = v2, j = c1), all v2, screened
      GemmParams g{};
      g.A = GemmOperand{Fvc.p, 1, ldvcF, 0, 0};                            // rows (v1, c2)
      g.B = GemmOperand{Fb.p + v2lo * ct, 1, ldB, 0, 0};
      g.C = H.p;
      g.c_m_inner = (int)ct; g.c_sm = h_ld; g.c_sm_outer = ct;             // c2 -> column part, v1 -> row part
      g.c_n_inner = (int)ct; g.c_sn = 1; g.c_sn_outer = ct * h_ld;         // c1 -> row part, v2l -> column
      g.M = (int)(vt * ct); g.N = (int)(ns * ct); g.K = (int)naux_glob; g.n_outer = 1; g.n_batch = 1;
      g.alpha = -(double)cd2; g.beta = 1.0;
      contract(g, ctx->ws, ctx->stream);
      ctx->sync();
      Fb.release();
    }
    if (cx && !hx_factorised && ncols > 0) {
      // Hx[(v1,c1),(v2,c2)] = sum_P Mvc[v1][P][c1] Mvc[v2][P][c2]
      GemmParams g{};
      g.A = GemmOperand{Fvc.p, 1, ldvcF, 0, 0};
      g.B = GemmOperand{Fvc.p + v2lo * ct, 1, ldvcF, 0, 0};
      g.C = H.p; g.c_sm = 1; g.c_sn = h_ld;
      g.M = (int)size; g.N = (int)ncols; g.K = (int)naux_glob; g.n_outer = 1; g.n_batch = 1;
      g.alpha = (double)cx; g.beta = 1.0;
      contract(g, ctx->ws, ctx->stream);
    }
    if (cqp && ncols > 0)
      k_bse_add_hqp(H.p, h_ld, (int)vt, (int)ct, (int)v2lo, (int)ns, hqp_dev.p, hs, (double)cqp, ctx->stream);
    ctx->sync();
    if (!hx_factorised) Fvc.release();         // the flat [P][(v,c)] operand stays for the exchange products
    ldvc = slabvc = ldvv = slabvv = ldcc = slabcc = ldcv = slabcv = 0;
    return;
  }
  ldvc = slabvc = ldvv = slabvv = ldcc = slabcc = ldcv = slabcv = 0;
  if (cx || cd2) window(Mvc, ldvc, slabvc, v0, (int)vt, c0, (int)ct, false);
  if (cd) {
    window(Mvv, ldvv, slabvv, v0, (int)vt, v0, (int)vt, true);
    window(Mcc, ldcc, slabcc, c0, (int)ct, c0, (int)ct, false);
  }
  if (cd2) window(Mcv, ldcv, slabcv, c0, (int)ct, v0, (int)vt, true);
  ctx->sync();
}

void BseOperator::diagonal_dev(double* d) {
  if (dense) {      // the owned columns' diagonal entries, zero elsewhere; summed over ranks
    XTPB_CUDA(cudaMemsetAsync(d, 0, (size_t)size * 8, ctx->stream));
    if (ns > 0) k_extract_diagonal(H.p + v2lo * ct, (int)(ns * ct), h_ld, d + v2lo * ct, ctx->stream);
    if (hx_factorised && ns > 0)               // + cx sum_P Mvc[v][P][c]^2 for the owned entries
      k_add_column_square_sums(d + v2lo * ct, Fvc.p + v2lo * ct, ldvcF, (int)tc_naux_glob, ns * ct, (double)cx,
                               ctx->stream);
    ctx->allreduce_sum(d, (size_t)size);
    return;
  }
  // partial sums over the local aux range; the P-independent Hqp part is contributed by rank 0 only
  k_bse_diagonal(d, (int)vt, (int)ct, (int)naux, Mvc.p, ldvc, slabvc, Mvv.p, ldvv, slabvv, Mcc.p, ldcc, slabcc, Mcv.p,
                 ldcv, slabcv, eps_inv_dev.p, hqp_diag_dev.p, ctx->rank == 0 ? cqp : 0, cx, cd, cd2, ctx->stream);
  ctx->allreduce_sum(d, (size_t)size);
}

void BseOperator::matmul_dev(const double* X, long long ldx, int k, double* Y, long long ldy) {
  XTPB_REQUIRE(k > 0, "matmul needs at least one column");
  ProfScope prof(PROF_BSE_MATMUL);
  cudaStream_t st = ctx->stream;
  bool first = true;
  auto beta = [&]() { const double b = first ? 0.0 : 1.0; first = false; return b; };
  const long long hs = vt + ct;

  XTPB_REQUIRE(k <= 8192, "more than 8192 trial vectors per matmul are not supported");

  if (dense) {      // Y = H[:, owned columns] X[owned rows, :]  (one HBM-bound pass over H), summed over ranks
    if (ns > 0) {
      GemmParams g{};
      g.A = GemmOperand{H.p, 1, h_ld, 0, 0};
      g.B = GemmOperand{X + v2lo * ct, ldx, 1, 0, 0};
      g.C = Y; g.c_sm = 1; g.c_sn = ldy;
      g.M = (int)size; g.N = k; g.K = (int)(ns * ct); g.n_outer = 1; g.n_batch = 1; g.alpha = 1.0; g.beta = 0.0;
      contract(g, ctx->ws, st);
    } else {
      XTPB_CUDA(cudaMemset2DAsync(Y, ldy * 8, 0, size * 8, k, st));
    }
    if (hx_factorised && ns > 0) {
      // exchange, factorised: T(P,kk) = sum_{i owned} Mvc(i,P) X(i,kk);  Y(i,kk) += cx sum_P Mvc(i,P) T(P,kk).
      // Linear in T, so the partial T of each rank goes straight into its partial Y (summed by the all-reduce below).
      const long long na = tc_naux_glob;
      T.ensure((size_t)(na * k));
      GemmParams g{};
      g.A = GemmOperand{Fvc.p + v2lo * ct, ldvcF, 1, 0, 0};
      g.B = GemmOperand{X + v2lo * ct, ldx, 1, 0, 0};
      g.C = T.p; g.c_sm = 1; g.c_sn = na;
      g.M = (int)na; g.N = k; g.K = (int)(ns * ct); g.n_outer = 1; g.n_batch = 1; g.alpha = 1.0; g.beta = 0.0;
      contract(g, ctx->ws, st);
      GemmParams h{};
      h.A = GemmOperand{Fvc.p, 1, ldvcF, 0, 0};
      h.B = GemmOperand{T.p, na, 1, 0, 0};
      h.C = Y; h.c_sm = 1; h.c_sn = ldy;
      h.M = (int)size; h.N = k; h.K = (int)na; h.n_outer = 1; h.n_batch = 1; h.alpha = (double)cx; h.beta = 1.0;
      contract(h, ctx->ws, st);
    }
    ctx->allreduce_sum(Y, (size_t)(ldy * (k - 1) + size));
    return;
  }

  if (cqp && ctx->rank == 0) {   // P-independent term: one rank contributes it, the all-reduce below spreads it
    // Y_k(c,v) = cqp * sum_c2 Hc(c,c2) X_k(c2,v)          (Hqp symmetric: K-contiguous view of the cc block)
    GemmParams g{};
    g.A = GemmOperand{hqp_dev.p + vt + vt * hs, hs, 1, 0, 0};
    g.B = GemmOperand{X, ct, 1, 0, ldx};
    g.C = Y; g.c_sm = 1; g.c_sn = ct; g.c_batch = ldy;
    g.M = (int)ct; g.N = (int)vt; g.K = (int)ct; g.n_outer = 1; g.n_batch = k;
    g.alpha = cqp; g.beta = beta();
    contract(g, ctx->ws, st);
    // Y_k(c,v) -= cqp * sum_v2 X_k(c,v2) Hv(v2,v)
    GemmParams h{};
    h.A = GemmOperand{X, 1, ct, 0, ldx};
    h.B = GemmOperand{hqp_dev.p, hs, 1, 0, 0};
    h.C = Y; h.c_sm = 1; h.c_sn = ct; h.c_batch = ldy;
    h.M = (int)ct; h.N = (int)vt; h.K = (int)vt; h.n_outer = 1; h.n_batch = k;
    h.alpha = -(double)cqp; h.beta = 1.0;
    contract(h, ctx->ws, st);
  }

  if (cx) {
    T.ensure((size_t)(naux * k));
    // T(P,k) = sum_v sum_c Mvc[v][P][c] X[(v,c),k]
    GemmParams g{};
    g.A = GemmOperand{Mvc.p, ldvc, 1, slabvc, 0};
    g.B = GemmOperand{X, ldx, 1, ct, 0};
    g.C = T.p; g.c_sm = 1; g.c_sn = naux;
    g.M = (int)naux; g.N = k; g.K = (int)ct; g.n_outer = (int)vt; g.n_batch = 1;
    g.alpha = 1.0; g.beta = 0.0;
    contract(g, ctx->ws, st);
    // Y[(v,c),k] += cx sum_P Mvc[v][P][c] T(P,k)            (batched over v)
    GemmParams h{};
    h.A = GemmOperand{Mvc.p, 1, ldvc, 0, slabvc};
    h.B = GemmOperand{T.p, naux, 1, 0, 0};
    h.C = Y; h.c_sm = 1; h.c_sn = ldy; h.c_batch = ct;
    h.M = (int)ct; h.N = k; h.K = (int)naux; h.n_outer = 1; h.n_batch = (int)vt;
    h.alpha = cx; h.beta = beta();
    contract(h, ctx->ws, st);
  }

  if (cd || cd2) {
    // packed copy of X when its leading dimension is not the operator size (column index (k,v2) must be affine)
    const double* Xp = X;
    DBuf Xpack;
    if (ldx != size) {
      Xpack.alloc((size_t)(size * k));
      k_copy_2d(Xpack.p, size, X, ldx, (int)size, k, st);
      Xp = Xpack.p;
    }
    // direct term (cd):   rows of step 1 are (c1,P) from Mcc, screened factor Mvv, output index (v1,c1)
    // direct term (cd2):  rows of step 1 are (v1,P) from Mvc, screened factor Mcv, output index (v1,c1) as well
    const long long r1 = cd ? ct : vt;                 // first index count of the step-1 tensor
    const double* A1 = cd ? Mcc.p : Mvc.p;
    const long long ld1 = cd ? ldcc : ldvc;
    const double* A2 = cd ? Mvv.p : Mcv.p;
    const long long ld2 = cd ? ldvv : ldcv, slab2 = cd ? slabvv : slabcv;
    const long long n2 = cd ? vt : ct;                 // output index delivered by the screened factor
    const long long ldu = round_up(vt, 2);
    const long long per_k = r1 * naux * ldu;
    // U[kl][r][P][v2] holds one trial-vector chunk of the half-contracted direct term.  The more vectors per chunk, the
    // wider the two contractions below (their N resp. M dimension is kc * v) and the less tile padding they carry, so
    // U may take up to a third of the free device memory, at most 24 GiB (XTPB_BSE_U_MAX_GB), at least 4 GiB.
    long long budget = 1LL << 29;                      // doubles (4 GiB)
    {
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        const char* env = getenv("XTPB_BSE_U_MAX_GB");
        const double cap_gb = env ? atof(env) : 24.0;
        const double want = std::min((double)free_b / 3.0 + (double)U.n * 8.0, cap_gb * 1073741824.0);
        budget = std::max<long long>(budget, (long long)(want / 8.0));
      }
    }
    const int kmax = (int)std::max<long long>(1, std::min<long long>(k, budget / per_k));
    const int kchunks = (k + kmax - 1) / kmax;
    const int kchunk = (k + kchunks - 1) / kchunks;        // equal chunks: no short (badly padded) last one
    U.ensure((size_t)(per_k * kchunk));
    const double coef = cd ? (double)cd : (double)cd2;
    const double b0 = beta();
    for (int k0 = 0; k0 < k; k0 += kchunk) {
      const int kc = std::min(kchunk, k - k0);
      // step 1: U[kl][r][P][v2] = sum_c2 A1[r][P][c2] X[k0+kl][v2*ct + c2]
      GemmParams g{};
      g.A = GemmOperand{A1, ld1, 1, 0, 0};
      g.B = GemmOperand{Xp + (long long)k0 * size, ct, 1, 0, 0};
      g.C = U.p; g.c_sm = ldu; g.c_sn = 1; g.c_n_inner = (int)vt; g.c_sn_outer = per_k;
      g.M = (int)(r1 * naux); g.N = (int)(kc * vt); g.K = (int)ct; g.n_outer = 1; g.n_batch = 1;
      g.alpha = 1.0; g.beta = 0.0;
      contract(g, ctx->ws, st);
      // step 2: Y[k0+kl][v1,c1] -= coef * sum_{P,v2} A2[n][P][v2] U[kl][r][P][v2]
      //   rows = (kl, r) from U, columns = n from the screened factor
      GemmParams h{};
      h.A = GemmOperand{U.p, naux * ldu, 1, ldu, 0};
      h.B = GemmOperand{A2, slab2, 1, ld2, 0};
      h.C = Y + (long long)k0 * ldy;
      if (cd) {   // row (kl,c1) -> kl*ldy + c1 ; col v1 -> v1*ct
        h.c_sm = 1; h.c_m_inner = (int)ct; h.c_sm_outer = ldy; h.c_sn = ct;
      } else {    // row (kl,v1) -> kl*ldy + v1*ct ; col c1 -> c1
        h.c_sm = ct; h.c_m_inner = (int)vt; h.c_sm_outer = ldy; h.c_sn = 1;
      }
      h.M = (int)(kc * r1); h.N = (int)n2; h.K = (int)vt; h.n_outer = (int)naux; h.n_batch = 1;
      h.alpha = -coef; h.beta = b0;
      contract(h, ctx->ws, st);
    }
  }
  if (first) XTPB_CUDA(cudaMemset2DAsync(Y, ldy * 8, 0, size * 8, k, st));   // all coefficients zero
  // one all-reduce of Y per matmul (C60, k = 15: 3.9 MB); Y columns are ldy apart, padding rows ride along
  ctx->allreduce_sum(Y, (size_t)(ldy * (k - 1) + size));
}

// ------------------------------------------------------------------ dense operator (Davidson tests)
DenseOperator::DenseOperator(Context* c, const double* A_host, long long n, long long lda_host) {
  ctx = c;
  size = n;
  lda = round_up(n, 2);
  A.alloc((size_t)(lda * n));
  A.zero(ctx->stream);
  ctx->h2d_2d(A.p, lda, A_host, lda_host, n, n);
  ctx->sync();
}
void DenseOperator::matmul_dev(const double* X, long long ldx, int k, double* Y, long long ldy) {
  GemmParams g{};
  g.A = GemmOperand{A.p, 1, lda, 0, 0};
  g.B = GemmOperand{X, ldx, 1, 0, 0};
  g.C = Y; g.c_sm = 1; g.c_sn = ldy;
  g.M = (int)size; g.N = k; g.K = (int)size; g.n_outer = 1; g.n_batch = 1;
  g.alpha = 1.0; g.beta = 0.0;
  contract(g, ctx->ws, ctx->stream);
}
void DenseOperator::diagonal_dev(double* d) { k_extract_diagonal(A.p, (int)size, lda, d, ctx->stream); }

// ------------------------------------------------------------------ DavidsonSolver::solve (symmetric)
namespace {

struct DavidsonWork {
  Context* ctx;
  long long n, ld, cap;
  DBuf V, AV, Q, R, small, vec, tmp;   // Q: Ritz vectors, R: residuals, tmp: block being re-orthonormalised
};

// C(s x k) = V[:, :s]^T W[:, :k]
void gemm_tn(Context* ctx, const double* V, long long ldv, int s, const double* W, long long ldw, int k, long long n,
             double* C, long long ldc) {
  GemmParams g{};
  g.A = GemmOperand{V, ldv, 1, 0, 0};
  g.B = GemmOperand{W, ldw, 1, 0, 0};
  g.C = C; g.c_sm = 1; g.c_sn = ldc;
  g.M = s; g.N = k; g.K = (int)n; g.n_outer = 1; g.n_batch = 1; g.alpha = 1.0; g.beta = 0.0;
  contract(g, ctx->ws, ctx->stream);
}
// C(n x k) = alpha * V[:, :s] * S(s x k) + beta * C
void gemm_nn(Context* ctx, const double* V, long long ldv, int s, const double* S, long long lds, int k, long long n,
             double* C, long long ldc, double alpha, double beta) {
  GemmParams g{};
  g.A = GemmOperand{V, 1, ldv, 0, 0};
  g.B = GemmOperand{S, lds, 1, 0, 0};
  g.C = C; g.c_sm = 1; g.c_sn = ldc;
  g.M = (int)n; g.N = k; g.K = s; g.n_outer = 1; g.n_batch = 1; g.alpha = alpha; g.beta = beta;
  contract(g, ctx->ws, ctx->stream);
}

// two-pass Gram-Schmidt of columns nstart..ncols-1 of V against all kept previous ones; dependent columns are
// dropped (kept columns are compacted).  Returns the new column count.
int gram_schmidt_columns(DavidsonWork& w, int nstart, int ncols) {
  Context* ctx = w.ctx;
  int kept = nstart;
  double* coef = w.small.p;       // up to cap doubles
  double* nrm_dev = w.small.p + w.cap;
  for (int j = nstart; j < ncols; ++j) {
    double* col = w.V.p + (long long)j * w.ld;
    if (kept > 0) {
      for (int pass = 0; pass < 2; ++pass) {
        gemm_tn(ctx, w.V.p, w.ld, kept, col, w.ld, 1, w.n, coef, w.cap);
        gemm_nn(ctx, w.V.p, w.ld, kept, coef, w.cap, 1, w.n, col, w.ld, -1.0, 1.0);
      }
    }
    k_col_norms(col, w.ld, w.n, 1, nrm_dev, ctx->stream);
    double nrm = 0.0;
    ctx->d2h(&nrm, nrm_dev, 1);
    if (!(nrm > 1e-10)) continue;
    double* dst = w.V.p + (long long)kept * w.ld;
    if (dst != col) XTPB_CUDA(cudaMemcpyAsync(dst, col, (size_t)w.n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    k_scale(dst, w.n, 1.0 / nrm, ctx->stream);
    ++kept;
  }
  return kept;
}


// Block version used on the hot path: two classical Gram-Schmidt passes of the whole new block against the kept
// basis (2 GEMMs each), then Cholesky-QR (twice) inside the block -- a handful of launches and one small D2H per
// call instead of ~8 launches and a host sync per column.  A (near-)dependent block (tiny Cholesky pivot) falls
// back to the column-by-column routine above, which drops dependent columns.
int gram_schmidt(DavidsonWork& w, int nstart, int ncols) {
  Context* ctx = w.ctx;
  const int k = ncols - nstart;
  if (k <= 0) return nstart;
  if (k == 1 || k > 64) return gram_schmidt_columns(w, nstart, ncols);
  double* W = w.V.p + (long long)nstart * w.ld;
  double* coef = w.small.p;                      // nstart x k (<= cap*cap)
  if (nstart > 0) {
    for (int pass = 0; pass < 2; ++pass) {
      gemm_tn(ctx, w.V.p, w.ld, nstart, W, w.ld, k, w.n, coef, nstart);
      gemm_nn(ctx, w.V.p, w.ld, nstart, coef, nstart, k, w.n, W, w.ld, -1.0, 1.0);
    }
  }
  std::vector<double> G((size_t)k * k), S((size_t)k * k);
  for (int pass = 0; pass < 2; ++pass) {
    gemm_tn(ctx, W, w.ld, k, W, w.ld, k, w.n, coef, k);
    ctx->d2h(G.data(), coef, (size_t)k * k);
    // Cholesky G = L L^T (lower), then S = L^{-T} (upper) so that W S is orthonormal
    std::vector<double> L((size_t)k * k, 0.0);
    double dmax = 0.0;
    for (int j = 0; j < k; ++j) dmax = std::max(dmax, G[j + (size_t)j * k]);
    bool ok = dmax > 0.0 && std::isfinite(dmax);
    for (int j = 0; j < k && ok; ++j) {
      double d = G[j + (size_t)j * k];
      for (int t = 0; t < j; ++t) d -= L[j + (size_t)t * k] * L[j + (size_t)t * k];
      if (!(d > 1e-12 * G[j + (size_t)j * k]) || !(G[j + (size_t)j * k] > 1e-20)) { ok = false; break; }
      const double ljj = std::sqrt(d);
      L[j + (size_t)j * k] = ljj;
      for (int i = j + 1; i < k; ++i) {
        double v = G[i + (size_t)j * k];
        for (int t = 0; t < j; ++t) v -= L[i + (size_t)t * k] * L[j + (size_t)t * k];
        L[i + (size_t)j * k] = v / ljj;
      }
    }
    if (!ok) return gram_schmidt_columns(w, nstart, ncols);
    // S = (L^{-1})^T: solve L X = I column by column (X lower), S(i,j) = X(j,i)
    std::fill(S.begin(), S.end(), 0.0);
    for (int c = 0; c < k; ++c) {
      std::vector<double> x((size_t)k, 0.0);
      for (int i = c; i < k; ++i) {
        double v = i == c ? 1.0 : 0.0;
        for (int t = c; t < i; ++t) v -= L[i + (size_t)t * k] * x[t];
        x[i] = v / L[i + (size_t)i * k];
      }
      for (int i = c; i < k; ++i) S[c + (size_t)i * k] = x[i];      // S(c, i) = X(i, c)
    }
    ctx->h2d(coef, S.data(), (size_t)k * k);
    gemm_nn(ctx, W, w.ld, k, coef, k, k, w.n, w.tmp.p, w.ld, 1.0, 0.0);
    k_copy_2d(W, w.ld, w.tmp.p, w.ld, (int)w.n, k, ctx->stream);
  }
  return ncols;
}

}  // namespace

void davidson_solve(Operator& A, long long neigen, const xtpb_davidson_options& opt, DavidsonResult& out) {
  Context* ctx = A.ctx;
  ProfScope prof(PROF_DAVIDSON);
  const long long n = A.size;
  XTPB_REQUIRE(neigen >= 1 && neigen <= n, "neigen out of range");
  long long max_space = opt.max_search_space;
  if (max_space < neigen) max_space = neigen * 5;
  if (max_space >= n) max_space = n;                    // DavidsonSolver::checkOptions clamps to the operator size
  long long guess = opt.size_initial_guess == 0 ? 2 * neigen : opt.size_initial_guess;
  guess = std::min(guess, n);
  long long su;
  switch (opt.size_update) {
    case XTPB_UPDATE_MIN: su = neigen; break;
    case XTPB_UPDATE_MAX: su = 2 * neigen; break;
    default: su = neigen < 20 ? (long long)(1.5 * (double)neigen) : neigen + 10; break;
  }
  su = std::min(su, guess);

  DavidsonWork w;
  w.ctx = ctx;
  w.n = n;
  w.ld = round_up(n, 2);
  w.cap = std::min<long long>(n, max_space + su) + guess + 2;
  w.V.alloc((size_t)(w.ld * w.cap));
  w.AV.alloc((size_t)(w.ld * w.cap));
  w.Q.alloc((size_t)(w.ld * su));
  w.R.alloc((size_t)(w.ld * su));
  w.tmp.alloc((size_t)(w.ld * std::max<long long>(su, 1)));
  w.small.alloc((size_t)(w.cap * (w.cap + 4) + 16));
  w.vec.alloc((size_t)(2 * n + 1024));
  w.V.zero(ctx->stream);

  // diagonal, initial guess = unit vectors on the smallest diagonal entries
  DBuf D((size_t)n);
  A.diagonal_dev(D.p);
  std::vector<double> Dh((size_t)n);
  ctx->d2h(Dh.data(), D.p, (size_t)n);
  std::vector<long long> order((size_t)n);
  std::iota(order.begin(), order.end(), 0LL);
  std::stable_sort(order.begin(), order.end(), [&](long long a, long long b) { return Dh[a] < Dh[b]; });
  {
    DBuf idx((size_t)guess);   // reinterpret as int64
    XTPB_CUDA(cudaMemcpyAsync(idx.p, order.data(), (size_t)guess * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_unit_vectors(w.V.p, w.ld, n, reinterpret_cast<const long long*>(idx.p), (int)guess, ctx->stream);
    ctx->sync();
  }

  int ncols = (int)guess, nold = 0;
  std::vector<double> T;            // projected matrix, ncols x ncols col-major (ld = ncols)
  std::vector<double> lam, Uh;
  bool have_ritz = false;
  DBuf Tdev, Udev((size_t)(w.cap * w.cap)), lamdev((size_t)w.cap), nrmdev((size_t)su + 1);
  out.info = 1;
  out.iterations = 0;
  std::vector<double> rn((size_t)su);

  // XTPB_TRACE=1: host seconds per phase of the iteration (with a stream synchronisation at every phase boundary, so
  // device time lands in the phase that issued it) on stderr -- a diagnostic, never on during a bench
  static const bool trace = [] { const char* e = getenv("XTPB_TRACE"); return e && e[0] == '1'; }();
  enum { PH_MATMUL, PH_PROJECT, PH_EIGH, PH_RITZ, PH_CORRECT, PH_ORTHO, PH_N };
  double phase[PH_N] = {};
  auto tnow = [] { return std::chrono::steady_clock::now(); };
  auto lap = [&](int ph, std::chrono::steady_clock::time_point& t0) {
    if (!trace) return;
    ctx->sync();
    const auto t1 = tnow();
    phase[ph] += std::chrono::duration<double>(t1 - t0).count();
    t0 = t1;
  };
  for (long long it = 0; it < opt.iter_max; ++it) {
    out.iterations = it + 1;
    auto t0 = tnow();
    if (ncols > max_space && have_ritz) {
      // restart: V <- Ritz vectors, AV <- AV U, T <- V^T AV
      gemm_nn(ctx, w.AV.p, w.ld, nold, Udev.p, nold, (int)su, n, w.R.p, w.ld, 1.0, 0.0);   // R = AV U (temp)
      XTPB_CUDA(cudaMemcpyAsync(w.V.p, w.Q.p, (size_t)(w.ld * su) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      XTPB_CUDA(cudaMemcpyAsync(w.AV.p, w.R.p, (size_t)(w.ld * su) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      ncols = gram_schmidt(w, 0, (int)su);
      T.assign((size_t)ncols * ncols, 0.0);
      gemm_tn(ctx, w.V.p, w.ld, ncols, w.AV.p, w.ld, ncols, n, w.small.p, ncols);
      ctx->d2h(T.data(), w.small.p, (size_t)ncols * ncols);
      nold = ncols;
    } else {
      // AV[:, nold:] = A V[:, nold:];  T[:, nold:] = V^T AV[:, nold:]
      const int nnew = ncols - nold;
      A.matmul_dev(w.V.p + (long long)nold * w.ld, w.ld, nnew, w.AV.p + (long long)nold * w.ld, w.ld);
      lap(PH_MATMUL, t0);
      gemm_tn(ctx, w.V.p, w.ld, ncols, w.AV.p + (long long)nold * w.ld, w.ld, nnew, n, w.small.p, ncols);
      std::vector<double> Tn((size_t)ncols * nnew);
      ctx->d2h(Tn.data(), w.small.p, Tn.size());
      std::vector<double> T2((size_t)ncols * ncols, 0.0);
      for (int j = 0; j < nold; ++j)
        for (int i = 0; i < nold; ++i) T2[i + (size_t)j * ncols] = T[i + (size_t)j * nold];
      for (int j = 0; j < nnew; ++j)
        for (int i = 0; i < ncols; ++i) {
          T2[i + (size_t)(nold + j) * ncols] = Tn[i + (size_t)j * ncols];
          if (i < nold) T2[(nold + j) + (size_t)i * ncols] = Tn[i + (size_t)j * ncols];
        }
      T.swap(T2);
      nold = ncols;
    }
    lap(PH_PROJECT, t0);
    // Ritz pairs of the symmetrised projected matrix
    {
      std::vector<double> Ts((size_t)ncols * ncols);
      for (int j = 0; j < ncols; ++j)
        for (int i = 0; i < ncols; ++i) Ts[i + (size_t)j * ncols] = 0.5 * (T[i + (size_t)j * ncols] + T[j + (size_t)i * ncols]);
      lam.resize((size_t)ncols);
      // the projected problem (<= max_search_space rows) is solved on the host, as upstream does with
      // Eigen::SelfAdjointEigenSolver: no cuSOLVER launch chain, no extra synchronisation
      if (ncols <= kHostEighMax && host_eigh(ncols, Ts.data(), lam.data())) {
        ctx->h2d(Udev.p, Ts.data(), Ts.size());
        ctx->h2d(lamdev.p, lam.data(), (size_t)ncols);
      } else {
        for (int j = 0; j < ncols; ++j)
          for (int i = 0; i < ncols; ++i) Ts[i + (size_t)j * ncols] = 0.5 * (T[i + (size_t)j * ncols] + T[j + (size_t)i * ncols]);
        ctx->h2d(Udev.p, Ts.data(), Ts.size());
        ctx->eigh(ncols, Udev.p, ncols, lamdev.p);
        ctx->d2h(lam.data(), lamdev.p, (size_t)ncols);
      }
    }
    lap(PH_EIGH, t0);
    const int nsu = (int)std::min<long long>(su, ncols);
    gemm_nn(ctx, w.V.p, w.ld, ncols, Udev.p, ncols, nsu, n, w.Q.p, w.ld, 1.0, 0.0);     // q = V U
    gemm_nn(ctx, w.AV.p, w.ld, ncols, Udev.p, ncols, nsu, n, w.R.p, w.ld, 1.0, 0.0);    // r = AV U
    k_residuals(w.R.p, w.ld, w.Q.p, w.ld, lamdev.p, n, nsu, ctx->stream);               //   - q lambda
    k_col_norms(w.R.p, w.ld, n, nsu, nrmdev.p, ctx->stream);
    ctx->d2h(rn.data(), nrmdev.p, (size_t)nsu);
    lap(PH_RITZ, t0);
    have_ritz = true;
    bool converged = true;
    for (long long j = 0; j < neigen; ++j) converged = converged && (j < nsu) && rn[j] < opt.tolerance;
    if (converged) {
      out.info = 0;
      break;
    }
    if (it == opt.iter_max - 1) break;
    // correction vectors for the unconverged roots
    int added = 0;
    for (int j = 0; j < nsu; ++j) {
      if (rn[j] < opt.tolerance) continue;
      XTPB_REQUIRE(ncols + added < w.cap, "Davidson search space overflow");
      double* dst = w.V.p + (long long)(ncols + added) * w.ld;
      k_davidson_correction(dst, w.R.p + (long long)j * w.ld, w.Q.p + (long long)j * w.ld, D.p, lam[j], n,
                            opt.correction == XTPB_DAVIDSON_OLSEN ? 1 : 0, w.vec.p, ctx->stream);
      ++added;
    }
    lap(PH_CORRECT, t0);
    const int before = ncols;
    ncols = gram_schmidt(w, before, before + added);
    lap(PH_ORTHO, t0);
    if (ncols == before) break;    // nothing independent left to add
  }
  if (trace)
    fprintf(stderr, "[xtpb trace] davidson n=%lld neigen=%lld iterations=%lld: matmul %.4f project %.4f eigh %.4f "
            "ritz+residual %.4f correction %.4f orthogonalise %.4f s\n", n, neigen, out.iterations, phase[PH_MATMUL],
            phase[PH_PROJECT], phase[PH_EIGH], phase[PH_RITZ], phase[PH_CORRECT], phase[PH_ORTHO]);
  out.evals.assign(lam.begin(), lam.begin() + std::min<size_t>((size_t)neigen, lam.size()));
  out.evecs.alloc((size_t)(n * neigen));
  k_copy_2d(out.evecs.p, n, w.Q.p, w.ld, (int)n, neigen, ctx->stream);
  ctx->sync();
}


// ------------------------------------------------------------------ full (non-TDA) BSE
// upstream BSE::Solve_nonhermitian_Davidson / HamiltonianOperator<A,B> / DavidsonSolver "HAM" mode:
//   [[A, B], [-B, -A]] [X; Y] = w [X; Y].
// Solved in the symmetric-product form: with M+ = A + B, M- = A - B and an orthonormal basis b of the vc-dimensional
// space, the reduced matrices m+- = b^T M+- b give (m-)^1/2 m+ (m-)^1/2 z = w^2 z; |X+Y> = b (m-)^1/2 z / sqrt(w),
// |X-Y> = b (m-)^-1/2 z sqrt(w) (so (X+Y)^T (X-Y) = X^T X - Y^T Y = 1); residuals M+|X+Y> - w|X-Y> and
// M-|X-Y> - w|X+Y> feed the diagonally preconditioned corrections.  Two operator applications (A, B) per new basis
// vector; the subspace algebra (<= max_search_space) is small dense work (cuSOLVER eigh + host products).
namespace {

// symmetric s x s (host, col-major): A <- eigenvectors, w <- ascending eigenvalues
void small_eigh(Context* ctx, int n, std::vector<double>& A, std::vector<double>& w, DBuf& dA, DBuf& dw) {
  w.resize((size_t)n);
  if (n <= kHostEighMax) {
    std::vector<double> keep = A;
    if (host_eigh(n, A.data(), w.data())) return;
    A.swap(keep);
  }
  dA.ensure((size_t)n * n);
  dw.ensure((size_t)n);
  ctx->h2d(dA.p, A.data(), (size_t)n * n);
  ctx->eigh(n, dA.p, n, dw.p);
  w.resize((size_t)n);
  ctx->d2h(w.data(), dw.p, (size_t)n);
  ctx->d2h(A.data(), dA.p, (size_t)n * n);
}
// C = A * B for s x s host matrices (col-major)
std::vector<double> small_mm(int n, const std::vector<double>& A, const std::vector<double>& B) {
  std::vector<double> C((size_t)n * n, 0.0);
  for (int j = 0; j < n; ++j)
    for (int k = 0; k < n; ++k) {
      const double b = B[k + (size_t)j * n];
      if (b == 0.0) continue;
      for (int i = 0; i < n; ++i) C[i + (size_t)j * n] += A[i + (size_t)k * n] * b;
    }
  return C;
}

}  // namespace

void btda_solve(Operator& A, Operator& B, long long neigen, const xtpb_davidson_options& opt, BtdaResult& out) {
  Context* ctx = A.ctx;
  ProfScope prof(PROF_DAVIDSON);
  const long long n = A.size;
  XTPB_REQUIRE(B.size == n, "A and B operators differ in size");
  XTPB_REQUIRE(neigen >= 1 && neigen <= n, "neigen out of range");
  const int k = (int)neigen;
  long long max_space = opt.max_search_space;
  if (max_space < 4 * neigen) max_space = 10 * neigen;
  max_space = std::min(max_space, n);
  long long guess = opt.size_initial_guess == 0 ? 2 * neigen : opt.size_initial_guess;
  guess = std::min(guess, n);

  DavidsonWork w;
  w.ctx = ctx;
  w.n = n;
  w.ld = round_up(n, 2);
  w.cap = std::min<long long>(n, max_space + 2 * k) + guess + 2;
  w.V.alloc((size_t)(w.ld * w.cap));
  w.V.zero(ctx->stream);
  w.small.alloc((size_t)(w.cap * (w.cap + 4) + 16));
  w.vec.alloc((size_t)(2 * n + 1024));
  w.tmp.alloc((size_t)(w.ld * 2 * k));
  DBuf P((size_t)(w.ld * w.cap)), Q((size_t)(w.ld * w.cap)), T1((size_t)(w.ld * w.cap)), T2((size_t)(w.ld * w.cap));
  DBuf XpY((size_t)(w.ld * k)), XmY((size_t)(w.ld * k)), RL((size_t)(w.ld * k)), RR((size_t)(w.ld * k));
  DBuf D((size_t)n), dsmall, dw, Ldev((size_t)(w.cap * k)), Rdev((size_t)(w.cap * k)), omdev((size_t)k), nrm((size_t)(2 * k));
  A.diagonal_dev(D.p);
  {
    std::vector<double> Dh((size_t)n);
    ctx->d2h(Dh.data(), D.p, (size_t)n);
    std::vector<long long> order((size_t)n);
    std::iota(order.begin(), order.end(), 0LL);
    std::stable_sort(order.begin(), order.end(), [&](long long a, long long b) { return Dh[a] < Dh[b]; });
    DBuf idx((size_t)guess);
    XTPB_CUDA(cudaMemcpyAsync(idx.p, order.data(), (size_t)guess * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_unit_vectors(w.V.p, w.ld, n, reinterpret_cast<const long long*>(idx.p), (int)guess, ctx->stream);
    ctx->sync();
  }
  auto apply = [&](int from, int to) {      // P, Q columns [from, to) = (A +- B) V
    const int cnt = to - from;
    if (cnt <= 0) return;
    const long long off = (long long)from * w.ld;
    A.matmul_dev(w.V.p + off, w.ld, cnt, T1.p, w.ld);
    B.matmul_dev(w.V.p + off, w.ld, cnt, T2.p, w.ld);
    const long long len = w.ld * (cnt - 1) + n;
    XTPB_CUDA(cudaMemcpyAsync(P.p + off, T1.p, (size_t)len * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    XTPB_CUDA(cudaMemcpyAsync(Q.p + off, T1.p, (size_t)len * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    k_axpby(P.p + off, T2.p, len, 1.0, 1.0, ctx->stream);
    k_axpby(Q.p + off, T2.p, len, -1.0, 1.0, ctx->stream);
  };

  int s = (int)guess, applied = 0;
  std::vector<double> omega((size_t)k, 0.0), rn((size_t)(2 * k));
  out.info = 1;
  out.iterations = 0;
  for (long long it = 0; it < opt.iter_max; ++it) {
    out.iterations = it + 1;
    apply(applied, s);
    applied = s;
    // reduced matrices
    std::vector<double> mp((size_t)s * s), mm((size_t)s * s);
    gemm_tn(ctx, w.V.p, w.ld, s, P.p, w.ld, s, n, w.small.p, s);
    ctx->d2h(mp.data(), w.small.p, (size_t)s * s);
    gemm_tn(ctx, w.V.p, w.ld, s, Q.p, w.ld, s, n, w.small.p, s);
    ctx->d2h(mm.data(), w.small.p, (size_t)s * s);
    for (int j = 0; j < s; ++j)
      for (int i = 0; i < j; ++i) {
        mp[i + (size_t)j * s] = mp[j + (size_t)i * s] = 0.5 * (mp[i + (size_t)j * s] + mp[j + (size_t)i * s]);
        mm[i + (size_t)j * s] = mm[j + (size_t)i * s] = 0.5 * (mm[i + (size_t)j * s] + mm[j + (size_t)i * s]);
      }
    // (m-)^{+-1/2}
    std::vector<double> U = mm, lam;
    small_eigh(ctx, s, U, lam, dsmall, dw);
    XTPB_REQUIRE(lam[0] > 0.0, "full BSE: A - B is not positive definite in the search space (triplet instability?)");
    std::vector<double> Sh((size_t)s * s, 0.0), Si((size_t)s * s, 0.0);
    for (int j = 0; j < s; ++j)
      for (int i = 0; i < s; ++i) {
        double a = 0.0, b = 0.0;
        for (int t = 0; t < s; ++t) {
          const double uu = U[i + (size_t)t * s] * U[j + (size_t)t * s];
          a += uu * std::sqrt(lam[t]);
          b += uu / std::sqrt(lam[t]);
        }
        Sh[i + (size_t)j * s] = a;
        Si[i + (size_t)j * s] = b;
      }
    std::vector<double> Cm = small_mm(s, small_mm(s, Sh, mp), Sh), w2;
    for (int j = 0; j < s; ++j)
      for (int i = 0; i < j; ++i) Cm[i + (size_t)j * s] = Cm[j + (size_t)i * s] = 0.5 * (Cm[i + (size_t)j * s] + Cm[j + (size_t)i * s]);
    small_eigh(ctx, s, Cm, w2, dsmall, dw);            // Cm <- Z
    const int kk = std::min(k, s);
    std::vector<double> L((size_t)s * kk), R((size_t)s * kk);
    for (int j = 0; j < kk; ++j) {
      XTPB_REQUIRE(w2[j] > 0.0, "full BSE: non-positive squared excitation energy");
      omega[j] = std::sqrt(w2[j]);
      const double sq = std::sqrt(omega[j]);
      for (int i = 0; i < s; ++i) {
        double a = 0.0, b = 0.0;
        for (int t = 0; t < s; ++t) {
          a += Sh[i + (size_t)t * s] * Cm[t + (size_t)j * s];
          b += Si[i + (size_t)t * s] * Cm[t + (size_t)j * s];
        }
        L[i + (size_t)j * s] = a / sq;
        R[i + (size_t)j * s] = b * sq;
      }
    }
    ctx->h2d(Ldev.p, L.data(), L.size());
    ctx->h2d(Rdev.p, R.data(), R.size());
    ctx->h2d(omdev.p, omega.data(), (size_t)kk);
    gemm_nn(ctx, w.V.p, w.ld, s, Ldev.p, s, kk, n, XpY.p, w.ld, 1.0, 0.0);
    gemm_nn(ctx, w.V.p, w.ld, s, Rdev.p, s, kk, n, XmY.p, w.ld, 1.0, 0.0);
    gemm_nn(ctx, P.p, w.ld, s, Ldev.p, s, kk, n, RL.p, w.ld, 1.0, 0.0);      // M+ |X+Y>
    gemm_nn(ctx, Q.p, w.ld, s, Rdev.p, s, kk, n, RR.p, w.ld, 1.0, 0.0);      // M- |X-Y>
    k_residuals(RL.p, w.ld, XmY.p, w.ld, omdev.p, n, kk, ctx->stream);       //   - w |X-Y>
    k_residuals(RR.p, w.ld, XpY.p, w.ld, omdev.p, n, kk, ctx->stream);       //   - w |X+Y>
    k_col_norms(RL.p, w.ld, n, kk, nrm.p, ctx->stream);
    k_col_norms(RR.p, w.ld, n, kk, nrm.p + kk, ctx->stream);
    ctx->d2h(rn.data(), nrm.p, (size_t)(2 * kk));
    bool converged = kk == k;
    for (int j = 0; j < kk; ++j) converged = converged && rn[j] < opt.tolerance && rn[kk + j] < opt.tolerance;
    if (converged) {
      out.info = 0;
      break;
    }
    if (it == opt.iter_max - 1) break;
    if (s + 2 * kk > max_space) {       // restart from the current Ritz pairs
      XTPB_CUDA(cudaMemcpyAsync(w.V.p, XpY.p, (size_t)(w.ld * kk) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      XTPB_CUDA(cudaMemcpyAsync(w.V.p + (long long)kk * w.ld, XmY.p, (size_t)(w.ld * kk) * 8, cudaMemcpyDeviceToDevice,
                                ctx->stream));
      s = gram_schmidt_columns(w, 0, 2 * kk);
      applied = 0;
    }
    int added = 0;
    for (int side = 0; side < 2; ++side)
      for (int j = 0; j < kk; ++j) {
        if (rn[side * kk + j] < opt.tolerance) continue;
        XTPB_REQUIRE(s + added < w.cap, "full BSE: search space overflow");
        const double* r = (side == 0 ? RL.p : RR.p) + (long long)j * w.ld;
        const double* x = (side == 0 ? XpY.p : XmY.p) + (long long)j * w.ld;
        k_davidson_correction(w.V.p + (long long)(s + added) * w.ld, r, x, D.p, omega[j], n, 0, w.vec.p, ctx->stream);
        ++added;
      }
    const int before = s;
    s = gram_schmidt(w, before, before + added);
    if (s == before) break;
  }
  out.evals.assign(omega.begin(), omega.end());
  out.X.alloc((size_t)(n * k));
  out.Y.alloc((size_t)(n * k));
  // X = (|X+Y> + |X-Y>)/2, Y = (|X+Y> - |X-Y>)/2
  k_copy_2d(out.X.p, n, XpY.p, w.ld, (int)n, k, ctx->stream);
  k_copy_2d(out.Y.p, n, XpY.p, w.ld, (int)n, k, ctx->stream);
  k_copy_2d(w.tmp.p, n, XmY.p, w.ld, (int)n, k, ctx->stream);
  k_axpby(out.X.p, w.tmp.p, n * k, 0.5, 0.5, ctx->stream);
  k_axpby(out.Y.p, w.tmp.p, n * k, -0.5, 0.5, ctx->stream);
  ctx->sync();
}

}  // namespace xtpb
