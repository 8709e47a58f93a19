// Sigma_Exact and Sigma_CDA on the device (upstream xtp/src/libxtp/gwbse/sigma_exact.cc, sigma_cda.cc,
// ImaginaryAxisIntegration.cc, gaussian_quadrature.cc, and RPA::Diagonalize_H2p in rpa.cc).
//
// Sigma_Exact: the two-particle RPA Hamiltonian (size o*u) is diagonalised once (cuSOLVER), the residues
//   R[level][s][m] = sum_P M[level](m,P) sum_{vc} M[v](c,P) (X+Y)_{vc,s} are three contractions on the DMMA engine,
//   and every (level, frequency) evaluation is one fused reduction over (s, m).  Dense (o*u)^2 storage: small
//   molecules only, single GPU.
// Sigma_CDA: kappa(i w_j) = eps^-1(i w_j) - 1 at the Gauss points and kappa(0) are built once (batched epsilon +
//   Cholesky inverses); the frequency-independent quadratic forms q_j[level][m] = M(m,:) K_j M(m,:)^T are batched
//   contractions, so the imaginary-axis integral and the Gaussian tail of any (level, frequency) pair reduce to one
//   fused kernel over (j, m).  Residues need eps^-1(|e_m - w|) per enclosed pole: batched epsilon + LU inverse +
//   a quadratic-form kernel.  Works on the multi-GPU layout (local second-index columns, all-reduced partials).
#include <algorithm>
#include <cmath>
#include <numeric>

#include "internal.h"

namespace xtpb {

namespace {

constexpr double kPi = 3.14159265358979323846264338327950288;

#define LAUNCH_CHECK_SO()              \
  do {                                 \
    XTPB_CUDA(cudaGetLastError());     \
    ++g_launch_count;                  \
  } while (0)

inline int nblocks(long long n, int threads, int cap = 8192) {
  long long b = (n + threads - 1) / threads;
  return (int)std::max<long long>(1, std::min<long long>(b, cap));
}

__device__ __forceinline__ double warp_sum_so(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum of two values (256 threads), valid in thread 0
__device__ __forceinline__ void block_sum2(double& a, double& b, double* sh) {
  a = warp_sum_so(a);
  b = warp_sum_so(b);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[w] = a; sh[8 + w] = b; }
  __syncthreads();
  if (threadIdx.x < 32) {
    a = threadIdx.x < 8 ? sh[threadIdx.x] : 0.0;
    b = threadIdx.x < 8 ? sh[8 + threadIdx.x] : 0.0;
    a = warp_sum_so(a);
    b = warp_sum_so(b);
  }
  __syncthreads();
}

// ---------------------------------------------------------------- Sigma_Exact kernels
// C = (ApB + diag(AmB)) scaled by s_i s_j, s = sqrt(AmB); AmB_i = e_c - e_v for i = v*nun + c
__global__ void exact_build_c_kernel(double* __restrict__ C, long long ld, int rs, int nocc, int nun,
                                     const double* __restrict__ e) {
  const long long total = (long long)rs * rs;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % rs), j = (int)(idx / rs);
    const double di = e[nocc + i % nun] - e[i / nun], dj = e[nocc + j % nun] - e[j / nun];
    double v = C[i + j * ld];
    if (i == j) v += di;
    C[i + j * ld] = v * sqrt(di) * sqrt(dj);
  }
}
// XpY[r,s] = sqrt(AmB_r) Z[r,s] / sqrt(omega_s)
__global__ void exact_xpy_kernel(double* __restrict__ Z, long long ld, int rs, int nocc, int nun,
                                 const double* __restrict__ e, const double* __restrict__ omega) {
  const long long total = (long long)rs * rs;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % rs), s = (int)(idx / rs);
    Z[r + s * ld] *= sqrt(e[nocc + r % nun] - e[r / nun]) / sqrt(omega[s]);
  }
}
// one CTA per (level, frequency) pair: 2 sum_{s,m} R^2 t/(t^2+eta^2), t = w - e_m +/- Omega_s, and its derivative
__global__ void __launch_bounds__(256) exact_pairs_kernel(const double* __restrict__ R, long long ldm, int rs,
                                                          int ntotal, int nocc, const double* __restrict__ e,
                                                          const double* __restrict__ omega_s,
                                                          const int* __restrict__ pair_level,
                                                          const double* __restrict__ pair_omega, double eta2,
                                                          double* __restrict__ values, double* __restrict__ derivs) {
  __shared__ double sh[16];
  const int pair = blockIdx.x;
  const double* Rl = R + (long long)pair_level[pair] * rs * ldm;
  const double om = pair_omega[pair];
  double val = 0.0, der = 0.0;
  const long long total = (long long)rs * ldm;
  for (long long idx = threadIdx.x; idx < total; idx += 256) {
    const int m = (int)(idx % ldm), s = (int)(idx / ldm);
    if (m >= ntotal) continue;
    const double r = Rl[idx];
    const double t = om - e[m] + (m < nocc ? omega_s[s] : -omega_s[s]);
    const double den = t * t + eta2, r2 = r * r;
    val += r2 * t / den;
    der += r2 * (eta2 - t * t) / (den * den);
  }
  block_sum2(val, der, sh);
  if (threadIdx.x == 0) {
    values[pair] = 2.0 * val;
    derivs[pair] = 2.0 * der;
  }
}
// W[l][s][m] = R[l][s][m] * t/(t^2+eta^2) at the level's own frequency
__global__ void exact_weighted_kernel(double* __restrict__ W, const double* __restrict__ R, long long ldm, int rs,
                                      int ntotal, int nocc, int nlevels, const double* __restrict__ e,
                                      const double* __restrict__ omega_s, const double* __restrict__ level_omega,
                                      double eta2) {
  const long long per = (long long)rs * ldm, total = per * nlevels;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(idx % ldm), s = (int)((idx / ldm) % rs), l = (int)(idx / per);
    double out = 0.0;
    if (m < ntotal) {
      const double t = level_omega[l] - e[m] + (m < nocc ? omega_s[s] : -omega_s[s]);
      out = R[idx] * t / (t * t + eta2);
    }
    W[idx] = out;
  }
}

// ---------------------------------------------------------------- Sigma_CDA kernels
// q[l][m] = sum_P S_l[P][m] * Y_l[P][m]   (column dots, coalesced over m)
__global__ void cda_column_dots_kernel(double* __restrict__ q, long long q_level_stride, const double* __restrict__ S,
                                       const double* __restrict__ Y, long long ldn, long long slab, int ntotal,
                                       int naux, int nlevels) {
  const long long total = (long long)nlevels * ntotal;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(idx % ntotal), l = (int)(idx / ntotal);
    const double* s = S + (long long)l * slab + m;
    const double* y = Y + (long long)l * slab + m;
    double acc = 0.0;
    for (int P = 0; P < naux; ++P) acc += s[(long long)P * ldn] * y[(long long)P * ldn];
    q[(long long)l * q_level_stride + m] = acc;
  }
}
// one CTA per pair: imaginary-axis Gauss quadrature + Gaussian tail from the precomputed quadratic forms
//   Q[level][j][m], j < order: kernel j; j == order: kappa(0)
__global__ void __launch_bounds__(256) cda_pairs_kernel(const double* __restrict__ Q, long long ldq, int order,
                                                        int ntotal, int nocc, const double* __restrict__ e,
                                                        const double* __restrict__ gq_points,
                                                        const double* __restrict__ gq_weights,
                                                        const int* __restrict__ pair_level,
                                                        const double* __restrict__ pair_omega, double eta,
                                                        double alpha, double* __restrict__ values) {
  __shared__ double sh[16];
  const int pair = blockIdx.x;
  const double* Ql = Q + (long long)pair_level[pair] * (order + 1) * ldq;
  const double om = pair_omega[pair];
  double gq = 0.0, tail = 0.0;
  for (int idx = threadIdx.x; idx < (order + 1) * ntotal; idx += 256) {
    const int m = idx % ntotal, j = idx / ntotal;
    const double qv = Ql[(long long)j * ldq + m];
    if (j < order) {
      const double x = om - e[m], sg = m < nocc ? eta : -eta, w = gq_points[j];
      const double x2 = x * x, a = sg + w, b = sg - w;
      gq += gq_weights[j] * qv * (x / (x2 + a * a) + x / (x2 + b * b));
    } else if (alpha != 0.0) {
      const double delta = e[m] - om;
      if (fabs(delta) > 1e-10) tail += qv * 0.5 * copysign(1.0, delta) * erfcx(fabs(alpha * delta));
    }
  }
  block_sum2(gq, tail, sh);
  if (threadIdx.x == 0) values[pair] = 0.5 / kPi * gq + tail;
}
// one CTA per residue item: v^T (Ainv - 1) v with v = slab[:, col]; Ainv symmetric column-major
// out[it] = v_it . (x_it - v_it) = v^T (eps^-1 - 1) v for the items this rank evaluates (active[it] >= 0), 0 otherwise
__global__ void __launch_bounds__(256) cda_residue_dot_kernel(const double* __restrict__ Vb,
                                                              const double* __restrict__ Xb, int naux,
                                                              const int* __restrict__ active, double* __restrict__ out) {
  __shared__ double sh[16];
  const int it = blockIdx.x;
  if (active[it] < 0) {
    if (threadIdx.x == 0) out[it] = 0.0;
    return;
  }
  const double* v = Vb + (long long)it * naux;
  const double* x = Xb + (long long)it * naux;
  double quad = 0.0, vv = 0.0;
  for (int P = threadIdx.x; P < naux; P += 256) {
    quad += v[P] * x[P];
    vv += v[P] * v[P];
  }
  block_sum2(quad, vv, sh);
  if (threadIdx.x == 0) out[it] = quad - vv;
}
// Vb[it][P] = slab(item)[P][col] for the items whose column this rank owns, 0 otherwise (summed over ranks afterwards)
__global__ void cda_gather_columns_kernel(double* __restrict__ Vb, int naux, const double* __restrict__ M, long long ldn,
                                          long long slab, const int* __restrict__ item_slab,
                                          const int* __restrict__ item_col, int n_items) {
  const long long total = (long long)n_items * naux;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int it = (int)(idx / naux), P = (int)(idx % naux);
    const int col = item_col[it];
    Vb[idx] = col < 0 ? 0.0 : M[(long long)item_slab[it] * slab + (long long)P * ldn + col];
  }
}

// ---------------------------------------------------------------- Gauss quadrature nodes (host)
// Golub-Welsch: nodes = eigenvalues of the Jacobi matrix (implicit QL), polished by Newton on the orthonormal
// three-term recurrence; weights from the Christoffel sum 1 / sum_k p_k(x)^2 (no cancellation).
void tridiag_eigenvalues(std::vector<double> d, std::vector<double> e, std::vector<double>& out) {
  const int n = (int)d.size();
  e.push_back(0.0);
  for (int l = 0; l < n; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; ++m) {
        const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
      }
      if (m != l) {
        XTPB_REQUIRE(iter++ < 200, "Gauss quadrature: QL iteration did not converge");
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + std::copysign(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i], b = c * e[i];
          e[i + 1] = (r = std::hypot(f, g));
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          d[i + 1] = g + (p = s * r);
          g = c * r - b;
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  std::sort(d.begin(), d.end());
  out = d;
}

void gauss_rule(const std::vector<double>& a, const std::vector<double>& b, double mu0, std::vector<double>& x,
                std::vector<double>& w) {
  // Jacobi matrix: diagonal a[0..n-1], off-diagonal b[1..n-1]; b[n] closes the recurrence for p_n
  const int n = (int)a.size();
  tridiag_eigenvalues(a, std::vector<double>(b.begin() + 1, b.begin() + n), x);
  w.resize((size_t)n);
  for (int i = 0; i < n; ++i) {
    for (int newton = 0; newton < 3; ++newton) {
      // p_{k+1} = ((x - a_k) p_k - b_k p_{k-1}) / b_{k+1}, with derivatives
      double pm = 0.0, p = 1.0 / std::sqrt(mu0), dpm = 0.0, dp = 0.0;
      for (int k = 0; k < n; ++k) {
        const double pn = ((x[i] - a[k]) * p - b[k] * pm) / b[k + 1];
        const double dpn = (p + (x[i] - a[k]) * dp - b[k] * dpm) / b[k + 1];
        pm = p; p = pn; dpm = dp; dp = dpn;
      }
      if (dp != 0.0 && std::isfinite(p / dp)) x[i] -= p / dp;
    }
    double pm = 0.0, p = 1.0 / std::sqrt(mu0), sum = 0.0;
    for (int k = 0; k < n; ++k) {
      sum += p * p;
      const double pn = ((x[i] - a[k]) * p - b[k] * pm) / b[k + 1];
      pm = p; p = pn;
    }
    w[i] = 1.0 / sum;
  }
}

}  // namespace

// GaussianQuadrature (gaussian_quadrature.cc): scaled points / weights on (0, inf)
void gaussian_quadrature(int scheme, long long order, std::vector<double>& points, std::vector<double>& weights) {
  XTPB_REQUIRE(order >= 2 && order <= 200, "quadrature order out of range");
  std::vector<double> a, b, x, w;
  points.clear();
  weights.clear();
  if (scheme == XTPB_QUAD_LEGENDRE) {
    const int n = (int)order;
    a.assign((size_t)n, 0.0);
    b.assign((size_t)n + 1, 0.0);
    for (int k = 1; k <= n; ++k) b[k] = k / std::sqrt(4.0 * k * k - 1.0);
    gauss_rule(a, b, 2.0, x, w);
    for (int i = 0; i < n; ++i) {     // omega = 0.5 (1+x)/(1-x) maps (-1,1) -> (0,inf)
      points.push_back(0.5 * (1.0 + x[i]) / (1.0 - x[i]));
      weights.push_back(w[i] / ((1.0 - x[i]) * (1.0 - x[i])));
    }
  } else if (scheme == XTPB_QUAD_LAGUERRE) {
    const int n = (int)order;
    a.resize((size_t)n);
    b.assign((size_t)n + 1, 0.0);
    for (int k = 0; k < n; ++k) a[k] = 2.0 * k + 1.0;
    for (int k = 1; k <= n; ++k) b[k] = (double)k;
    gauss_rule(a, b, 1.0, x, w);
    for (int i = 0; i < n; ++i) {
      points.push_back(x[i]);
      weights.push_back(w[i] * std::exp(x[i]));
    }
  } else if (scheme == XTPB_QUAD_HERMITE) {
    const int n = 2 * (int)order;
    a.assign((size_t)n, 0.0);
    b.assign((size_t)n + 1, 0.0);
    for (int k = 1; k <= n; ++k) b[k] = std::sqrt(0.5 * k);
    gauss_rule(a, b, std::sqrt(kPi), x, w);
    for (int i = 0; i < n; ++i)
      if (x[i] > 0.0) {
        points.push_back(x[i]);
        weights.push_back(w[i] * std::exp(x[i] * x[i]));
      }
  } else {
    throw Error("xtpb: unknown quadrature scheme");
  }
}

// ================================================================ Sigma_Exact
void GW::prepare_exact() {
  tc->flush();
  ProfScope prof(PROF_EXACT);
  XTPB_REQUIRE(ctx->world == 1, "Sigma_Exact holds a dense (o*u)^2 matrix and runs on a single GPU");
  const long long na = tc->naux, nocc = n_occ, nun = rpatotal - n_occ;
  const long long rs = nocc * nun;
  XTPB_REQUIRE(rs <= 24576, "Sigma_Exact: o*u too large for the dense two-particle Hamiltonian (small molecules only)");
  rpasize = rs;
  const long long rld = round_up(rs, 2), nald = round_up(na, 2);
  // I[(v,c), P] = M[v](c,P), column-major rs x naux
  DBuf I((size_t)(rld * na)), C((size_t)(rld * rs)), w2((size_t)rs);
  I.zero(ctx->stream);
  k_extract_window(I.p, rld, nun, tc->M.p, tc->ldn, tc->slab, 0, (int)nocc, (int)nocc, (int)nun, (int)na, nullptr,
                   ctx->stream);
  {   // ApB - diag = 4 I I^T
    GemmParams g{};
    g.A = op_rows_contig(I.p, rld);
    g.B = g.A;
    g.C = C.p; g.c_sm = 1; g.c_sn = rld;
    g.M = g.N = (int)rs; g.K = (int)na; g.n_outer = 1; g.n_batch = 1; g.alpha = 4.0; g.lower = 1;
    contract(g, ctx->ws, ctx->stream);
    symmetrize_from_lower(C.p, (int)rs, rld, 0.0, ctx->stream);
  }
  exact_build_c_kernel<<<nblocks(rs * rs, 256), 256, 0, ctx->stream>>>(C.p, rld, (int)rs, (int)nocc, (int)nun,
                                                                      energies_dev.p);
  LAUNCH_CHECK_SO();
  ctx->eigh((int)rs, C.p, rld, w2.p);            // C <- Z
  rpa_omegas.resize((size_t)rs);
  ctx->d2h(rpa_omegas.data(), w2.p, (size_t)rs);
  for (auto& v : rpa_omegas) {
    XTPB_REQUIRE(v > 0.0, "Sigma_Exact: the RPA two-particle Hamiltonian is not positive definite");
    v = std::sqrt(v);
  }
  exact_omega_dev.ensure((size_t)rs);
  ctx->h2d(exact_omega_dev.p, rpa_omegas.data(), (size_t)rs);
  exact_xpy_kernel<<<nblocks(rs * rs, 256), 256, 0, ctx->stream>>>(C.p, rld, (int)rs, (int)nocc, (int)nun,
                                                                  energies_dev.p, exact_omega_dev.p);
  LAUNCH_CHECK_SO();
  // T[P,s] = sum_r I[r,P] XpY[r,s]
  DBuf T((size_t)(nald * rs));
  {
    GemmParams g{};
    g.A = op_k_contig(I.p, rld);
    g.B = op_k_contig(C.p, rld);
    g.C = T.p; g.c_sm = 1; g.c_sn = nald;
    g.M = (int)na; g.N = (int)rs; g.K = (int)rs; g.n_outer = 1; g.n_batch = 1; g.alpha = 1.0;
    contract(g, ctx->ws, ctx->stream);
  }
  // residues[level][s][m] = sum_P M[level](m,P) T[P,s]
  const long long ldm = tc->ldn;
  residues.alloc((size_t)(qptotal * rs * ldm));
  residues.zero(ctx->stream);
  {
    GemmParams g{};
    g.A = GemmOperand{tc->slab_ptr(q0), 1, tc->ldn, 0, tc->slab};
    g.B = op_k_contig(T.p, nald);
    g.C = residues.p; g.c_sm = 1; g.c_sn = ldm; g.c_batch = rs * ldm;
    g.M = (int)tc->ntotal; g.N = (int)rs; g.K = (int)na; g.n_outer = 1; g.n_batch = (int)qptotal; g.alpha = 1.0;
    contract(g, ctx->ws, ctx->stream);
  }
  ctx->sync();
}

// ================================================================ Sigma_CDA
void GW::prepare_cda() {
  tc->flush();
  ProfScope prof(PROF_CDA);
  const long long na = tc->naux, nn = na * na;
  gaussian_quadrature(opt.quadrature_scheme, opt.order, quad_points, quad_weights);
  const int order = (int)quad_points.size();
  cda_kernels.ensure((size_t)((order + 1) * nn));
  // Frequency sharding (BASELINE.json configs[2], north_star "frequency points for epsilon and Sigma"): with several
  // ranks, eps(i w_j) is summed onto rank j % world only, which alone inverts it (the N_aux^3 work per node is done
  // once, not world times); kappa(0) is owned by rank order % world.  The finished kernels are then broadcast, because
  // every rank needs all of them for the quadratic forms over ITS share of the second tensor index.
  const int world = ctx->world, rank = ctx->rank;
  auto owner = [&](int j) { return j % world; };
  rpa_epsilon_dev(*tc, energies_dev.p, n_occ, opt.eta, quad_points.data(), order, true, 0.0, cda_kernels.p, 0);
  const double zero = 0.0;
  double* kappa0 = cda_kernels.p + (long long)order * nn;
  rpa_epsilon_dev(*tc, energies_dev.p, n_occ, opt.eta, &zero, 1, false, 0.0, kappa0, order);
  if (owner(order) == rank) {
    ctx->spd_inverse((int)na, kappa0, na);
    k_add_diagonal(kappa0, (int)na, na, -1.0, ctx->stream);
  }
  ctx->bcast(kappa0, (size_t)nn, owner(order));
  for (int j = 0; j < order; ++j) {
    if (owner(j) != rank) continue;
    double* K = cda_kernels.p + (long long)j * nn;
    ctx->spd_inverse((int)na, K, na);
    // dielinv_j = -(eps^-1 - 1) + exp(-(alpha w_j)^2) kappa0
    const double c = std::exp(-(opt.alpha * quad_points[j]) * (opt.alpha * quad_points[j]));
    k_axpby(K, kappa0, nn, c, -1.0, ctx->stream);
    k_add_diagonal(K, (int)na, na, 1.0, ctx->stream);
  }
  ctx->group_start();
  for (int j = 0; j < order; ++j) ctx->bcast(cda_kernels.p + (long long)j * nn, (size_t)nn, owner(j));
  ctx->group_end();
  cda_points_dev.ensure((size_t)(2 * order));
  ctx->h2d(cda_points_dev.p, quad_points.data(), (size_t)order);
  ctx->h2d(cda_points_dev.p + order, quad_weights.data(), (size_t)order);

  // frequency-independent quadratic forms Q[level][j][m] = M(m,:) K_j M(m,:)^T over the local columns m
  const long long ldq = tc->ldn;
  cda_q.alloc((size_t)(qptotal * (order + 1) * ldq));
  cda_q.zero(ctx->stream);
  const long long budget = 1LL << 28;                     // doubles (2 GiB) for Y = K_j * slab
  const long long lc = std::max<long long>(1, std::min<long long>(qptotal, budget / tc->slab));
  DBuf Y((size_t)(lc * tc->slab));
  for (int j = 0; j <= order; ++j) {
    const double* K = cda_kernels.p + (long long)j * nn;
    for (long long l0 = 0; l0 < qptotal; l0 += lc) {
      const long long cnt = std::min(lc, qptotal - l0);
      GemmParams g{};
      g.A = op_k_contig(K, na);                            // K symmetric
      g.B = GemmOperand{tc->slab_ptr(q0 + l0), 1, tc->ldn, 0, tc->slab};
      g.C = Y.p; g.c_sm = tc->ldn; g.c_sn = 1; g.c_batch = tc->slab;
      g.M = (int)na; g.N = (int)tc->ntotal; g.K = (int)na; g.n_outer = 1; g.n_batch = (int)cnt; g.alpha = 1.0;
      contract(g, ctx->ws, ctx->stream);
      cda_column_dots_kernel<<<nblocks(cnt * tc->ntotal, 128), 128, 0, ctx->stream>>>(
          cda_q.p + (l0 * (order + 1) + j) * ldq, (long long)(order + 1) * ldq, tc->slab_ptr(q0 + l0), Y.p, tc->ldn,
          tc->slab, (int)tc->ntotal, (int)na, (int)cnt);
      LAUNCH_CHECK_SO();
    }
  }
  ctx->sync();
}

namespace {
// Sigma_CDA::CalcResiduePrefactor
double residue_prefactor(double e_f, double e_m, double frequency) {
  const double tol = 1e-10;
  if (e_f < e_m && e_m < frequency) return 1.0;
  if (e_f > e_m && e_m > frequency) return -1.0;
  if (std::fabs(e_m - frequency) < tol && e_f > e_m) return -0.5;
  if (std::fabs(e_m - frequency) < tol && e_f < e_m) return 0.5;
  return 0.0;
}
}  // namespace

void GW::cda_values(long long n, const long long* levels, const double* freqs, double* values) {
  ProfScope prof(PROF_CDA);
  const int order = (int)quad_points.size();
  const long long na = tc->naux, nn = na * na, ldq = tc->ldn;
  // ---- quadrature + tail: one fused kernel over the local columns
  std::vector<int> lv((size_t)n);
  for (long long i = 0; i < n; ++i) {
    XTPB_REQUIRE(levels[i] >= 0 && levels[i] < qptotal, "gw level out of range");
    lv[i] = (int)levels[i];
  }
  DBuf buf((size_t)(2 * n + (n + 1) / 2 + 1));
  double* om = buf.p;
  double* val = om + n;
  int* lvd = reinterpret_cast<int*>(val + n);
  ctx->h2d(om, freqs, (size_t)n);
  XTPB_CUDA(cudaMemcpyAsync(lvd, lv.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  cda_pairs_kernel<<<(unsigned)n, 256, 0, ctx->stream>>>(cda_q.p, ldq, order, (int)tc->ntotal, (int)n_occ_loc, e_loc,
                                                        cda_points_dev.p, cda_points_dev.p + order, lvd, om, opt.eta,
                                                        opt.alpha, val);
  LAUNCH_CHECK_SO();
  ctx->allreduce_sum(val, (size_t)n);
  ctx->d2h(values, val, (size_t)n);

  // ---- residues: every enclosed pole needs eps^-1(|e_i - w|) (host decides which, identical on every rank)
  struct Item { long long pair; int level, i; double fac, delta; };
  std::vector<Item> items;
  const long long homo = opt.homo - opt.rpamin;
  const double fermi = 0.5 * (rpa_energies[homo] + rpa_energies[homo + 1]);
  for (long long p = 0; p < n; ++p)
    for (long long i = 0; i < rpatotal; ++i) {
      const double fac = residue_prefactor(fermi, rpa_energies[i], freqs[p]);
      if (std::fabs(fac) > 1e-10) items.push_back(Item{p, lv[p], (int)i, fac, std::fabs(rpa_energies[i] - freqs[p])});
    }
  if (items.empty()) return;
  const long long bmax = std::max<long long>(1, std::min<long long>((long long)items.size(), (1LL << 28) / nn));
  DBuf eps((size_t)(bmax * nn)), res((size_t)bmax + 1);
  DBuf meta((size_t)(3 * ((bmax + 1) / 2 + 1)));
  int* slab_d = reinterpret_cast<int*>(meta.p);
  int* col_d = slab_d + bmax + 1;
  int* act_d = col_d + bmax + 1;
  // Pole sharding (world > 1): item k's eps(|e_i - w|) is summed onto rank k % world, which alone factorises it and
  // solves eps x = v; the vector v (one column of a slab, owned by the rank holding second-index column i) reaches it
  // through a small all-reduced gather buffer.  One LU factorisation + one triangular solve per item (the residue
  // needs v^T eps^-1 v for a single vector), not a full inverse.
  const int world = ctx->world, rank = ctx->rank;
  DBuf Vb((size_t)(bmax * na)), Xb((size_t)(bmax * na));
  std::vector<int> hs, hc, ha;
  std::vector<double> deltas, hres;
  for (size_t b0 = 0; b0 < items.size(); b0 += (size_t)bmax) {
    const long long cnt = (long long)std::min<size_t>((size_t)bmax, items.size() - b0);
    deltas.resize((size_t)cnt); hs.resize((size_t)cnt); hc.resize((size_t)cnt); ha.resize((size_t)cnt);
    for (long long k = 0; k < cnt; ++k) {
      const Item& it = items[b0 + k];
      deltas[k] = it.delta;
      hs[k] = (int)(q0 + it.level);
      hc[k] = (it.i % tc->world == tc->rank) ? it.i / tc->world : -1;     // owner of second-index column i
      ha[k] = (k % world == rank) ? 0 : -1;                               // >= 0: this rank evaluates item k
    }
    rpa_epsilon_dev(*tc, energies_dev.p, n_occ, opt.eta, deltas.data(), (int)cnt, false, 0.0, eps.p, 0);
    XTPB_CUDA(cudaMemcpyAsync(slab_d, hs.data(), (size_t)cnt * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    XTPB_CUDA(cudaMemcpyAsync(col_d, hc.data(), (size_t)cnt * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    XTPB_CUDA(cudaMemcpyAsync(act_d, ha.data(), (size_t)cnt * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    cda_gather_columns_kernel<<<nblocks(cnt * na, 256), 256, 0, ctx->stream>>>(Vb.p, (int)na, tc->M.p, tc->ldn, tc->slab,
                                                                             slab_d, col_d, (int)cnt);
    LAUNCH_CHECK_SO();
    ctx->allreduce_sum(Vb.p, (size_t)(cnt * na));
    XTPB_CUDA(cudaMemcpyAsync(Xb.p, Vb.p, (size_t)(cnt * na) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    for (long long k = 0; k < cnt; ++k)
      if (k % world == rank) ctx->lu_solve_vector((int)na, eps.p + k * nn, na, Xb.p + k * na);
    cda_residue_dot_kernel<<<(unsigned)cnt, 256, 0, ctx->stream>>>(Vb.p, Xb.p, (int)na, act_d, res.p);
    LAUNCH_CHECK_SO();
    ctx->allreduce_sum(res.p, (size_t)cnt);
    hres.resize((size_t)cnt);
    ctx->d2h(hres.data(), res.p, (size_t)cnt);
    for (long long k = 0; k < cnt; ++k) values[items[b0 + k].pair] += items[b0 + k].fac * hres[k];
  }
}

// ================================================================ dispatch for the non-PPM self-energies
void GW::sigma_c_diag_elements_other(long long n, const long long* levels, const double* freqs, double* values,
                                     double* derivs) {
  if (opt.sigma_integration == XTPB_SIGMA_EXACT) {
    ProfScope prof(PROF_EXACT);
    std::vector<int> lv((size_t)n);
    for (long long i = 0; i < n; ++i) {
      XTPB_REQUIRE(levels[i] >= 0 && levels[i] < qptotal, "gw level out of range");
      lv[i] = (int)levels[i];
    }
    DBuf buf((size_t)(3 * n + (n + 1) / 2 + 1));
    double* om = buf.p;
    double* val = om + n;
    double* der = val + n;
    int* lvd = reinterpret_cast<int*>(der + n);
    ctx->h2d(om, freqs, (size_t)n);
    XTPB_CUDA(cudaMemcpyAsync(lvd, lv.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    exact_pairs_kernel<<<(unsigned)n, 256, 0, ctx->stream>>>(residues.p, tc->ldn, (int)rpasize, (int)tc->ntotal,
                                                            (int)n_occ, energies_dev.p, exact_omega_dev.p, lvd, om,
                                                            opt.eta * opt.eta, val, der);
    LAUNCH_CHECK_SO();
    ctx->d2h(values, val, (size_t)n);
    if (derivs) ctx->d2h(derivs, der, (size_t)n);
    return;
  }
  XTPB_REQUIRE(opt.sigma_integration == XTPB_SIGMA_CDA, "unknown sigma_integration");
  cda_values(n, levels, freqs, values);
  if (derivs) {      // Sigma_CDA::CalcCorrelationDiagElementDerivative: central difference, h = 1e-3
    const double h = 1e-3;
    std::vector<double> fp(freqs, freqs + n), fm(freqs, freqs + n), vp((size_t)n), vm((size_t)n);
    for (long long i = 0; i < n; ++i) { fp[i] += h; fm[i] -= h; }
    cda_values(n, levels, fp.data(), vp.data());
    cda_values(n, levels, fm.data(), vm.data());
    for (long long i = 0; i < n; ++i) derivs[i] = (vp[i] - vm[i]) / (2.0 * h);
  }
}

void GW::sigma_c_offdiag_other(const double* freqs, double* out_host) {
  const long long q = qptotal;
  if (opt.sigma_integration == XTPB_SIGMA_CDA) {
    // upstream Sigma_CDA provides diagonal elements only; its off-diagonal correlation is zero
    std::fill(out_host, out_host + q * q, 0.0);
    return;
  }
  ProfScope prof(PROF_EXACT);
  const long long rs = rpasize, ldm = tc->ldn, per = rs * ldm;
  DBuf W((size_t)(q * per)), S((size_t)(q * q)), om((size_t)q);
  ctx->h2d(om.p, freqs, (size_t)q);
  exact_weighted_kernel<<<nblocks(q * per, 256), 256, 0, ctx->stream>>>(W.p, residues.p, ldm, (int)rs,
                                                                       (int)tc->ntotal, (int)n_occ, (int)q,
                                                                       energies_dev.p, exact_omega_dev.p, om.p,
                                                                       opt.eta * opt.eta);
  LAUNCH_CHECK_SO();
  GemmParams g{};
  g.A = op_k_contig(W.p, per);
  g.B = op_k_contig(residues.p, per);
  g.C = S.p; g.c_sm = 1; g.c_sn = q;
  g.M = g.N = (int)q; g.K = (int)per; g.n_outer = 1; g.n_batch = 1; g.alpha = 1.0;
  contract(g, ctx->ws, ctx->stream);
  std::vector<double> s((size_t)(q * q));
  ctx->d2h(s.data(), S.p, (size_t)(q * q));
  for (long long j = 0; j < q; ++j)
    for (long long i = 0; i < q; ++i) out_host[i + j * q] = i == j ? 0.0 : s[i + j * q] + s[j + i * q];
}

}  // namespace xtpb
