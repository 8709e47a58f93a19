/* CPU oracle, C part II -- TEST / BASELINE INFRASTRUCTURE ONLY (see gwbse_oracle.py header; PARITY UNPINNED).
 *
 * The quasiparticle solver and the Sigma_c loops of the GW step as complete OpenMP C routines, so that a FULL
 * reference-structure CPU step (oracle/cpu_step.py) can be timed at benzene / pentacene size instead of being
 * extrapolated from samples:
 *   solve_qp_grid_ppm       GW::SolveQP with qp_solver = grid for every gw level (upstream gw.cc: SolveQP_Grid ->
 *                           SolveQP_Bisection, root with the smallest |dSigma/dw - 1|, SolveQP_Linearisation fallback),
 *                           OpenMP over the levels as upstream; every Sigma_c value is one
 *                           Sigma_PPM::CalcCorrelationDiagElement call that streams the level's slab (sigma_ppm.cc).
 *   sigma_ppm_grid_batched  SAME-ALGORITHM variant of the grid scan (what the CUDA path does): one pass over the slab
 *                           per level, all grid frequencies of a block accumulated in registers.
 *   sigma_ppm_offdiag_all   Sigma_base::CalcCorrelationOffDiag: loop over level pairs, one
 *                           Sigma_PPM::CalcCorrelationOffDiagElement per pair (sigma_base.cc / sigma_ppm.cc).
 *   sigma_ppm_weighted_slab SAME-ALGORITHM building block of the off-diagonal part: W[l][P][m] = fac_P g(w_l - z) M[l][P][m]
 *                           (the q x q matrix is then one GEMM).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static const double kFourPi2 = 12.566370614359172953850573533118;

static inline double stab_inv(double x) {
  const double ax = fabs(x);
  if (ax >= 0.25) return 1.0 / x;
  if (x == 0.0) return 0.0;
  return 0.5 * (1.0 - cos(kFourPi2 * x)) / x;
}

/* Sigma_PPM::CalcCorrelationDiagElement (and its derivative) for one level at one frequency */
static double sigma_ppm_eval(const double* slab, long long ld, int ntotal, int naux, int n_occ, const double* e,
                             const double* ppm_freq, const double* ppm_fac, double om, double* deriv) {
  double val = 0.0, der = 0.0;
  for (int P = 0; P < naux; ++P) {
    const double fac = ppm_fac[P];
    if (fac == 0.0) continue;
    const double Om = ppm_freq[P];
    const double* row = slab + (size_t)P * (size_t)ld;
    double v1 = 0.0, d1 = 0.0;
    for (int m = 0; m < ntotal; ++m) {
      const double g = stab_inv(om - e[m] + (m < n_occ ? Om : -Om));
      const double a = row[m] * row[m] * g;
      v1 += a;
      d1 -= a * g;
    }
    val += fac * v1;
    der += fac * d1;
  }
  if (deriv) *deriv = der;
  return val;
}

/* M: slabs of the gw levels, level l at M + l*slab_stride, each [naux][ld] (level index m fastest).
 * result[l]: QP energy; converged[l]: 1 grid root, 0 linearised; evaluations: Sigma_c evaluations performed. */
void solve_qp_grid_ppm(const double* M, long long slab_stride, long long ld, int ntotal, int naux, int n_occ,
                       const double* energies, const double* ppm_freq, const double* ppm_fac, int n_levels,
                       const double* intercept, const double* frequency0, int steps, double spacing, double limit,
                       double* result, int* converged, long long* evaluations) {
  long long total_evals = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total_evals)
  for (int l = 0; l < n_levels; ++l) {
    const double* slab = M + (size_t)l * (size_t)slab_stride;
    const double range = spacing * (double)(steps - 1) / 2.0;
    long long evals = 0;
    double fprev = frequency0[l] - range;
    double tprev = sigma_ppm_eval(slab, ld, ntotal, naux, n_occ, energies, ppm_freq, ppm_fac, fprev, NULL) +
                   intercept[l] - fprev;
    ++evals;
    double best = 0.0, grad_max = INFINITY;
    int found = 0;
    for (int i = 1; i < steps; ++i) {
      const double f = frequency0[l] - range + (double)i * spacing;
      const double t = sigma_ppm_eval(slab, ld, ntotal, naux, n_occ, energies, ppm_freq, ppm_fac, f, NULL) +
                       intercept[l] - f;
      ++evals;
      if (tprev * t < 0.0) {
        /* GW::SolveQP_Bisection */
        double lo = fprev, flo = tprev, hi = f, root;
        for (;;) {
          const double c = 0.5 * (lo + hi);
          if (fabs(hi - lo) < limit) { root = c; break; }
          const double yc = sigma_ppm_eval(slab, ld, ntotal, naux, n_occ, energies, ppm_freq, ppm_fac, c, NULL) +
                            intercept[l] - c;
          ++evals;
          if (fabs(yc) < limit) { root = c; break; }
          if (yc * flo > 0) { lo = c; flo = yc; } else { hi = c; }
        }
        double d = 0.0;
        sigma_ppm_eval(slab, ld, ntotal, naux, n_occ, energies, ppm_freq, ppm_fac, root, &d);
        ++evals;
        const double g = fabs(d - 1.0);
        if (g < grad_max) { best = root; grad_max = g; found = 1; }
      }
      fprev = f;
      tprev = t;
    }
    if (found) {
      result[l] = best;
      converged[l] = 1;
    } else {                                   /* GW::SolveQP_Linearisation */
      double d = 0.0;
      const double s = sigma_ppm_eval(slab, ld, ntotal, naux, n_occ, energies, ppm_freq, ppm_fac, frequency0[l], &d);
      ++evals;
      const double Z = 1.0 - d;
      result[l] = fabs(Z) > 1e-9 ? frequency0[l] + (intercept[l] - frequency0[l] + s) / Z : frequency0[l];
      converged[l] = 0;
    }
    total_evals += evals;
  }
  if (evaluations) *evaluations = total_evals;
}

/* values[l*steps + j] = Sigma_c(level l, omega0[l] + j*spacing): one pass over the slab, frequencies blocked */
void sigma_ppm_grid_batched(const double* M, long long slab_stride, long long ld, int ntotal, int naux, int n_occ,
                            const double* energies, const double* ppm_freq, const double* ppm_fac, int n_levels,
                            const double* omega0, double spacing, int steps, double* values) {
  enum { WB = 16 };
  const int nblk = (steps + WB - 1) / WB;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int l = 0; l < n_levels; ++l)
    for (int b = 0; b < nblk; ++b) {
      const double* slab = M + (size_t)l * (size_t)slab_stride;
      const int j0 = b * WB, cnt = steps - j0 < WB ? steps - j0 : WB;
      double acc[WB];
      for (int j = 0; j < WB; ++j) acc[j] = 0.0;
      const double w0 = omega0[l] + (double)j0 * spacing;
      const double wlo = w0, whi = w0 + (double)(WB - 1) * spacing;
      for (int P = 0; P < naux; ++P) {
        const double fac = ppm_fac[P];
        if (fac == 0.0) continue;
        const double Om = ppm_freq[P];
        const double* row = slab + (size_t)P * (size_t)ld;
        for (int m = 0; m < ntotal; ++m) {
          const double z = energies[m] - (m < n_occ ? Om : -Om);      /* pole position */
          const double a = fac * row[m] * row[m];
          if (z < wlo - 0.25 || z > whi + 0.25) {                     /* no frequency of the block is damped */
#pragma omp simd
            for (int j = 0; j < WB; ++j) acc[j] += a / (w0 + (double)j * spacing - z);
          } else {
            for (int j = 0; j < WB; ++j) acc[j] += a * stab_inv(w0 + (double)j * spacing - z);
          }
        }
      }
      for (int j = 0; j < cnt; ++j) values[(size_t)l * (size_t)steps + (size_t)(j0 + j)] = acc[j];
    }
}

/* out[l1*q + l2] = out[l2*q + l1] = Sigma_PPM::CalcCorrelationOffDiagElement(l1, l2, w[l1], w[l2]), zero diagonal */
void sigma_ppm_offdiag_all(const double* M, long long slab_stride, long long ld, int ntotal, int naux, int n_occ,
                           const double* energies, const double* ppm_freq, const double* ppm_fac, int q,
                           const double* w, double* out) {
  for (int i = 0; i < q * q; ++i) out[i] = 0.0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int l1 = 0; l1 < q; ++l1) {
    const double* s1 = M + (size_t)l1 * (size_t)slab_stride;
    for (int l2 = l1 + 1; l2 < q; ++l2) {
      const double* s2 = M + (size_t)l2 * (size_t)slab_stride;
      double tot = 0.0;
      for (int P = 0; P < naux; ++P) {
        const double fac = ppm_fac[P];
        if (fac == 0.0) continue;
        const double Om = ppm_freq[P];
        const double* r1 = s1 + (size_t)P * (size_t)ld;
        const double* r2 = s2 + (size_t)P * (size_t)ld;
        double acc = 0.0;
        for (int m = 0; m < ntotal; ++m) {
          const double sh = energies[m] - (m < n_occ ? Om : -Om);
          acc += r1[m] * r2[m] * (stab_inv(w[l1] - sh) + stab_inv(w[l2] - sh));
        }
        tot += fac * acc;
      }
      out[(size_t)l1 * q + l2] = out[(size_t)l2 * q + l1] = 0.5 * tot;
    }
  }
}

/* W[l][P][m] = fac_P g(w[l] - z_{P,m}) M[l][P][m]   (same layout as M, ld_w = ntotal) */
void sigma_ppm_weighted_slab(const double* M, long long slab_stride, long long ld, int ntotal, int naux, int n_occ,
                             const double* energies, const double* ppm_freq, const double* ppm_fac, int q,
                             const double* w, double* W) {
#pragma omp parallel for schedule(static) collapse(2)
  for (int l = 0; l < q; ++l)
    for (int P = 0; P < naux; ++P) {
      const double* row = M + (size_t)l * (size_t)slab_stride + (size_t)P * (size_t)ld;
      double* dst = W + ((size_t)l * (size_t)naux + (size_t)P) * (size_t)ntotal;
      const double fac = ppm_fac[P], Om = ppm_freq[P];
      for (int m = 0; m < ntotal; ++m)
        dst[m] = fac == 0.0 ? 0.0 : fac * stab_inv(w[l] - energies[m] + (m < n_occ ? Om : -Om)) * row[m];
    }
}
