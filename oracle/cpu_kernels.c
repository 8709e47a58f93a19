/* CPU oracle, C part -- TEST / BASELINE INFRASTRUCTURE ONLY (see gwbse_oracle.py header; PARITY UNPINNED).
 *
 * Plain C + OpenMP restatement of the loops of the GW-BSE path that are not GEMMs, so that the CPU baseline
 * of bench.py uses every host core for them the way the reference's OpenMP loops do:
 *   sigma_ppm_diag   Sigma_PPM::CalcCorrelationDiagElement evaluated for one gw level at n_omega frequencies
 *                    (upstream xtp/src/libxtp/gwbse/sigma_ppm.cc; the QP grid solver GW::SolveQP_Grid of gw.cc calls
 *                    it qp_grid_steps times per level; the reference parallelises over levels, here the
 *                    frequencies of one level are spread over the threads -- same work per core).
 *   unpack_symmetric packed lower triangle -> full symmetric matrix (host mirror of k_unpack_symmetric).
 * The GEMM-shaped stages go through numpy's OpenBLAS (oracle/cpu_reference.py).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file.
 */
#include <math.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static const double kFourPi = 12.566370614359172953850573533118;

/* Sigma_PPM::Stabilize: 1/x for |x| >= 0.25, 0.5 (1 - cos 4 pi x) / x below, 0 at x = 0 */
static inline double stabilized_inverse(double x) {
  const double ax = fabs(x);
  if (ax >= 0.25) return 1.0 / x;
  if (x == 0.0) return 0.0;
  return 0.5 * (1.0 - cos(kFourPi * x)) / x;
}

/* slab: [naux][ld] (P-major, level index m fastest: slab[P*ld + m] = M~[level](m, P)); out[n_omega] */
void sigma_ppm_diag(const double* slab, long long ld, int ntotal, int naux, int n_occ, const double* energies,
                    const double* ppm_freq, const double* ppm_fac, const double* omegas, int n_omega, double* out,
                    double* out_deriv) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int w = 0; w < n_omega; ++w) {
    const double om = omegas[w];
    double val = 0.0, der = 0.0;
    for (int P = 0; P < naux; ++P) {
      const double fac = ppm_fac[P];
      if (fac == 0.0) continue;
      const double Om = ppm_freq[P];
      const double* row = slab + (size_t)P * (size_t)ld;
      double v1 = 0.0, d1 = 0.0;
      for (int m = 0; m < ntotal; ++m) {
        const double g = stabilized_inverse(om - energies[m] + (m < n_occ ? Om : -Om));
        const double a = row[m] * row[m] * g;
        v1 += a;
        d1 -= a * g;
      }
      val += fac * v1;
      der += fac * d1;
    }
    out[w] = val;
    if (out_deriv) out_deriv[w] = der;
  }
}

void unpack_symmetric(const double* packed, int n, double* full, long long ld) {
#pragma omp parallel for schedule(static)
  for (int mu = 0; mu < n; ++mu) {
    const double* row = packed + (size_t)mu * (size_t)(mu + 1) / 2;
    for (int nu = 0; nu <= mu; ++nu) {
      full[(size_t)mu + (size_t)nu * (size_t)ld] = row[nu];
      full[(size_t)nu + (size_t)mu * (size_t)ld] = row[nu];
    }
  }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU baseline is meant to use every host core */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
