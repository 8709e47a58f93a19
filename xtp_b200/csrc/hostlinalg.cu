// Small dense symmetric eigenproblems on the host.  The Davidson solvers diagonalise a projected matrix of at most a
// few hundred rows every iteration (upstream DavidsonSolver uses Eigen::SelfAdjointEigenSolver on the host for the
// same step, davidsonsolver.cc); a cuSOLVER Dsyevd call of that size costs milliseconds of launch latency and three
// stream synchronisations, the host solve below tens of microseconds.
//
// Algorithm: Householder reduction to tridiagonal form followed by the implicit-shift QL iteration, both accumulating
// the transformation (the classical tred2 / tql2 pair of Wilkinson & Reinsch, Handbook for Automatic Computation II).
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "internal.h"

namespace xtpb {

namespace {

// V (n x n, column-major, ld = n): symmetric input -> orthogonal Q with Q^T A Q tridiagonal (d diagonal, e sub-diagonal
// in e[1..n-1]).
void tridiagonalise(int n, double* V, double* d, double* e) {
  auto v = [&](int i, int j) -> double& { return V[(size_t)i + (size_t)j * n]; };
  for (int j = 0; j < n; ++j) d[j] = v(n - 1, j);
  for (int i = n - 1; i > 0; --i) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; ++j) {
        d[j] = v(i - 1, j);
        v(i, j) = 0.0;
        v(j, i) = 0.0;
      }
    } else {
      for (int k = 0; k < i; ++k) {
        d[k] /= scale;
        h += d[k] * d[k];
      }
      double f = d[i - 1];
      double g = f > 0 ? -std::sqrt(h) : std::sqrt(h);
      e[i] = scale * g;
      h -= f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.0;
      for (int j = 0; j < i; ++j) {
        f = d[j];
        v(j, i) = f;
        g = e[j] + v(j, j) * f;
        for (int k = j + 1; k <= i - 1; ++k) {
          g += v(k, j) * d[k];
          e[k] += v(k, j) * f;
        }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; ++j) {
        e[j] /= h;
        f += e[j] * d[j];
      }
      const double hh = f / (h + h);
      for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
      for (int j = 0; j < i; ++j) {
        f = d[j];
        g = e[j];
        for (int k = j; k <= i - 1; ++k) v(k, j) -= (f * e[k] + g * d[k]);
        d[j] = v(i - 1, j);
        v(i, j) = 0.0;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; ++i) {
    v(n - 1, i) = v(i, i);
    v(i, i) = 1.0;
    const double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; ++k) d[k] = v(k, i + 1) / h;
      for (int j = 0; j <= i; ++j) {
        double g = 0.0;
        for (int k = 0; k <= i; ++k) g += v(k, i + 1) * v(k, j);
        for (int k = 0; k <= i; ++k) v(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; ++k) v(k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; ++j) {
    d[j] = v(n - 1, j);
    v(n - 1, j) = 0.0;
  }
  v(n - 1, n - 1) = 1.0;
  e[0] = 0.0;
}

// implicit QL on the tridiagonal (d, e), rotations accumulated into V; returns false when an eigenvalue needs more
// than 60 sweeps
bool ql_implicit(int n, double* V, double* d, double* e) {
  auto v = [&](int i, int j) -> double& { return V[(size_t)i + (size_t)j * n]; };
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = std::ldexp(1.0, -52);
  for (int l = 0; l < n; ++l) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n) {
      if (std::fabs(e[m]) <= eps * tst1) break;
      ++m;
    }
    if (m == n) m = n - 1;
    if (m > l) {
      int iter = 0;
      do {
        if (++iter > 60) return false;
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = std::hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; ++i) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c;
        const double el1 = e[l + 1];
        double s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; --i) {
          c3 = c2;
          c2 = c;
          s2 = s;
          g = c * e[i];
          h = c * p;
          r = std::hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int k = 0; k < n; ++k) {
            h = v(k, i + 1);
            v(k, i + 1) = s * v(k, i) + c * h;
            v(k, i) = c * v(k, i) - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1);
    }
    d[l] += f;
    e[l] = 0.0;
  }
  return true;
}

}  // namespace

// A (n x n col-major, ld = n, symmetric; the lower triangle is read) <- eigenvectors (columns), w <- ascending
// eigenvalues.  Returns false when the QL iteration does not converge (the caller falls back to cuSOLVER).
bool host_eigh(int n, double* A, double* w) {
  if (n <= 0) return true;
  if (n == 1) {
    w[0] = A[0];
    A[0] = 1.0;
    return true;
  }
  for (int j = 0; j < n; ++j)          // the reduction reads the full matrix: mirror the lower triangle
    for (int i = j + 1; i < n; ++i) A[(size_t)j + (size_t)i * n] = A[(size_t)i + (size_t)j * n];
  std::vector<double> e((size_t)n);
  tridiagonalise(n, A, w, e.data());
  if (!ql_implicit(n, A, w, e.data())) return false;
  // ascending order (selection on an index permutation, then permute the columns)
  std::vector<int> order((size_t)n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return w[a] < w[b]; });
  std::vector<double> V((size_t)n * n), ws((size_t)n);
  for (int j = 0; j < n; ++j) {
    ws[j] = w[order[j]];
    std::copy(A + (size_t)order[j] * n, A + (size_t)(order[j] + 1) * n, V.begin() + (size_t)j * n);
  }
  std::copy(ws.begin(), ws.end(), w);
  std::copy(V.begin(), V.end(), A);
  return true;
}

}  // namespace xtpb

// ---------------------------------------------------------------------------------------------------------------
// Anderson mixing of the evGW quasiparticle energies (upstream xtp/src/libxtp/anderson_mixing.cc, used by
// GW::CalculateGWPerturbation when gw_mixing_order > 1).  History of the last `order` input / output vectors; the new
// guess minimises |sum_j c_j (out_j - in_j)| over the affine combinations of the history and is then damped:
//   x = alpha * Out_mixed + (1 - alpha) * In_mixed.
// order 1 degenerates to linear mixing of the last pair.
namespace xtpb {

void Anderson::configure(int order, double alpha) {
  XTPB_REQUIRE(order >= 1 && alpha > 0.0 && alpha <= 1.0, "Anderson mixing needs order >= 1 and 0 < alpha <= 1");
  order_ = order;
  alpha_ = alpha;
  input_.clear();
  output_.clear();
}
void Anderson::update_input(const std::vector<double>& x) {
  if ((int)input_.size() > order_ - 1) input_.erase(input_.begin());
  input_.push_back(x);
}
void Anderson::update_output(const std::vector<double>& x) {
  if ((int)output_.size() > order_ - 1) output_.erase(output_.begin());
  output_.push_back(x);
}

// minimum-norm least-squares solution of the symmetric positive semi-definite h x h system A c = b through the
// eigendecomposition A = U diag(l) U^T: c = sum_i u_i (u_i . b) / l_i over l_i > 1e-12 l_max.  (Upstream solves with
// Eigen's fullPivHouseholderQr; for a full-rank history the two agree, for a rank-deficient one -- repeated residuals
// -- the minimum-norm choice is the reproducible one.)
static std::vector<double> solve_psd_min_norm(int h, std::vector<double> A, const std::vector<double>& b) {
  std::vector<double> lam((size_t)h), c((size_t)h, 0.0);
  XTPB_REQUIRE(host_eigh(h, A.data(), lam.data()), "Anderson mixing: eigensolver did not converge");
  const double lmax = std::max(std::fabs(lam.front()), std::fabs(lam.back()));
  for (int i = 0; i < h; ++i) {
    if (!(lam[i] > 1e-12 * lmax)) continue;
    double ub = 0.0;
    for (int k = 0; k < h; ++k) ub += A[(size_t)k + (size_t)i * h] * b[k];
    for (int k = 0; k < h; ++k) c[k] += A[(size_t)k + (size_t)i * h] * ub / lam[i];
  }
  return c;
}

std::vector<double> Anderson::mix_history() const {
  XTPB_REQUIRE(!output_.empty() && output_.size() == input_.size(), "Anderson mixing: input/output history out of step");
  const int iteration = (int)output_.size(), used = iteration - 1;
  std::vector<double> out = output_.back(), in = input_.back();
  const size_t n = out.size();
  if (iteration > 1 && order_ > 1) {
    std::vector<double> dN(n);
    for (size_t i = 0; i < n; ++i) dN[i] = out[i] - in[i];
    // D_m = DeltaN - (out_{used-m} - in_{used-m}), m = 1 .. used
    std::vector<std::vector<double>> D((size_t)used, std::vector<double>(n));
    for (int m = 1; m <= used; ++m)
      for (size_t i = 0; i < n; ++i) D[m - 1][i] = dN[i] - output_[used - m][i] + input_[used - m][i];
    std::vector<double> A((size_t)used * used), c((size_t)used);
    for (int m = 0; m < used; ++m) {
      c[m] = std::inner_product(D[m].begin(), D[m].end(), dN.begin(), 0.0);
      for (int j = 0; j < used; ++j) A[(size_t)m + (size_t)j * used] = std::inner_product(D[m].begin(), D[m].end(), D[j].begin(), 0.0);
    }
    const std::vector<double> coef = solve_psd_min_norm(used, A, c);
    for (int k = 1; k <= used; ++k)
      for (size_t i = 0; i < n; ++i) {
        out[i] += coef[k - 1] * (output_[used - k][i] - output_[used][i]);
        in[i] += coef[k - 1] * (input_[used - k][i] - input_[used][i]);
      }
  }
  std::vector<double> x(n);
  for (size_t i = 0; i < n; ++i) x[i] = alpha_ * out[i] + (1.0 - alpha_) * in[i];
  return x;
}

}  // namespace xtpb
