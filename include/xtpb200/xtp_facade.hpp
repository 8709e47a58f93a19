// Header-only C++17 facade over libxtpb200's C ABI (xtpb200.h) that re-creates the class names and member
// functions of VOTCA-XTP's GW-BSE path, so that gwbse.cc-style driver code compiles against it unchanged in
// shape:  TCMatrix_gwbse, RPA, GW (+ the Sigma_base calls it owns), BSE, BSE_OPERATOR<cqp,cx,cd,cd2> and its
// typedefs, MatrixFreeOperator, DavidsonSolver.
//
// Upstream files mirrored (votca/votca, xtp/ subtree; no line numbers: /root/reference/README.md:1 is a redirect
// stub):
//   xtp/include/votca/xtp/threecenter.h, rpa.h, gw.h, sigma_base.h, bse.h, bse_operator.h,
//   matrixfreeoperator.h, davidsonsolver.h
//
// The reference's matrices are Eigen::MatrixXd / Eigen::VectorXd.  Eigen is not a dependency of this header: every
// class is a template over a `Matrix` type that offers  Matrix(rows, cols), rows(), cols(), data()  with
// column-major storage, and a `Vector` type with  Vector(n), size(), data().  Eigen::MatrixXd / Eigen::VectorXd
// satisfy both as they are; `xtpb200::DenseMatrix` / `DenseVector` below are minimal stand-ins.
//
// Error behaviour: the reference throws std::runtime_error; so does this facade (message from xtpb_last_error()).
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <fstream>
#include <string>
#include <utility>
#include <array>
#include <vector>

#include "xtpb200.h"

namespace xtpb200 {

using Index = xtpb_index;   // Eigen::Index on LP64

inline void check(int status) {
  if (status != 0) throw std::runtime_error(xtpb_last_error());
}

// ---- minimal column-major containers (used when Eigen is not around)
class DenseVector {
 public:
  DenseVector() = default;
  explicit DenseVector(Index n) : v_(static_cast<size_t>(n), 0.0) {}
  Index size() const { return static_cast<Index>(v_.size()); }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double& operator()(Index i) { return v_[static_cast<size_t>(i)]; }
  double operator()(Index i) const { return v_[static_cast<size_t>(i)]; }

 private:
  std::vector<double> v_;
};
class DenseMatrix {
 public:
  DenseMatrix() = default;
  DenseMatrix(Index r, Index c) : r_(r), c_(c), v_(static_cast<size_t>(r * c), 0.0) {}
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double& operator()(Index i, Index j) { return v_[static_cast<size_t>(i + j * r_)]; }
  double operator()(Index i, Index j) const { return v_[static_cast<size_t>(i + j * r_)]; }

 private:
  Index r_ = 0, c_ = 0;
  std::vector<double> v_;
};

// ---- one CUDA device + stream; stands where the reference passes a Logger / OpenMP_CUDA around
class Context {
 public:
  explicit Context(int device = 0) { check(xtpb_ctx_create(device, &h_)); }
  ~Context() { if (h_) xtpb_ctx_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  xtpb_ctx* handle() const { return h_; }
  void sync() const { check(xtpb_ctx_sync(h_)); }
  // multi-GPU (one process per GPU): rank 0 calls UniqueId(), the host program distributes the 128 bytes, every
  // rank calls JoinCommunicator before creating its TCMatrix_gwbse.  No counterpart upstream (OpenMP threads there).
  static std::array<char, 128> UniqueId() {
    std::array<char, 128> id{};
    check(xtpb_comm_unique_id(id.data()));
    return id;
  }
  void JoinCommunicator(const std::array<char, 128>& id, int rank, int world) {
    check(xtpb_ctx_comm_init(h_, id.data(), rank, world));
  }

 private:
  xtpb_ctx* h_ = nullptr;
};

// ---- GaussianQuadrature (gaussian_quadrature.h): scaled nodes/weights of the Sigma_CDA imaginary-axis integral
class GaussianQuadrature {
 public:
  struct options { Index order = 12; int qptype = XTPB_QUAD_LEGENDRE; };
  void configure(const options& opt) {
    points_.assign(static_cast<std::size_t>(2 * opt.order + 2), 0.0);
    weights_ = points_;
    Index n = 0;
    check(xtpb_gaussian_quadrature(opt.qptype, opt.order, points_.data(), weights_.data(), &n));
    points_.resize(static_cast<std::size_t>(n));
    weights_.resize(static_cast<std::size_t>(n));
  }
  Index Order() const { return static_cast<Index>(points_.size()); }
  double ScaledPoint(Index i) const { return points_[static_cast<std::size_t>(i)]; }
  double ScaledWeight(Index i) const { return weights_[static_cast<std::size_t>(i)]; }

 private:
  std::vector<double> points_, weights_;
};

// ---- TCMatrix_gwbse (threecenter.h / threecenter_gwbse.cc)
template <class Matrix = DenseMatrix>
class TCMatrix_gwbse {
 public:
  explicit TCMatrix_gwbse(Context& ctx) : ctx_(ctx) {}
  ~TCMatrix_gwbse() { if (h_) xtpb_tc_destroy(h_); }
  TCMatrix_gwbse(const TCMatrix_gwbse&) = delete;
  TCMatrix_gwbse& operator=(const TCMatrix_gwbse&) = delete;

  void Initialize(Index basissize, Index mmin, Index mmax, Index nmin, Index nmax) {
    if (h_) { xtpb_tc_destroy(h_); h_ = nullptr; }
    check(xtpb_tc_create(ctx_.handle(), basissize, mmin, mmax, nmin, nmax, &h_));
    aux_ = basissize; mmin_ = mmin; mmax_ = mmax; nmin_ = nmin; nmax_ = nmax;
  }
  Index auxsize() const { return aux_; }
  Index msize() const { return mmax_ - mmin_ + 1; }
  Index nsize() const { return nmax_ - nmin_ + 1; }
  Index get_mmin() const { return mmin_; }
  Index get_mmax() const { return mmax_; }
  Index get_nmin() const { return nmin_; }
  Index get_nmax() const { return nmax_; }

  // Fill3cMO, streamed: the caller's integral loop (libint2 in the reference) hands over blocks of symmetric
  // AO slices (P|mu nu), full or packed-lower.
  void Fill3cMO_begin(const Matrix& dft_orbitals) {
    check(xtpb_tc_fill_begin(h_, dft_orbitals.rows(), dft_orbitals.data(), dft_orbitals.rows()));
  }
  void Fill3cMO_block(Index P0, Index nP, const double* ao3c, Index ld_ao) {
    check(xtpb_tc_fill_block(h_, P0, nP, ao3c, ld_ao));
  }
  void Fill3cMO_block_packed(Index P0, Index nP, const double* ao3c_packed) {
    check(xtpb_tc_fill_block_packed(h_, P0, nP, ao3c_packed));
  }
  // Optional, BEFORE the Fill3cMO calls: announce the matrices of the metric step (on the eigensolver path its first
  // decomposition then runs on the library's helper thread underneath the MO transform); ApplyCoulombMetric with the
  // SAME matrices (which must stay alive and unmoved until then) consumes the hint.
  void CoulombMetricBegin(const Matrix& aux_coulomb, const Matrix* aux_overlap = nullptr) {
    check(xtpb_tc_coulomb_metric_begin(h_, aux_coulomb.data(), aux_coulomb.rows(),
                                       aux_overlap ? aux_overlap->data() : nullptr,
                                       aux_overlap ? aux_overlap->rows() : 0));
  }
  // Optional, after FillBegin on a single rank: Sigma_PPM::PrepareScreening will run with these RPA input energies
  // (rpamin..rpamax); the plasmon-pole model's two epsilon matrices are then accumulated while the aux blocks arrive
  // (xtpb_tc_ppm_prefetch_begin; pays only when the fill waits for a slow host link).
  template <class Vector>
  void PpmPrefetchBegin(const Vector& rpa_input_energies, Index homo, double eta = 1e-3) {
    check(xtpb_tc_ppm_prefetch_begin(h_, rpa_input_energies.data(), homo, eta));
  }
  // which path the metric step took so far: {Cholesky factor, reference eigensolver construction}
  std::pair<Index, Index> MetricPathInfo() const {
    Index chol = 0, eig = 0;
    check(xtpb_tc_metric_path_info(h_, &chol, &eig));
    return {chol, eig};
  }
  // second half of Fill: Pseudo_InvSqrt_GWBSE(auxoverlap, 5e-7) + MultiplyRightWithAuxMatrix
  Index ApplyCoulombMetric(const Matrix& aux_coulomb, const Matrix* aux_overlap = nullptr, double etol = 5e-7) {
    Index removed = 0;
    check(xtpb_tc_apply_coulomb_metric(h_, aux_coulomb.data(), aux_coulomb.rows(),
                                       aux_overlap ? aux_overlap->data() : nullptr,
                                       aux_overlap ? aux_overlap->rows() : 0, etol, &removed));
    removedfunctions_ = removed;
    return removed;
  }
  Index Removedfunctions() const { return removedfunctions_; }
  void MultiplyRightWithAuxMatrix(const Matrix& matrix) {
    check(xtpb_tc_multiply_right_with_aux_matrix(h_, matrix.data(), matrix.rows()));
  }
  // operator[](i): the reference returns a reference into host storage; the tensor lives in HBM here, so this
  // returns a host copy of slab i (nsize x auxsize).
  Matrix operator[](Index i) const {
    Matrix slab(nsize(), auxsize());
    check(xtpb_tc_get_slab(h_, i, slab.data()));
    return slab;
  }
  xtpb_tc* handle() const { return h_; }
  Context& context() const { return ctx_; }

 private:
  Context& ctx_;
  xtpb_tc* h_ = nullptr;
  Index aux_ = 0, mmin_ = 0, mmax_ = 0, nmin_ = 0, nmax_ = 0, removedfunctions_ = 0;
};

// ---- RPA (rpa.h / rpa.cc)
template <class Matrix = DenseMatrix, class Vector = DenseVector>
class RPA {
 public:
  explicit RPA(const TCMatrix_gwbse<Matrix>& Mmn) : Mmn_(Mmn) {}
  void configure(Index homo, Index rpamin, Index rpamax) { homo_ = homo; rpamin_ = rpamin; rpamax_ = rpamax; }
  double getEta() const { return eta_; }
  void setRPAInputEnergies(const Vector& e) { energies_ = e; }
  const Vector& getRPAInputEnergies() const { return energies_; }
  // RPA::UpdateRPAInputEnergies(dftenergies, gwaenergies, qpmin)
  void UpdateRPAInputEnergies(const Vector& dft, const Vector& gwa, Index qpmin) {
    const Index rpatotal = rpamax_ - rpamin_ + 1, gwsize = gwa.size(), lumo = homo_ + 1, qpmax = qpmin + gwsize - 1;
    Vector e(rpatotal);
    for (Index i = 0; i < rpatotal; ++i) e.data()[i] = dft.data()[rpamin_ + i];
    for (Index i = 0; i < gwsize; ++i) e.data()[qpmin - rpamin_ + i] = gwa.data()[i];
    const double dftgap = dft.data()[lumo] - dft.data()[homo_];
    const double qpgap = gwa.data()[lumo - qpmin] - gwa.data()[homo_ - qpmin];
    const double shift = qpgap - dftgap;
    for (Index i = qpmax + 1 - rpamin_; i < rpatotal; ++i) e.data()[i] += shift;
    for (Index i = 0; i < qpmin - rpamin_; ++i) e.data()[i] -= shift;
    energies_ = e;
  }
  Matrix calculate_epsilon_i(double frequency) const { return epsilon(frequency, 1); }
  Matrix calculate_epsilon_r(double frequency) const { return epsilon(frequency, 0); }

 private:
  Matrix epsilon(double w, int imag) const {
    Matrix eps(Mmn_.auxsize(), Mmn_.auxsize());
    check(xtpb_rpa_epsilon(Mmn_.handle(), energies_.data(), homo_, rpamin_, rpamax_, eta_, &w, 1, imag, eps.data()));
    return eps;
  }
  const TCMatrix_gwbse<Matrix>& Mmn_;
  Vector energies_;
  Index homo_ = 0, rpamin_ = 0, rpamax_ = 0;
  double eta_ = 1e-3;
};

// ---- GW + Sigma_base (gw.h, sigma_base.h; Sigma_PPM / Sigma_Exact / Sigma_CDA chosen by opt.sigma_integration)
template <class Matrix = DenseMatrix, class Vector = DenseVector>
class GW {
 public:
  using options = xtpb_gw_options;
  static options default_options() { options o; xtpb_gw_options_default(&o); return o; }

  GW(TCMatrix_gwbse<Matrix>& Mmn, const Matrix& vxc, const Vector& dft_energies)
      : Mmn_(Mmn), vxc_(vxc), dft_energies_(dft_energies) {}
  ~GW() { if (h_) xtpb_gw_destroy(h_); }
  GW(const GW&) = delete;
  GW& operator=(const GW&) = delete;

  void configure(const options& opt) {
    if (h_) { xtpb_gw_destroy(h_); h_ = nullptr; }
    opt_ = opt;
    qptotal_ = opt.qpmax - opt.qpmin + 1;
    rpatotal_ = opt.rpamax - opt.rpamin + 1;
    check(xtpb_gw_create(Mmn_.context().handle(), Mmn_.handle(), &opt_, vxc_.data(), vxc_.rows(), dft_energies_.data(),
                         dft_energies_.size(), &h_));
  }
  void CalculateGWPerturbation() { check(xtpb_gw_calculate_gw_perturbation(h_)); }
  void CalculateHQP() { check(xtpb_gw_calculate_hqp(h_)); }
  Vector getGWAResults() const { Vector v(qptotal_); check(xtpb_gw_get_gwa_results(h_, v.data())); return v; }
  Matrix getHQP() const { Matrix m(qptotal_, qptotal_); check(xtpb_gw_get_hqp(h_, m.data())); return m; }
  Vector RPAInputEnergies() const {
    Vector v(rpatotal_);
    check(xtpb_gw_get_rpa_input_energies(h_, v.data()));
    return v;
  }
  // Eigen::SelfAdjointEigenSolver<MatrixXd> in the reference: eigenvalues + eigenvectors of Hqp
  std::pair<Vector, Matrix> DiagonalizeQPHamiltonian() const {
    Vector w(qptotal_);
    Matrix v(qptotal_, qptotal_);
    check(xtpb_gw_diagonalize_qp_hamiltonian(h_, w.data(), v.data()));
    return {w, v};
  }
  // Sigma_base interface (the GW object owns its Sigma in the reference as std::unique_ptr<Sigma_base>)
  void PrepareScreening() { check(xtpb_gw_prepare_screening(h_)); }
  Matrix CalcExchangeMatrix() const {
    Matrix m(qptotal_, qptotal_);
    check(xtpb_gw_sigma_exchange(h_, m.data()));
    return m;
  }
  Vector CalcCorrelationDiag(const Vector& frequencies) const {
    Vector v(qptotal_);
    check(xtpb_gw_sigma_c_diag(h_, frequencies.data(), v.data()));
    return v;
  }
  Matrix CalcCorrelationOffDiag(const Vector& frequencies) const {
    Matrix m(qptotal_, qptotal_);
    check(xtpb_gw_sigma_c_offdiag(h_, frequencies.data(), m.data()));
    return m;
  }
  double CalcCorrelationDiagElement(Index gw_level, double frequency) const {
    double v = 0.0;
    check(xtpb_gw_sigma_c_diag_elements(h_, 1, &gw_level, &frequency, &v, nullptr));
    return v;
  }
  double CalcCorrelationDiagElementDerivative(Index gw_level, double frequency) const {
    double v = 0.0, d = 0.0;
    check(xtpb_gw_sigma_c_diag_elements(h_, 1, &gw_level, &frequency, &v, &d));
    return d;
  }
  // batched form (north_star's "CalcCorrelationDiagElements"): n (level, frequency) pairs in one launch
  void CalcCorrelationDiagElements(Index n, const Index* levels, const double* frequencies, double* values,
                                   double* derivatives = nullptr) const {
    check(xtpb_gw_sigma_c_diag_elements(h_, n, levels, frequencies, values, derivatives));
  }
  // Sigma_c of every gw level on its QP grid (what GW::SolveQP_Grid scans and GW::PlotSigma tabulates):
  // result(level, j) = Sigma_c(level, center[level] + (j - (steps-1)/2) * qp_grid_spacing)
  Matrix CalcCorrelationGrid(const Vector& center_frequencies) const {
    Matrix rowmajor(opt_.qp_grid_steps, qptotal_);      // column `level` holds that level's grid (contiguous)
    check(xtpb_gw_sigma_c_grid(h_, center_frequencies.data(), rowmajor.data()));
    return rowmajor;
  }
  // GW::PlotSigma(filename, steps, spacing, states): `states` are gw-level indices (0 = qpmin); writes upstream's
  // tab-separated table (frequency / Sigma_c + e_KS + Sigma_x - Vxc column pair per state) and returns it
  Matrix PlotSigma(const std::string& filename, Index steps, double spacing, const std::vector<Index>& states) const {
    Matrix table(steps, 2 * static_cast<Index>(states.size()));
    check(xtpb_gw_plot_sigma(h_, steps, spacing, static_cast<Index>(states.size()), states.data(), table.data()));
    if (!filename.empty()) {
      std::ofstream out(filename);
      for (std::size_t i = 0; i < states.size(); ++i)
        out << (i ? "\t" : "") << "#frequency_" << opt_.qpmin + states[i] << "\tSigma_c_" << opt_.qpmin + states[i];
      out << "\n";
      out.precision(12);
      for (Index gp = 0; gp < steps; ++gp) {
        for (Index c = 0; c < table.cols(); ++c) out << (c ? "\t" : "") << table.data()[gp + c * steps];
        out << "\n";
      }
    }
    return table;
  }
  xtpb_gw* handle() const { return h_; }

 private:
  TCMatrix_gwbse<Matrix>& Mmn_;
  Matrix vxc_;
  Vector dft_energies_;
  options opt_{};
  Index qptotal_ = 0, rpatotal_ = 0;
  xtpb_gw* h_ = nullptr;
};

// ---- MatrixFreeOperator (matrixfreeoperator.h)
template <class Matrix = DenseMatrix, class Vector = DenseVector>
class MatrixFreeOperator {
 public:
  MatrixFreeOperator() = default;
  explicit MatrixFreeOperator(xtpb_op* h) : h_(h) { check(xtpb_op_size(h_, &size_)); }
  virtual ~MatrixFreeOperator() { if (h_) xtpb_op_destroy(h_); }
  MatrixFreeOperator(const MatrixFreeOperator&) = delete;
  MatrixFreeOperator& operator=(const MatrixFreeOperator&) = delete;
  Index rows() const { return size_; }
  Index cols() const { return size_; }
  Index size() const { return size_; }
  virtual Matrix matmul(const Matrix& input) const {
    Matrix out(size_, input.cols());
    check(xtpb_op_matmul(h_, input.data(), input.rows(), input.cols(), out.data(), size_));
    return out;
  }
  virtual Vector diagonal() const {
    Vector d(size_);
    check(xtpb_op_diagonal(h_, d.data()));
    return d;
  }
  Matrix get_full_matrix() const {
    Matrix H(size_, size_);
    check(xtpb_op_get_full_matrix(h_, H.data(), size_));
    return H;
  }
  xtpb_op* handle() const { return h_; }

 protected:
  void adopt(xtpb_op* h) { h_ = h; check(xtpb_op_size(h_, &size_)); }
  xtpb_op* h_ = nullptr;
  Index size_ = 0;
};

// ---- BSE_OPERATOR<cqp,cx,cd,cd2> (bse_operator.h / bse_operator.cc)
struct BSEOperator_Options {
  Index homo, rpamin, qpmin, vmin, cmax;
};
template <Index cqp, Index cx, Index cd, Index cd2, class Matrix = DenseMatrix, class Vector = DenseVector>
class BSE_OPERATOR final : public MatrixFreeOperator<Matrix, Vector> {
 public:
  BSE_OPERATOR(const Vector& epsilon_0_inv, const TCMatrix_gwbse<Matrix>& Mmn, const Matrix& Hqp)
      : eps_(epsilon_0_inv), Mmn_(Mmn), Hqp_(Hqp) {}
  void configure(BSEOperator_Options opt) {
    xtpb_op* h = nullptr;
    check(xtpb_bse_operator_create_raw(Mmn_.context().handle(), Mmn_.handle(), opt.homo, opt.rpamin, opt.vmin, opt.cmax,
                                       eps_.data(), Hqp_.data(), Hqp_.rows(), (int)cqp, (int)cx, (int)cd, (int)cd2, &h));
    this->adopt(h);
  }

 private:
  Vector eps_;
  const TCMatrix_gwbse<Matrix>& Mmn_;
  Matrix Hqp_;
};
template <class M = DenseMatrix, class V = DenseVector> using SingletOperator_TDA = BSE_OPERATOR<1, 2, 1, 0, M, V>;
template <class M = DenseMatrix, class V = DenseVector> using TripletOperator_TDA = BSE_OPERATOR<1, 0, 1, 0, M, V>;
template <class M = DenseMatrix, class V = DenseVector> using SingletOperator_BTDA_B = BSE_OPERATOR<0, 2, 0, 1, M, V>;
template <class M = DenseMatrix, class V = DenseVector> using TripletOperator_BTDA_B = BSE_OPERATOR<0, 0, 0, 1, M, V>;
template <class M = DenseMatrix, class V = DenseVector> using HxOperator = BSE_OPERATOR<0, 1, 0, 0, M, V>;
template <class M = DenseMatrix, class V = DenseVector> using HdOperator = BSE_OPERATOR<0, 0, 1, 0, M, V>;
template <class M = DenseMatrix, class V = DenseVector> using Hd2Operator = BSE_OPERATOR<0, 0, 0, 1, M, V>;
template <class M = DenseMatrix, class V = DenseVector> using HqpOperator = BSE_OPERATOR<1, 0, 0, 0, M, V>;

// ---- DavidsonSolver (davidsonsolver.h / davidsonsolver.cc)
template <class Matrix = DenseMatrix, class Vector = DenseVector>
class DavidsonSolver {
 public:
  DavidsonSolver() { xtpb_davidson_options_default(&opt_); }
  void set_iter_max(Index n) { opt_.iter_max = n; }
  void set_max_search_space(Index n) { opt_.max_search_space = n; }
  void set_tolerance(const std::string& tol) {
    if (tol == "loose") opt_.tolerance = 1e-3;
    else if (tol == "normal") opt_.tolerance = 1e-4;
    else if (tol == "strict") opt_.tolerance = 1e-5;
    else if (tol == "lapack") opt_.tolerance = 1e-9;
    else throw std::runtime_error(tol + " is not a valid Davidson tolerance");
  }
  void set_correction(const std::string& method) {
    if (method == "DPR") opt_.correction = XTPB_DAVIDSON_DPR;
    else if (method == "OLSEN") opt_.correction = XTPB_DAVIDSON_OLSEN;
    else throw std::runtime_error(method + " is not a valid Davidson correction method");
  }
  void set_size_update(const std::string& update) {
    if (update == "min") opt_.size_update = XTPB_UPDATE_MIN;
    else if (update == "safe") opt_.size_update = XTPB_UPDATE_SAFE;
    else if (update == "max") opt_.size_update = XTPB_UPDATE_MAX;
    else throw std::runtime_error(update + " is not a valid Davidson update option");
  }
  void set_matrix_type(const std::string& mt) {
    if (mt != "SYMM") throw std::runtime_error("only SYMM Davidson problems are supported by libxtpb200 in this build");
  }
  template <class Operator>
  void solve(const Operator& A, Index neigen, Index size_initial_guess = 0) {
    opt_.size_initial_guess = size_initial_guess;
    evals_ = Vector(neigen);
    evecs_ = Matrix(A.rows(), neigen);
    check(xtpb_davidson_solve(A.handle(), neigen, &opt_, evals_.data(), evecs_.data(), A.rows(), &info_, &iterations_));
  }
  const Vector& eigenvalues() const { return evals_; }
  const Matrix& eigenvectors() const { return evecs_; }
  int info() const { return info_; }           // 0 = Eigen::Success, 1 = Eigen::NoConvergence
  Index num_iterations() const { return iterations_; }

 private:
  xtpb_davidson_options opt_{};
  Vector evals_;
  Matrix evecs_;
  int info_ = 1;
  Index iterations_ = 0;
};

// ---- BSE (bse.h / bse.cc): configure = AdjustHqpSize + SetupDirectInteractionOperator
template <class Matrix = DenseMatrix, class Vector = DenseVector>
class BSE {
 public:
  using options = xtpb_bse_options;
  explicit BSE(TCMatrix_gwbse<Matrix>& Mmn) : Mmn_(Mmn) {}
  ~BSE() { if (h_) xtpb_bse_destroy(h_); }
  BSE(const BSE&) = delete;
  BSE& operator=(const BSE&) = delete;
  void configure(const options& opt, const Vector& RPAInputEnergies, const Matrix& Hqp_in) {
    if (h_) { xtpb_bse_destroy(h_); h_ = nullptr; }
    opt_ = opt;
    check(xtpb_bse_create(Mmn_.context().handle(), Mmn_.handle(), &opt_, RPAInputEnergies.data(), Hqp_in.data(),
                          Hqp_in.rows(), 0, &h_));
  }
  Vector epsilon_0_inv() const {
    Vector v(Mmn_.auxsize());
    check(xtpb_bse_get_epsilon_0_inv(h_, v.data()));
    return v;
  }
  struct Result { Vector energies; Matrix eigenvectors; int info; Index iterations; };
  Result Solve_singlets() const { return solve(1, 2, 1, 0); }     // TDA
  Result Solve_triplets() const { return solve(1, 0, 1, 0); }     // TDA
  // useTDA = false (upstream Solve_nonhermitian_Davidson): X = coefficients, Y = coefficients_AR, X^T X - Y^T Y = 1
  struct FullResult { Vector energies; Matrix X, Y; int info; Index iterations; };
  FullResult Solve_singlets_BTDA() const { return solve_full(1); }
  FullResult Solve_triplets_BTDA() const { return solve_full(0); }
  // BSE::CalcCoupledTransition_Dipoles: 3 x n_states (column s = dipole of state s); ao_dipoles = x, y, z matrices
  Matrix CalcCoupledTransition_Dipoles(const Matrix& dft_orbitals, const std::array<Matrix, 3>& ao_dipoles,
                                       const Matrix& X, const Matrix* Y = nullptr) const {
    const Index nb = dft_orbitals.rows();
    std::vector<double> r(static_cast<std::size_t>(3 * nb * nb));
    for (int i = 0; i < 3; ++i)
      std::copy(ao_dipoles[i].data(), ao_dipoles[i].data() + nb * nb, r.begin() + static_cast<std::size_t>(i) * nb * nb);
    Matrix d(3, X.cols());
    check(xtpb_bse_transition_dipoles(h_, nb, dft_orbitals.data(), nb, r.data(), X.cols(), X.data(),
                                      Y ? Y->data() : nullptr, X.rows(), d.data()));
    return d;
  }
  // BSE::Perturbative_DynamicalScreening(type, orb): dynamically screened excitation energies of the given states
  // (X = coefficients, Y = coefficients_AR or nullptr for the TDA); upstream option names and defaults
  Vector Perturbative_DynamicalScreening(const Vector& energies, const Matrix& X, const Matrix* Y = nullptr,
                                         Index max_dyn_iter = 10, double dyn_tolerance = 1e-5) const {
    Vector dyn(energies.size());
    check(xtpb_bse_perturbative_dynamical_screening(h_, energies.size(), energies.data(), X.data(),
                                                    Y ? Y->data() : nullptr, X.rows(), max_dyn_iter, dyn_tolerance,
                                                    dyn.data(), nullptr));
    return dyn;
  }
  // Orbitals::Oscillatorstrengths
  static Vector Oscillatorstrengths(const Vector& energies, const Matrix& dipoles) {
    Vector f(energies.size());
    check(xtpb_oscillator_strengths(energies.size(), energies.data(), dipoles.data(), f.data()));
    return f;
  }

 private:
  Result solve(int cqp, int cx, int cd, int cd2) const {
    xtpb_op* op = nullptr;
    check(xtpb_bse_operator_create(h_, cqp, cx, cd, cd2, &op));
    MatrixFreeOperator<Matrix, Vector> H(op);
    DavidsonSolver<Matrix, Vector> ds;
    ds.set_max_search_space(10 * opt_.nmax);
    ds.solve(H, opt_.nmax);
    return Result{ds.eigenvalues(), ds.eigenvectors(), ds.info(), ds.num_iterations()};
  }
  FullResult solve_full(int singlet) const {
    xtpb_davidson_options o;
    xtpb_davidson_options_default(&o);
    o.max_search_space = 10 * opt_.nmax;
    const Index n = (opt_.homo - opt_.vmin + 1) * (opt_.cmax - opt_.homo);
    FullResult r{Vector(opt_.nmax), Matrix(n, opt_.nmax), Matrix(n, opt_.nmax), 1, 0};
    check(xtpb_bse_solve_btda(h_, singlet, &o, r.energies.data(), r.X.data(), r.Y.data(), n, &r.info, &r.iterations));
    return r;
  }
  TCMatrix_gwbse<Matrix>& Mmn_;
  options opt_{};
  xtpb_bse* h_ = nullptr;
};

// ---- GWBSE::Initialize level ranges (gwbse.cc `ranges` option): default / factor / explicit / full + clamps
inline xtpb_gwbse_ranges LevelRanges(int mode, Index n_levels, Index n_occ, double rpamax = 0, double qpmin = 0,
                                     double qpmax = 0, double bsemin = 0, double bsemax = 0, Index n_core_ignored = 0) {
  xtpb_gwbse_range_options o{mode, n_levels, n_occ, n_core_ignored, rpamax, qpmin, qpmax, bsemin, bsemax};
  xtpb_gwbse_ranges r{};
  check(xtpb_gwbse_level_ranges(&o, &r));
  return r;
}

// ---- GWBSE (gwbse.h / gwbse.cc): the driver.  Initialize fixes level ranges and options, Evaluate runs
//      Fill -> GW (G0W0 or evGW) -> Hqp -> BSE in the reference's order and returns what the reference stores into
//      Orbitals (QPpert energies, Hqp, RPA input energies, BSE energies / coefficients).  The AO three-centre slices
//      come from the caller's integral loop: `ao3c` holds auxsize symmetric n_basis x n_basis slices, full storage.
template <class Matrix = DenseMatrix, class Vector = DenseVector>
class GWBSE {
 public:
  struct options {
    int ranges = XTPB_RANGES_DEFAULT;
    double rpamax = 0, qpmin = 0, qpmax = 0, bsemin = 0, bsemax = 0;   // factors or level indices (see LevelRanges)
    Index ignore_corelevels = 0;
    bool do_singlets = true, do_triplets = false, useTDA = true;
    Index nmax = 5;
    xtpb_gw_options gw;                                                  // ranges are overwritten by Initialize
    options() { xtpb_gw_options_default(&gw); }
  };
  struct Results {
    xtpb_gwbse_ranges ranges;
    Vector QPpert_energies, RPA_input_energies, QPdiag_energies;
    Matrix Hqp, QPdiag_coefficients;
    Vector BSE_singlet_energies, BSE_triplet_energies;
    Matrix BSE_singlet_coefficients, BSE_singlet_coefficients_AR, BSE_triplet_coefficients, BSE_triplet_coefficients_AR;
  };

  explicit GWBSE(Context& ctx) : ctx_(ctx) {}
  void Initialize(const options& opt, Index n_levels, Index n_occ) {
    opt_ = opt;
    ranges_ = LevelRanges(opt.ranges, n_levels, n_occ, opt.rpamax, opt.qpmin, opt.qpmax, opt.bsemin, opt.bsemax,
                          opt.ignore_corelevels);
    opt_.gw.homo = ranges_.homo; opt_.gw.qpmin = ranges_.qpmin; opt_.gw.qpmax = ranges_.qpmax;
    opt_.gw.rpamin = ranges_.rpamin; opt_.gw.rpamax = ranges_.rpamax;
  }
  const xtpb_gwbse_ranges& ranges() const { return ranges_; }

  // vxc: qptotal x qptotal block of the exchange-correlation matrix in the MO basis; dft_energies: all levels
  Results Evaluate(const Matrix& dft_orbitals, const Vector& dft_energies, const Matrix& vxc, Index auxsize,
                   const double* ao3c, const Matrix& aux_coulomb, const Matrix* aux_overlap = nullptr) const {
    Results r;
    r.ranges = ranges_;
    TCMatrix_gwbse<Matrix> Mmn(ctx_);
    const Index mmax = ranges_.qpmax > ranges_.cmax ? ranges_.qpmax : ranges_.cmax;
    Mmn.Initialize(auxsize, ranges_.rpamin, mmax, ranges_.rpamin, ranges_.rpamax);
    Mmn.Fill3cMO_begin(dft_orbitals);
    Mmn.Fill3cMO_block(0, auxsize, ao3c, dft_orbitals.rows());
    Mmn.ApplyCoulombMetric(aux_coulomb, aux_overlap);
    GW<Matrix, Vector> gw(Mmn, vxc, dft_energies);
    gw.configure(opt_.gw);
    gw.CalculateGWPerturbation();
    r.QPpert_energies = gw.getGWAResults();
    gw.CalculateHQP();
    r.Hqp = gw.getHQP();
    r.RPA_input_energies = gw.RPAInputEnergies();
    auto diag = gw.DiagonalizeQPHamiltonian();
    r.QPdiag_energies = diag.first;
    r.QPdiag_coefficients = diag.second;
    if (opt_.do_singlets || opt_.do_triplets) {
      BSE<Matrix, Vector> bse(Mmn);
      xtpb_bse_options bo{ranges_.homo, ranges_.rpamin, ranges_.rpamax, ranges_.qpmin, ranges_.qpmax, ranges_.vmin,
                          ranges_.cmax, opt_.nmax, 1};
      bse.configure(bo, r.RPA_input_energies, r.Hqp);
      if (opt_.do_singlets) {
        if (opt_.useTDA) {
          auto s = bse.Solve_singlets();
          r.BSE_singlet_energies = s.energies; r.BSE_singlet_coefficients = s.eigenvectors;
        } else {
          auto s = bse.Solve_singlets_BTDA();
          r.BSE_singlet_energies = s.energies; r.BSE_singlet_coefficients = s.X; r.BSE_singlet_coefficients_AR = s.Y;
        }
      }
      if (opt_.do_triplets) {
        if (opt_.useTDA) {
          auto t = bse.Solve_triplets();
          r.BSE_triplet_energies = t.energies; r.BSE_triplet_coefficients = t.eigenvectors;
        } else {
          auto t = bse.Solve_triplets_BTDA();
          r.BSE_triplet_energies = t.energies; r.BSE_triplet_coefficients = t.X; r.BSE_triplet_coefficients_AR = t.Y;
        }
      }
    }
    return r;
  }

 private:
  Context& ctx_;
  options opt_{};
  xtpb_gwbse_ranges ranges_{};
};

}  // namespace xtpb200
