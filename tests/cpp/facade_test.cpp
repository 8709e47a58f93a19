// Compiles the header-only C++ facade (include/xtpb200/xtp_facade.hpp) without Eigen and, when run on a GPU box,
// drives a tiny G0W0+BSE step the way GWBSE::Evaluate does (upstream xtp/src/libxtp/gwbse/gwbse.cc), reading its
// inputs from a flat binary written by tests/test_cpp_facade.py and writing QP / BSE energies back as text.
//   usage: facade_test <input.bin> <output.txt>      (no arguments: instantiate-only self check, exit 0)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "xtpb200/xtp_facade.hpp"

using namespace xtpb200;

// the driver template is compile-checked here (its members are the calls the step below makes one by one)
template class xtpb200::GWBSE<>;

static std::vector<double> read_all(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) throw std::runtime_error("cannot open input");
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<double> v((size_t)n / 8);
  if (std::fread(v.data(), 8, v.size(), f) != v.size()) throw std::runtime_error("short read");
  std::fclose(f);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 3) {   // nothing to run without a device; the templates above were instantiated at compile time
    // GaussianQuadrature is host code: the Legendre weights mapped back to (-1,1) must sum to 2
    GaussianQuadrature gq;
    GaussianQuadrature::options qo;
    qo.order = 12;
    gq.configure(qo);
    double sum = 0.0;
    for (Index i = 0; i < gq.Order(); ++i) {      // w' = w / (1-x)^2 with omega = 0.5 (1+x)/(1-x)  =>  1-x = 1/(omega+0.5)
      const double one_minus_x = 1.0 / (gq.ScaledPoint(i) + 0.5);
      sum += gq.ScaledWeight(i) * one_minus_x * one_minus_x;
    }
    if (gq.Order() != 12 || std::abs(sum - 2.0) > 1e-12) {
      std::printf("GaussianQuadrature check failed: order %lld sum %.15f\n", (long long)gq.Order(), sum);
      return 1;
    }
    // GWBSE::Initialize level ranges are host code too: `default` on 17 levels / 5 occupied (methane, 3-21G-like)
    const xtpb_gwbse_ranges r = LevelRanges(XTPB_RANGES_DEFAULT, 17, 5);
    if (r.homo != 4 || r.rpamax != 16 || r.qpmin != 0 || r.qpmax != 9 || r.vmin != 0 || r.cmax != 9 || r.bse_size != 25) {
      std::printf("LevelRanges check failed\n");
      return 1;
    }
    std::printf("facade compiled; version %d\n", xtpb_version());
    return 0;
  }
  try {
    const std::vector<double> in = read_all(argv[1]);
    size_t p = 0;
    auto next = [&]() { return in[p++]; };
    const Index nb = (Index)next(), naux = (Index)next(), homo = (Index)next(), qpmax = (Index)next(),
                cmax = (Index)next(), nmax = (Index)next(), grid = (Index)next();
    const Index rpamax = nb - 1, mmax = qpmax > cmax ? qpmax : cmax;
    DenseMatrix C(nb, nb), V(naux, naux), vxc(qpmax + 1, qpmax + 1);
    DenseVector e(nb);
    for (Index i = 0; i < nb * nb; ++i) C.data()[i] = next();
    for (Index i = 0; i < nb; ++i) e.data()[i] = next();
    for (Index i = 0; i < (qpmax + 1) * (qpmax + 1); ++i) vxc.data()[i] = next();
    for (Index i = 0; i < naux * naux; ++i) V.data()[i] = next();
    const double* ao = in.data() + p;   // naux slices nb x nb

    Context ctx(0);
    TCMatrix_gwbse<> Mmn(ctx);
    Mmn.Initialize(naux, 0, mmax, 0, rpamax);
    Mmn.CoulombMetricBegin(V);          // the metric's eigensolver runs underneath the MO transform
    Mmn.Fill3cMO_begin(C);
    Mmn.Fill3cMO_block(0, naux, ao, nb);
    Mmn.ApplyCoulombMetric(V);

    GW<> gw(Mmn, vxc, e);
    GW<>::options opt = GW<>::default_options();
    opt.homo = homo; opt.qpmin = 0; opt.qpmax = qpmax; opt.rpamin = 0; opt.rpamax = rpamax; opt.qp_grid_steps = grid;
    gw.configure(opt);
    gw.CalculateGWPerturbation();
    const DenseVector qp = gw.getGWAResults();
    gw.CalculateHQP();
    const DenseMatrix Hqp = gw.getHQP();

    BSE<> bse(Mmn);
    BSE<>::options bo{};
    bo.homo = homo; bo.rpamin = 0; bo.rpamax = rpamax; bo.qpmin = 0; bo.qpmax = qpmax; bo.vmin = 0; bo.cmax = cmax;
    bo.nmax = nmax; bo.use_Hqp_offdiag = 1;
    bse.configure(bo, gw.RPAInputEnergies(), Hqp);
    const auto singlets = bse.Solve_singlets();
    // full BSE and oscillator strengths (unit "dipole" matrices are enough to exercise the call path)
    const auto full = bse.Solve_singlets_BTDA();
    std::array<DenseMatrix, 3> rdip{DenseMatrix(nb, nb), DenseMatrix(nb, nb), DenseMatrix(nb, nb)};
    for (int k = 0; k < 3; ++k)
      for (Index i = 0; i < nb; ++i)
        for (Index j = 0; j < nb; ++j) rdip[k](i, j) = (i == j) ? 0.0 : 1.0 / double(1 + k + (i > j ? i - j : j - i));
    const DenseMatrix dip = bse.CalcCoupledTransition_Dipoles(C, rdip, full.X, &full.Y);
    const DenseVector fosc = BSE<>::Oscillatorstrengths(full.energies, dip);

    // GW::PlotSigma table for HOMO and LUMO, BSE::Perturbative_DynamicalScreening of the TDA singlets
    const DenseMatrix plot = gw.PlotSigma("", 11, 0.05, std::vector<Index>{homo, homo + 1});
    const DenseVector dyn = bse.Perturbative_DynamicalScreening(singlets.energies, singlets.eigenvectors);

    // operator-level check through the template typedefs: Hx via BSE_OPERATOR<0,1,0,0>
    DenseVector eps_inv = bse.epsilon_0_inv();
    HxOperator<> hx(eps_inv, Mmn, Hqp);
    hx.configure(BSEOperator_Options{homo, 0, 0, 0, cmax});
    const DenseVector d = hx.diagonal();

    FILE* out = std::fopen(argv[2], "w");
    for (Index i = 0; i < qp.size(); ++i) std::fprintf(out, "qp %.15e\n", qp(i));
    for (Index i = 0; i < singlets.energies.size(); ++i) std::fprintf(out, "singlet %.15e\n", singlets.energies(i));
    std::fprintf(out, "davidson_info %d\n", singlets.info);
    for (Index i = 0; i < full.energies.size(); ++i) std::fprintf(out, "btda %.15e\n", full.energies(i));
    for (Index i = 0; i < fosc.size(); ++i) std::fprintf(out, "fosc %.15e\n", fosc(i));
    std::fprintf(out, "hx_diag0 %.15e\n", d(0));
    for (Index i = 0; i < dyn.size(); ++i) std::fprintf(out, "dynamic %.15e\n", dyn(i));
    std::fprintf(out, "plot_rows %lld\n", (long long)plot.rows());
    std::fclose(out);
    return 0;
  } catch (const std::exception& ex) {
    std::fprintf(stderr, "facade_test: %s\n", ex.what());
    return 1;
  }
}
