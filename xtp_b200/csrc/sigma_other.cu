// Sigma_Exact and Sigma_CDA (upstream xtp/src/libxtp/gwbse/sigma_exact.cc, sigma_cda.cc,
// ImaginaryAxisIntegration.cc, gaussian_quadrature.cc).
#include "internal.h"

namespace xtpb {

void GW::prepare_exact() { throw Error("xtpb: Sigma_Exact is not implemented yet in this build"); }
void GW::prepare_cda() { throw Error("xtpb: Sigma_CDA is not implemented yet in this build"); }
void GW::sigma_c_diag_elements_other(long long, const long long*, const double*, double*, double*) {
  throw Error("xtpb: only the PPM self-energy is implemented in this build");
}
void GW::sigma_c_offdiag_other(const double*, double*) {
  throw Error("xtpb: only the PPM self-energy is implemented in this build");
}

}  // namespace xtpb
