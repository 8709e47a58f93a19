"""The C/OpenMP part of the CPU oracle (oracle/cpu_kernels.c) against the numpy oracle, and the sampled CPU
baseline's bookkeeping."""
import numpy as np

from oracle import cpu_reference as cr
from oracle import gwbse_oracle as orc
from xtp_b200 import synth


def test_sigma_ppm_c_kernel_matches_numpy_oracle():
    prob = synth.make_problem("tiny")
    sz = prob["sizes"]
    tc = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.Fill(prob["ao3c"], prob["C"], prob["aux_coulomb"])
    rpa = orc.RPA(tc)
    rpa.configure(sz.homo, sz.rpamin, sz.rpamax)
    rpa.setRPAInputEnergies(prob["energies"][sz.rpamin:sz.rpamax + 1])
    sig = orc.Sigma_PPM(tc, rpa)
    sig.configure(orc.SigmaOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax))
    sig.PrepareScreening()
    level = 3
    omegas = np.concatenate([np.linspace(-2.0, 2.0, 41), prob["energies"][:4] + 0.3])   # includes |x| < 0.25 cases
    slab = tc.M[level + sz.qpmin - sz.rpamin]
    val, der = cr.sigma_ppm_diag(slab, sz.n_occ, rpa.getRPAInputEnergies(), sig.ppm.ppm_freq, sig._fac(), omegas, True)
    ref = np.array([sig.CalcCorrelationDiagElement(level, w) for w in omegas])
    refd = np.array([sig.CalcCorrelationDiagElementDerivative(level, w) for w in omegas])
    np.testing.assert_allclose(val, ref, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(der, refd, rtol=1e-11, atol=1e-13)


def test_unpack_symmetric():
    rng = np.random.default_rng(3)
    n = 37
    A = rng.standard_normal((n, n))
    A = A + A.T
    il = np.tril_indices(n)
    np.testing.assert_array_equal(cr.unpack_symmetric(A[il], n), A)


def test_sampled_step_reports_every_stage():
    r = cr.sampled_step(synth.WORKLOADS["tiny"], scale=0.5)
    assert set(r["stage_seconds"]) == {"fill", "metric", "epsilon", "ppm", "sigma_x", "sigma_c", "offdiag",
                                       "bse_setup", "davidson"}
    assert r["seconds"] > 0 and r["threads"] >= 1
