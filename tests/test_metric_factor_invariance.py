"""CPU check (numpy oracle only) of the identity the Cholesky path of the Coulomb-metric step rests on
(xtp_b200/csrc/tc.cu: TCMatrix::apply_coulomb_metric): when no auxiliary function is removed,
every factor R with R R^T = V^-1 gives the same epsilon spectrum, the same plasmon-pole parameters, the same Sigma_x
and the same G0W0 quasiparticle energies as the reference's symmetric R = S^-1/2 (S^-1/2 V S^-1/2)^-1/2 -- and the
definiteness tests that decide "no function is removed" agree with the eigenvalue criterion of the reference."""
import copy

import numpy as np
import pytest

from oracle import gwbse_oracle as orc
from xtp_b200 import synth


@pytest.fixture(scope="module")
def prob():
    sz = synth.WORKLOADS["tiny"]
    p = synth.make_problem("tiny")
    tc = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.Fill3cMO(p["ao3c"], p["C"])
    p["tc_raw"] = tc
    return p


def _cholesky_factor(V):
    """R = U^-1 from V = U^T U (upper Cholesky factor): R R^T = V^-1, what the library keeps pending."""
    U = np.linalg.cholesky(V).T
    return np.linalg.inv(U)


def _g0w0(prob, R):
    sz = prob["sizes"]
    tc = copy.deepcopy(prob["tc_raw"])
    tc.MultiplyRightWithAuxMatrix(R)
    gw = orc.GW(tc, prob["vxc"], prob["energies"])
    gw.configure(orc.GWOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax, qp_grid_steps=201))
    gw.CalculateGWPerturbation()
    return gw, tc


@pytest.mark.parametrize("with_overlap", [False, True])
def test_any_factor_of_the_inverse_metric_gives_the_same_g0w0(prob, with_overlap):
    sz = prob["sizes"]
    V = prob["aux_coulomb"]
    S = None
    if with_overlap:
        B = np.random.default_rng(3).standard_normal((sz.n_aux, sz.n_aux))
        S = B @ B.T / sz.n_aux + 0.5 * np.eye(sz.n_aux)
    R_sym, removed = orc.Pseudo_InvSqrt_GWBSE(V, S)
    assert removed == 0
    R_chol = _cholesky_factor(V)
    np.testing.assert_allclose(R_sym @ R_sym.T, np.linalg.inv(V), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(R_chol @ R_chol.T, np.linalg.inv(V), rtol=1e-9, atol=1e-12)
    gw_s, tc_s = _g0w0(prob, R_sym)
    gw_c, tc_c = _g0w0(prob, R_chol)
    # epsilon: orthogonally similar -> same spectrum; Sigma_x and the quasiparticle energies: identical
    e = prob["energies"][sz.rpamin:sz.rpamax + 1]
    spectra = []
    for R in (R_sym, R_chol):
        t = copy.deepcopy(prob["tc_raw"])
        t.MultiplyRightWithAuxMatrix(R)
        rpa = orc.RPA(t)
        rpa.configure(sz.homo, sz.rpamin, sz.rpamax)
        rpa.setRPAInputEnergies(e)
        spectra.append((np.linalg.eigvalsh(rpa.calculate_epsilon_r(0.0)), np.linalg.eigvalsh(rpa.calculate_epsilon_i(0.5))))
    np.testing.assert_allclose(spectra[0][0], spectra[1][0], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(spectra[0][1], spectra[1][1], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(gw_c.sigma.CalcExchangeMatrix(), gw_s.sigma.CalcExchangeMatrix(), rtol=0, atol=1e-11)
    np.testing.assert_allclose(np.sort(gw_c.sigma.ppm.ppm_weight), np.sort(gw_s.sigma.ppm.ppm_weight), rtol=0, atol=1e-9)
    np.testing.assert_allclose(gw_c.getGWAResults(), gw_s.getGWAResults(), rtol=0, atol=1e-9)
    # after the plasmon-pole rotation the tensors agree up to the sign of each eigenvector (generic spectrum)
    for m in (0, sz.homo - sz.rpamin, sz.mmax - sz.rpamin):
        np.testing.assert_allclose(np.abs(tc_c[m]), np.abs(tc_s[m]), rtol=0, atol=1e-8)


def test_definiteness_tests_equal_the_eigenvalue_criterion():
    """S - etol and V - etol S positive definite  <=>  no eigenvalue of S or of S^-1/2 V S^-1/2 below etol."""
    rng = np.random.default_rng(11)
    n, etol = 40, 5e-7

    def pd(A):
        try:
            np.linalg.cholesky(A)
            return True
        except np.linalg.LinAlgError:
            return False

    for trial in range(12):
        B = rng.standard_normal((n, n))
        S = B @ B.T / n + 0.3 * np.eye(n)
        w, U = np.linalg.eigh(rng.standard_normal((n, n)))
        lam = rng.uniform(0.5, 3.0, n)
        if trial % 3 == 1:
            lam[:2] = 1e-8                      # below etol in the S-orthogonalised metric
        if trial % 3 == 2:
            lam[0] = 5e-6                       # small but kept
        Sh = np.linalg.cholesky(S)
        V = Sh @ ((U * lam) @ U.T) @ Sh.T       # S^-1/2 V S^-1/2 is similar to U diag(lam) U^T
        _, removed = orc.Pseudo_InvSqrt_GWBSE(V, S, etol)
        none_removed = pd(S - etol * np.eye(n)) and pd(V - etol * S)
        assert none_removed == (removed == 0), (trial, removed)
        _, removed0 = orc.Pseudo_InvSqrt_GWBSE(V, None, etol)
        assert pd(V - etol * np.eye(n)) == (removed0 == 0), (trial, removed0)
