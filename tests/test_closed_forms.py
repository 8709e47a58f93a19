"""External pins of the CPU oracle (the reference's golden vectors are absent): closed forms
and an independent quadrature that share no code and no prefactor with oracle/gwbse_oracle.py -- see
tests/closed_forms.py for the derivations.  tests/test_gpu_closed_forms.py runs the same checks on the CUDA path."""
import math

import numpy as np
import pytest

import closed_forms as cf
from oracle import gwbse_oracle as orc
from xtp_b200 import synth


def _two_level_oracle(sysm, cls, **kw):
    tc = orc.TCMatrix_gwbse().Initialize(sysm["n_aux"], 0, 1, 0, 1)
    tc.set_raw(sysm["M"])
    rpa = orc.RPA(tc)
    rpa.configure(0, 0, 1)
    rpa.setRPAInputEnergies(sysm["energies"])
    s = cls(tc, rpa)
    s.configure(orc.SigmaOptions(0, 0, 1, 0, 1, **kw))
    s.PrepareScreening()
    return s


FREQS = [-0.9, -0.3, 0.0, 0.45, 0.9]      # in and outside the gap (-0.45, 0.20); all >= 0.25 away from e_l -+ W


@pytest.mark.parametrize("seed", [5, 6])
def test_two_level_plasmon_pole_model_is_exact(seed):
    """Pinned to the closed form: PPM weight/frequency of the single pole and Sigma_c from Sigma_PPM."""
    sysm = cf.two_level_system(seed=seed)
    s = _two_level_oracle(sysm, orc.Sigma_PPM, eta=1e-7)
    w_ref, W_ref = cf.two_level_ppm_parameters(sysm)
    live = s.ppm.ppm_weight > 1e-9
    assert live.sum() == 1
    np.testing.assert_allclose(s.ppm.ppm_weight[live], [w_ref], rtol=1e-10)
    np.testing.assert_allclose(s.ppm.ppm_freq[live], [W_ref], rtol=1e-10)
    e = sysm["energies"]
    for w in FREQS:
        assert min(abs(w - e[0] + W_ref), abs(w - e[1] - W_ref)) >= 0.25      # outside the damping window
        for level in (0, 1):
            np.testing.assert_allclose(s.CalcCorrelationDiagElement(level, w), cf.two_level_sigma_c(sysm, level, w),
                                       rtol=1e-10)


def test_two_level_exact_and_cda_match_the_closed_form():
    """Sigma_Exact (eta -> 0) to 1e-9, Sigma_CDA to the quadrature error of a 100-point rule."""
    sysm = cf.two_level_system()
    ex = _two_level_oracle(sysm, orc.Sigma_Exact, eta=1e-7)
    cda = _two_level_oracle(sysm, orc.Sigma_CDA, eta=1e-7, order=100, alpha=1e-3)
    for w in FREQS:
        for level in (0, 1):
            ref = cf.two_level_sigma_c(sysm, level, w)
            np.testing.assert_allclose(ex.CalcCorrelationDiagElement(level, w), ref, rtol=1e-9)
            assert abs(cda.CalcCorrelationDiagElement(level, w) - ref) < 2e-4 * max(1.0, abs(ref)), (level, w)


def test_cda_error_decreases_with_quadrature_order():
    """SURVEY.md section 8c item 5 as an assertion: the CDA error against the closed form falls as the order grows."""
    sysm = cf.two_level_system()
    errs = []
    for order in (8, 16, 40, 100):
        cda = _two_level_oracle(sysm, orc.Sigma_CDA, eta=1e-7, order=order, alpha=1e-3)
        errs.append(max(abs(cda.CalcCorrelationDiagElement(l, w) - cf.two_level_sigma_c(sysm, l, w))
                        for l in (0, 1) for w in FREQS))
    assert errs[-1] < 2e-4 and errs[-1] < 0.05 * errs[0], errs
    assert all(b <= 1.5 * a for a, b in zip(errs, errs[1:])), errs


def _two_level_bse(sysm):
    tc = orc.TCMatrix_gwbse().Initialize(sysm["n_aux"], 0, 1, 0, 1)
    tc.set_raw(sysm["M"])
    bse = orc.BSE(tc)
    bse.configure(orc.BSEOptions(0, 0, 1, 0, 1, 0, 1, nmax=1, davidson_tolerance="lapack"), sysm["energies"], sysm["hqp"])
    return bse


@pytest.mark.parametrize("seed", [5, 6])
def test_two_level_bse_closed_form(seed):
    """BSE::configure (eps(0), eigenbasis rotation, eps^-1) + BSE_OPERATOR + solvers against the rank-one closed form:
    TDA and full BSE, singlet and triplet."""
    sysm = cf.two_level_system(seed=seed)
    ref = cf.two_level_bse(sysm)
    bse = _two_level_bse(sysm)
    np.testing.assert_allclose(bse.make_operator("SingletOperator_TDA").get_full_matrix(), [[ref["singlet_tda"]]], rtol=1e-11)
    np.testing.assert_allclose(bse.make_operator("TripletOperator_TDA").get_full_matrix(), [[ref["triplet_tda"]]], rtol=1e-11)
    np.testing.assert_allclose(bse.Solve_singlets_TDA()[0], [ref["singlet_tda"]], rtol=1e-10)
    np.testing.assert_allclose(bse.Solve_triplets_TDA()[0], [ref["triplet_tda"]], rtol=1e-10)
    np.testing.assert_allclose(bse.Solve_singlets_BTDA()[0], [ref["singlet_full"]], rtol=1e-10)
    np.testing.assert_allclose(bse.Solve_triplets_BTDA()[0], [ref["triplet_full"]], rtol=1e-10)


def test_sigma_exact_equals_imaginary_axis_integral():
    """Sigma_Exact on the generic 'tiny' problem against the adaptive imaginary-axis quadrature (frequencies inside the
    gap): pins the sign, the spin factors of chi0 and of the residues, and the pole positions e_m -+ Omega_s."""
    p = synth.make_problem("tiny")
    sz = p["sizes"]
    tc = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.Fill(p["ao3c"], p["C"], p["aux_coulomb"])
    e = p["energies"][sz.rpamin:sz.rpamax + 1]
    rpa = orc.RPA(tc); rpa.configure(sz.homo, sz.rpamin, sz.rpamax); rpa.setRPAInputEnergies(e)
    ex = orc.Sigma_Exact(tc, rpa)
    ex.configure(orc.SigmaOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax, eta=1e-6))
    ex.PrepareScreening()
    nocc = sz.homo - sz.rpamin + 1
    gap_lo, gap_hi = e[nocc - 1], e[nocc]
    Mia = tc.M[:nocc][:, :, nocc:]
    for level in (0, sz.homo - sz.qpmin, sz.homo + 1 - sz.qpmin):
        for t in (0.25, 0.5, 0.8):
            w = gap_lo + t * (gap_hi - gap_lo)
            ref, err = cf.sigma_c_imaginary_axis(tc.M[level + sz.qpmin - sz.rpamin], Mia, e, nocc, w)
            got = ex.CalcCorrelationDiagElement(level, w)
            assert abs(got - ref) < 1e-8 + 10 * err, (level, w, got, ref, err)


def test_h2_minimal_basis_cis_matches_szabo_ostlund():
    """Literature pin of BSE_OPERATOR's exchange and direct terms: with unit screening and HF orbital energies the TDA
    BSE is configuration interaction singles.  H2 / STO-3G at R = 1.4 bohr (Szabo & Ostlund, Modern Quantum Chemistry,
    section 3.5.2 and table 3.11 ff.: e1 = -0.5782, e2 = 0.6703, J12 = 0.6636, K12 = 0.1813 Ha):
        singlet = (e2 - e1) + 2 K12 - J12,   triplet = (e2 - e1) - J12.
    The operator sees the integrals only through the RI tensor, so agreement is limited by the fit error of the
    even-tempered aux basis (~1e-4 Ha) and the four printed digits of the book (5e-4)."""
    from xtp_b200 import molecule as ml
    mol = ml.Molecule([("H", (0.0, 0.0, 0.0)), ("H", (0.0, 0.0, 1.4))])
    inp = ml.gwbse_inputs(mol)
    assert abs(inp["scf"]["energy"] - (-1.1167)) < 1e-4
    np.testing.assert_allclose(inp["energies"], [-0.5782, 0.6703], atol=1e-4)
    tc = orc.TCMatrix_gwbse().Initialize(inp["n_aux"], 0, 1, 0, 1)
    tc.Fill(inp["ao3c"], inp["C"], inp["aux_coulomb"])
    J12 = float(tc.M[0, :, 0] @ tc.M[1, :, 1])
    K12 = float(tc.M[0, :, 1] @ tc.M[0, :, 1])
    assert abs(J12 - 0.6636) < 6e-4 and abs(K12 - 0.1813) < 6e-4
    hqp = np.diag(inp["energies"])
    ones = np.ones(inp["n_aux"])
    opts = orc.BSEOperator_Options(0, 0, 0, 0, 1)
    out = {}
    for name in ("SingletOperator_TDA", "TripletOperator_TDA"):
        op = orc.BSE_OPERATOR(*orc.OPERATOR_TYPES[name], ones, tc, hqp)
        op.configure(opts)
        out[name] = float(op.get_full_matrix()[0, 0])
    d = 0.6703 + 0.5782
    assert abs(out["SingletOperator_TDA"] - (d + 2 * 0.1813 - 0.6636)) < 1.5e-3
    assert abs(out["TripletOperator_TDA"] - (d - 0.6636)) < 1e-3
    # and exactly (to rounding) against the same combination of the RI integrals
    np.testing.assert_allclose(out["SingletOperator_TDA"], inp["energies"][1] - inp["energies"][0] + 2 * K12 - J12, rtol=1e-12)


@pytest.mark.parametrize("hqp_gap", [0.55, 0.83])
def test_two_level_dynamical_screening_closed_form(hqp_gap):
    """BSE::Perturbative_DynamicalScreening on the two-level exciton against the scalar fixed-point iteration
    (hqp_gap 0.55: excitation below the KS transition; 0.83: above it, where eps(w) has a negative eigenvalue that
    SetupDirectInteractionOperator drops)."""
    sysm = cf.two_level_system(hqp_gap=hqp_gap)
    bse = _two_level_bse(sysm)
    for solve in (bse.Solve_singlets_TDA, bse.Solve_triplets_TDA):
        e, X = solve()
        dyn, its = bse.Perturbative_DynamicalScreening(sysm["energies"], e, X)
        ref, ref_its = cf.two_level_dynamical_screening(sysm, float(e[0]))
        np.testing.assert_allclose(dyn, [ref], rtol=0, atol=1e-10)
        assert its[0] == ref_its
        assert abs(dyn[0] - e[0]) > 1e-6          # the correction is not trivially zero
