"""-m gpu: the CUDA path against closed forms and an independent quadrature (tests/closed_forms.py) -- results that do
not go through oracle/gwbse_oracle.py at all, so a mistake shared by the oracle and the kernels (both were written
from the same recollection of upstream) cannot hide here."""
import numpy as np
import pytest

import closed_forms as cf
from xtp_b200 import synth

pytestmark = pytest.mark.gpu

FREQS = [-0.9, -0.3, 0.0, 0.45, 0.9]


@pytest.fixture(scope="module")
def ctx():
    from xtp_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def _gw(ctx, sysm, **kw):
    from xtp_b200 import api
    tc = api.TCMatrix_gwbse(ctx).Initialize(sysm["n_aux"], 0, 1, 0, 1)
    tc.set_raw(sysm["M"])
    gw = api.GW(ctx, tc, np.zeros((2, 2)), sysm["energies"])
    gw.configure(api.gw_options(homo=0, qpmin=0, qpmax=1, rpamin=0, rpamax=1, **kw))
    gw.PrepareScreening()
    return gw, tc


@pytest.mark.parametrize("seed", [5, 6])
def test_two_level_sigma_ppm(ctx, seed):
    sysm = cf.two_level_system(seed=seed)
    gw, _ = _gw(ctx, sysm, eta=1e-7)
    weight, freq = gw.getPpm()
    w_ref, W_ref = cf.two_level_ppm_parameters(sysm)
    live = weight > 1e-9
    assert live.sum() == 1
    np.testing.assert_allclose(weight[live], [w_ref], rtol=1e-9)
    np.testing.assert_allclose(freq[live], [W_ref], rtol=1e-9)
    levels = np.repeat([0, 1], len(FREQS))
    freqs = np.tile(FREQS, 2)
    val = gw.CalcCorrelationDiagElements(levels, freqs)
    ref = [cf.two_level_sigma_c(sysm, int(l), float(w)) for l, w in zip(levels, freqs)]
    np.testing.assert_allclose(val, ref, rtol=1e-9)
    # the QP-grid kernels (compressed and pole-by-pole scans) on the same closed form, away from the damped window
    grid = gw.CalcCorrelationGrid(np.array([0.0, 0.0]))
    steps, spacing = 1001, 0.01
    om = (np.arange(steps) - (steps - 1) / 2) * spacing
    e = sysm["energies"]
    far = (np.abs(om - e[0] + W_ref) >= 0.25) & (np.abs(om - e[1] - W_ref) >= 0.25)
    for level in (0, 1):
        ref = np.array([cf.two_level_sigma_c(sysm, level, float(w)) for w in om[far]])
        np.testing.assert_allclose(grid[level][far], ref, rtol=1e-8)


def test_two_level_sigma_exact_and_cda(ctx):
    sysm = cf.two_level_system()
    levels = np.repeat([0, 1], len(FREQS))
    freqs = np.tile(FREQS, 2)
    ref = np.array([cf.two_level_sigma_c(sysm, int(l), float(w)) for l, w in zip(levels, freqs)])
    gw, _ = _gw(ctx, sysm, eta=1e-7, sigma_integration="exact")
    np.testing.assert_allclose(gw.CalcCorrelationDiagElements(levels, freqs), ref, rtol=1e-8)
    gw, _ = _gw(ctx, sysm, eta=1e-7, sigma_integration="cda", order=100, alpha=1e-3)
    val = gw.CalcCorrelationDiagElements(levels, freqs)
    assert np.all(np.abs(val - ref) < 2e-4 * np.maximum(1.0, np.abs(ref)))


@pytest.mark.parametrize("seed", [5, 6])
@pytest.mark.parametrize("mode", ["dense", "factorised"])
def test_two_level_bse(ctx, seed, mode, monkeypatch):
    from xtp_b200 import api
    monkeypatch.setenv("XTPB_BSE_DENSE_MAX_GB", "0" if mode == "factorised" else "32")
    sysm = cf.two_level_system(seed=seed)
    ref = cf.two_level_bse(sysm)
    tc = api.TCMatrix_gwbse(ctx).Initialize(sysm["n_aux"], 0, 1, 0, 1)
    tc.set_raw(sysm["M"])
    bse = api.BSE(ctx, tc)
    bse.configure(0, 0, 1, 0, 1, 0, 1, 1, sysm["energies"], sysm["hqp"], davidson_tolerance="lapack")
    np.testing.assert_allclose(bse.Solve_singlets_TDA()[0], [ref["singlet_tda"]], rtol=1e-10)
    np.testing.assert_allclose(bse.Solve_triplets_TDA()[0], [ref["triplet_tda"]], rtol=1e-10)
    np.testing.assert_allclose(bse.Solve_singlets_BTDA()[0], [ref["singlet_full"]], rtol=1e-9)
    np.testing.assert_allclose(bse.Solve_triplets_BTDA()[0], [ref["triplet_full"]], rtol=1e-9)


def test_sigma_exact_equals_imaginary_axis_integral(ctx):
    """Sigma_Exact on the device against the adaptive imaginary-axis quadrature of the host (generic 'tiny' problem,
    frequencies inside the gap)."""
    from xtp_b200 import api
    p = synth.make_problem("tiny")
    sz = p["sizes"]
    tc = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.Fill(p["ao3c"], p["C"], p["aux_coulomb"])
    M = tc.get_raw()
    gw = api.GW(ctx, tc, p["vxc"], p["energies"])
    gw.configure(api.gw_options(homo=sz.homo, qpmin=sz.qpmin, qpmax=sz.qpmax, rpamin=sz.rpamin, rpamax=sz.rpamax,
                                eta=1e-6, sigma_integration="exact"))
    gw.PrepareScreening()
    e = p["energies"][sz.rpamin:sz.rpamax + 1]
    nocc = sz.homo - sz.rpamin + 1
    lv, fr, ref = [], [], []
    for level in (0, sz.homo - sz.qpmin, sz.homo + 1 - sz.qpmin):
        for t in (0.25, 0.5, 0.8):
            w = e[nocc - 1] + t * (e[nocc] - e[nocc - 1])
            r, err = cf.sigma_c_imaginary_axis(M[level + sz.qpmin - sz.rpamin], M[:nocc][:, :, nocc:], e, nocc, w)
            lv.append(level); fr.append(w); ref.append(r)
    np.testing.assert_allclose(gw.CalcCorrelationDiagElements(np.array(lv), np.array(fr)), ref, rtol=0, atol=1e-8)


def test_h2_cis_matches_szabo_ostlund(ctx):
    """BSE_OPERATOR on the device with unit screening and HF energies = CIS of H2 / STO-3G (Szabo & Ostlund)."""
    from xtp_b200 import api, molecule as ml
    mol = ml.Molecule([("H", (0.0, 0.0, 0.0)), ("H", (0.0, 0.0, 1.4))])
    inp = ml.gwbse_inputs(mol)
    tc = api.TCMatrix_gwbse(ctx).Initialize(inp["n_aux"], 0, 1, 0, 1)
    tc.Fill(inp["ao3c"], inp["C"], inp["aux_coulomb"])
    hqp = np.diag(inp["energies"])
    ones = np.ones(inp["n_aux"])
    d = 0.6703 + 0.5782
    for name, ref, tol in (("SingletOperator_TDA", d + 2 * 0.1813 - 0.6636, 1.5e-3), ("TripletOperator_TDA", d - 0.6636, 1e-3)):
        cqp, cx, cdd, cd2 = api.OPERATOR_TYPES[name]
        op = api.BSE_OPERATOR(ctx, cqp, cx, cdd, cd2, ones, tc, hqp, 0, 0, 0, 1)
        assert abs(float(op.get_full_matrix()[0, 0]) - ref) < tol
        op.close()


@pytest.mark.parametrize("hqp_gap", [0.55, 0.83])
def test_two_level_dynamical_screening(ctx, hqp_gap):
    from xtp_b200 import api
    sysm = cf.two_level_system(hqp_gap=hqp_gap)
    tc = api.TCMatrix_gwbse(ctx).Initialize(sysm["n_aux"], 0, 1, 0, 1)
    tc.set_raw(sysm["M"])
    bse = api.BSE(ctx, tc)
    bse.configure(0, 0, 1, 0, 1, 0, 1, 1, sysm["energies"], sysm["hqp"], davidson_tolerance="lapack")
    for solve in (bse.Solve_singlets_TDA, bse.Solve_triplets_TDA):
        e, X = solve()
        dyn = bse.Perturbative_DynamicalScreening(e, X)
        ref, ref_its = cf.two_level_dynamical_screening(sysm, float(e[0]))
        np.testing.assert_allclose(dyn, [ref], rtol=0, atol=1e-9)
        assert bse.dynamical_iterations[0] == ref_its
