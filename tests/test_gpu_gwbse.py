"""-m gpu parity tests: every stage of the GW-BSE path through the C ABI (ctypes) against the CPU oracle on
identical seeded inputs.  Tolerances are those of BASELINE.json's north_star: M_mn^P 1e-10 relative,
energies 1e-6 Hartree (tests use tighter bounds where FP64 allows)."""
import copy

import numpy as np
import pytest

from oracle import gwbse_oracle as orc
from xtp_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from xtp_b200 import api
    c = api.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module", params=["tiny", "ch4-svp-shape", "odd"])
def prob(request):
    if request.param == "odd":      # odd level counts everywhere: exercises the 8-byte (unaligned) load paths
        sz = synth.Sizes(n_basis=45, n_aux=91, homo=6, qpmax=15, cmax=15)
        p = synth.make_problem(sz, seed=77)
    else:
        p = synth.make_problem(request.param)
    sz = p["sizes"]
    tc = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.Fill(p["ao3c"], p["C"], p["aux_coulomb"])
    p["tc_o"] = tc
    return p


def gpu_tc(ctx, prob, raw=None):
    from xtp_b200 import api
    sz = prob["sizes"]
    tc = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.set_raw(prob["tc_o"].M if raw is None else raw)
    return tc


def rel(a, b):
    return np.abs(a - b).max() / max(1e-300, np.abs(b).max())


def test_set_raw_roundtrip(ctx, prob):
    tc = gpu_tc(ctx, prob)
    np.testing.assert_array_equal(tc.get_raw(), prob["tc_o"].M)
    np.testing.assert_array_equal(tc[1], prob["tc_o"][1])


def test_fill_matches_oracle(ctx, prob):
    """TCMatrix_gwbse::Fill: K1 (MO transform) + Coulomb metric + K2 (aux rotation)."""
    from xtp_b200 import api
    sz = prob["sizes"]
    tc = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.Fill3cMO(prob["ao3c"], prob["C"], block=7)
    ref = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    ref.Fill3cMO(prob["ao3c"], prob["C"])
    assert rel(tc.get_raw(), ref.M) < 1e-12
    removed = tc.apply_coulomb_metric(prob["aux_coulomb"])
    assert removed == 0
    assert rel(tc.get_raw(), prob["tc_o"].M) < 1e-10


def test_fill_packed_matches_full(ctx, prob):
    """packed lower-triangular AO slices (pinned host buffer, double-buffered H2D) give the same M as full slices."""
    from xtp_b200 import api
    sz = prob["sizes"]
    full = gpu_tc(ctx, prob)
    full.Fill3cMO(prob["ao3c"], prob["C"])
    ref = full.get_raw()
    packed = api.pack_lower(prob["ao3c"])
    pin = api.PinnedBuffer(packed.size)
    pin.array[:] = packed.reshape(-1)
    tc = gpu_tc(ctx, prob)
    tc.fill_begin(prob["C"])
    half = sz.n_aux // 2
    view = pin.array.reshape(packed.shape)
    tc.fill_block_packed(0, view[:half])
    tc.fill_block_packed(half, view[half:])
    assert rel(tc.get_raw(), ref) < 1e-12      # split-K choice depends on the batch size: not bit-identical
    pin.close()


def test_coulomb_metric_with_overlap_and_removed_functions(ctx, prob):
    from xtp_b200 import api
    sz = prob["sizes"]
    rng = np.random.default_rng(5)
    B = rng.standard_normal((sz.n_aux, sz.n_aux))
    S = B @ B.T / sz.n_aux + 0.5 * np.eye(sz.n_aux)
    # rank-deficient Coulomb matrix: three functions must be removed
    w, U = np.linalg.eigh(prob["aux_coulomb"])
    w[:3] = 1e-9
    V = (U * w) @ U.T
    tc = gpu_tc(ctx, prob)
    removed = tc.apply_coulomb_metric(V, S)
    ref = copy.deepcopy(prob["tc_o"])
    R, removed_ref = orc.Pseudo_InvSqrt_GWBSE(V, S)
    ref.MultiplyRightWithAuxMatrix(R)
    assert removed == removed_ref
    assert rel(tc.get_raw(), ref.M) < 1e-8


def test_coulomb_metric_cholesky_path(ctx, prob, monkeypatch):
    """When no function would be removed the metric factor comes from a Cholesky factorisation (R = U^-1, R R^T = V^-1)
    and stays pending: epsilon and the G0W0 energies do not depend on which factor is used, and reading the tensor
    builds the reference's symmetric factor after all.  A matrix with eigenvalues below etol takes the eigensolver."""
    from xtp_b200 import api
    sz = prob["sizes"]
    rng = np.random.default_rng(15)
    B = rng.standard_normal((sz.n_aux, sz.n_aux))
    S = B @ B.T / sz.n_aux + 0.5 * np.eye(sz.n_aux)
    for overlap in (None, S):
        tc = gpu_tc(ctx, prob)
        assert tc.apply_coulomb_metric(prob["aux_coulomb"], overlap) == 0
        info = tc.metric_path_info()
        assert info == {"cholesky_calls": 1, "eigensolver_calls": 0}
        # epsilon is handed to the caller in the reference's aux basis: asking for it builds the symmetric factor
        e = prob["energies"][sz.rpamin:sz.rpamax + 1]
        rpa_c = api.RPA(tc)
        rpa_c.configure(sz.homo, sz.rpamin, sz.rpamax)
        rpa_c.setRPAInputEnergies(e)
        eps_c = rpa_c.calculate_epsilon_i(0.5)
        assert tc.metric_path_info() == {"cholesky_calls": 1, "eigensolver_calls": 1}
        ref = copy.deepcopy(prob["tc_o"])
        R, _ = orc.Pseudo_InvSqrt_GWBSE(prob["aux_coulomb"], overlap)
        ref.MultiplyRightWithAuxMatrix(R)
        rpa_o = orc.RPA(ref)
        rpa_o.configure(sz.homo, sz.rpamin, sz.rpamax)
        rpa_o.setRPAInputEnergies(e)
        assert rel(eps_c, rpa_o.calculate_epsilon_i(0.5)) < 1e-9
        assert rel(tc.get_raw(), ref.M) < 1e-9
        assert tc.metric_path_info() == {"cholesky_calls": 1, "eigensolver_calls": 1}
    # a caller's own rotation right after the metric step meets the reference's symmetric factor, not the Cholesky one
    tc = gpu_tc(ctx, prob)
    tc.apply_coulomb_metric(prob["aux_coulomb"])
    Ruser = np.random.default_rng(16).standard_normal((sz.n_aux, sz.n_aux))
    tc.MultiplyRightWithAuxMatrix(Ruser)
    assert tc.metric_path_info() == {"cholesky_calls": 1, "eigensolver_calls": 1}
    ref = copy.deepcopy(prob["tc_o"])
    R, _ = orc.Pseudo_InvSqrt_GWBSE(prob["aux_coulomb"], None)
    ref.MultiplyRightWithAuxMatrix(R)
    ref.MultiplyRightWithAuxMatrix(Ruser)
    assert rel(tc.get_raw(), ref.M) < 1e-9
    # G0W0 + BSE through both factors
    out = {}
    for mode in ("1", "0"):
        tc = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
        tc.Fill3cMO(prob["ao3c"], prob["C"])
        if mode == "1":
            tc.apply_coulomb_metric(prob["aux_coulomb"])
            assert tc.metric_path_info()["cholesky_calls"] == 1
        else:
            tc.apply_coulomb_metric(prob["aux_coulomb"])
            tc.get_raw()                                   # forces the symmetric factor onto the tensor
            assert tc.metric_path_info()["eigensolver_calls"] == 1
        gw = api.GW(ctx, tc, prob["vxc"], prob["energies"])
        gw.configure(api.gw_options(homo=sz.homo, qpmin=sz.qpmin, qpmax=sz.qpmax, rpamin=sz.rpamin, rpamax=sz.rpamax))
        gw.CalculateGWPerturbation()
        out[mode] = gw.getGWAResults()
        if mode == "1":
            assert tc.metric_path_info()["eigensolver_calls"] == 0      # folded into the PPM rotation, never built
    np.testing.assert_allclose(out["1"], out["0"], rtol=0, atol=1e-8)
    # eigenvalues below etol: the Cholesky test must fail and the eigensolver path must remove them
    w, U = np.linalg.eigh(prob["aux_coulomb"])
    w[:2] = 1e-9
    tc = gpu_tc(ctx, prob)
    assert tc.apply_coulomb_metric((U * w) @ U.T) == 2
    assert tc.metric_path_info() == {"cholesky_calls": 0, "eigensolver_calls": 1}


def test_ppm_epsilon_prefetch_under_fill(ctx):
    """xtpb_tc_ppm_prefetch_begin: the plasmon-pole model's two epsilon matrices are accumulated panel by panel while the
    aux blocks arrive (several 256-row panels here) and PrepareScreening picks them up; G0W0 energies equal those of the
    plain sequence.  Other energies, a rotation in between, or an out-of-order fill: the screening is recomputed."""
    from xtp_b200 import api
    sz = synth.Sizes(n_basis=96, n_aux=600, homo=11)
    p = synth.make_problem(sz, seed=77)
    e_rpa = p["energies"][sz.rpamin:sz.rpamax + 1]

    def step(hint, order=None, energies=None, rotate=False):
        tc = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
        tc.fill_begin(p["C"])
        if hint:
            tc.ppm_prefetch_begin(e_rpa if energies is None else energies, sz.homo)
        blocks = list(range(0, sz.n_aux, 100))
        for p0 in (blocks if order is None else order):
            tc.fill_block(p0, p["ao3c"][p0:p0 + 100])
        tc.apply_coulomb_metric(p["aux_coulomb"])
        if rotate:
            tc.MultiplyRightWithAuxMatrix(np.eye(sz.n_aux))
        info = tc.ppm_prefetch_info()
        gw = api.GW(ctx, tc, p["vxc"], p["energies"])
        gw.configure(api.gw_options(homo=sz.homo, qpmin=sz.qpmin, qpmax=sz.qpmax, rpamin=sz.rpamin, rpamax=sz.rpamax,
                                    qp_grid_steps=201))
        gw.CalculateGWPerturbation()
        return gw.getGWAResults(), info, tc.ppm_prefetch_info()

    plain, info0, _ = step(False)
    assert not info0["complete"]
    qp, info, after = step(True)
    assert info["complete"] and info["aux_functions_done"] == sz.n_aux and after["matrices_used"] == 2
    np.testing.assert_allclose(qp, plain, rtol=0, atol=1e-9)
    # energies that are not the ones the GW object uses: nothing is picked up, same result
    qp2, info2, after2 = step(True, energies=e_rpa + 1e-3)
    assert info2["complete"] and after2["matrices_used"] == 0
    np.testing.assert_allclose(qp2, plain, rtol=0, atol=1e-9)
    # the tensor changes after the fill: stale
    qp3, _, after3 = step(True, rotate=True)
    assert after3["matrices_used"] == 0
    np.testing.assert_allclose(qp3, plain, rtol=0, atol=1e-8)
    # out-of-order blocks disarm the prefetch
    qp4, info4, after4 = step(True, order=[100, 0, 200, 300, 400, 500])
    assert not info4["complete"] and after4["matrices_used"] == 0
    np.testing.assert_allclose(qp4, plain, rtol=0, atol=1e-9)


def test_multiply_right_with_aux_matrix(ctx, prob):
    sz = prob["sizes"]
    R = np.random.default_rng(2).standard_normal((sz.n_aux, sz.n_aux))
    tc = gpu_tc(ctx, prob)
    tc.MultiplyRightWithAuxMatrix(R)
    ref = copy.deepcopy(prob["tc_o"])
    ref.MultiplyRightWithAuxMatrix(R)
    assert rel(tc.get_raw(), ref.M) < 1e-12


def _rpa_pair(ctx, prob):
    from xtp_b200 import api
    sz = prob["sizes"]
    e = prob["energies"][sz.rpamin:sz.rpamax + 1]
    tc = gpu_tc(ctx, prob)
    g = api.RPA(tc); g.configure(sz.homo, sz.rpamin, sz.rpamax); g.setRPAInputEnergies(e)
    o = orc.RPA(prob["tc_o"]); o.configure(sz.homo, sz.rpamin, sz.rpamax); o.setRPAInputEnergies(e)
    return g, o


def test_epsilon_imag_and_real(ctx, prob):
    g, o = _rpa_pair(ctx, prob)
    for w in [0.0, 0.5, 3.0]:
        assert rel(g.calculate_epsilon_i(w), o.calculate_epsilon_i(w)) < 1e-12
    for w in [0.0, 0.2]:
        assert rel(g.calculate_epsilon_r(w), o.calculate_epsilon_r(w)) < 1e-12
    batch = g.calculate_epsilon_batch([0.1, 0.7, 2.0], imag=True)
    for i, w in enumerate([0.1, 0.7, 2.0]):
        assert rel(batch[i], o.calculate_epsilon_i(w)) < 1e-12


def _gw_pair(ctx, prob, **kw):
    from xtp_b200 import api
    sz = prob["sizes"]
    tc = gpu_tc(ctx, prob)
    gw = api.GW(ctx, tc, prob["vxc"], prob["energies"])
    gw.configure(api.gw_options(homo=sz.homo, qpmin=sz.qpmin, qpmax=sz.qpmax, rpamin=sz.rpamin, rpamax=sz.rpamax, **kw))
    tco = copy.deepcopy(prob["tc_o"])
    okw = dict(kw)
    if "sigma_integration" in okw:
        okw["sigma_integration"] = okw["sigma_integration"]
    gwo = orc.GW(tco, prob["vxc"], prob["energies"])
    gwo.configure(orc.GWOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax, **okw))
    return gw, gwo, tc, tco


def test_sigma_exchange(ctx, prob):
    gw, gwo, _, _ = _gw_pair(ctx, prob)
    assert rel(gw.CalcExchangeMatrix(), gwo.sigma.CalcExchangeMatrix()) < 1e-12


def test_ppm_parameters_and_sigma_c(ctx, prob):
    sz = prob["sizes"]
    gw, gwo, tc, tco = _gw_pair(ctx, prob)
    gwo.rpa.setRPAInputEnergies(prob["energies"][sz.rpamin:sz.rpamax + 1])
    gw.PrepareScreening()
    gwo.sigma.PrepareScreening()
    w, f = gw.getPpm()
    np.testing.assert_allclose(w, gwo.sigma.ppm.ppm_weight, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(f, gwo.sigma.ppm.ppm_freq, rtol=1e-8, atol=1e-12)
    q = sz.qptotal
    rng = np.random.default_rng(4)
    levels = rng.integers(0, q, 40)
    freqs = rng.uniform(-1.5, 1.5, 40)
    val, der = gw.CalcCorrelationDiagElements(levels, freqs, derivative=True)
    ref = np.array([gwo.sigma.CalcCorrelationDiagElement(int(l), float(x)) for l, x in zip(levels, freqs)])
    refd = np.array([gwo.sigma.CalcCorrelationDiagElementDerivative(int(l), float(x)) for l, x in zip(levels, freqs)])
    np.testing.assert_allclose(val, ref, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(der, refd, rtol=1e-9, atol=1e-10)
    fr = prob["energies"][sz.qpmin:sz.qpmax + 1]
    np.testing.assert_allclose(gw.CalcCorrelationDiag(fr), gwo.sigma.CalcCorrelationDiag(fr), rtol=1e-9, atol=1e-11)
    off = gw.CalcCorrelationOffDiag(fr)
    np.testing.assert_allclose(off, gwo.sigma.CalcCorrelationOffDiag(fr), rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("mode", ["compressed", "direct"])
@pytest.mark.parametrize("steps,spacing", [(173, 0.013), (1001, 0.01)])
def test_sigma_c_qp_grid(ctx, prob, mode, steps, spacing, monkeypatch):
    """The grid scan of GW::SolveQP_Grid against element-wise oracle evaluations, both ways it can run: compressed
    (far poles through Chebyshev moments, near poles one by one) and the plain pole-by-pole kernel (damped-window
    branch).  Every level; a 173-point grid (odd, ragged against the 32-point chunks) that crosses many poles and the
    reference's default 1001 x 0.01 Ha grid."""
    sz = prob["sizes"]
    monkeypatch.setenv("XTPB_SIGMA_GRID", mode)
    gw, gwo, _, _ = _gw_pair(ctx, prob, qp_grid_steps=steps, qp_grid_spacing=spacing)
    gwo.rpa.setRPAInputEnergies(prob["energies"][sz.rpamin:sz.rpamax + 1])
    gw.PrepareScreening()
    gwo.sigma.PrepareScreening()
    centers = prob["energies"][sz.qpmin:sz.qpmax + 1].copy()
    grid = gw.CalcCorrelationGrid(centers)
    assert grid.shape == (sz.qptotal, steps)
    info = gw.grid_scan_info()
    assert info["compressed"] == (mode == "compressed")
    assert 0 < info["direct_evaluations"] and info["equivalent_evaluations"] == sz.ntotal * sz.n_aux * steps * sz.qptotal
    rng = np.random.default_rng(9)
    for level in rng.choice(sz.qptotal, size=min(6, sz.qptotal), replace=False):
        js = np.unique(np.concatenate([[0, steps - 1, steps // 2], rng.integers(0, steps, 20)]))
        ref = np.array([gwo.sigma.CalcCorrelationDiagElement(int(level), centers[level] + (j - (steps - 1) / 2) * spacing)
                        for j in js])
        np.testing.assert_allclose(grid[level, js], ref, rtol=1e-9, atol=1e-11)
    # the pair kernel and the grid kernel must agree where both are evaluated
    lv = np.arange(sz.qptotal)
    np.testing.assert_allclose(grid[:, (steps - 1) // 2], gw.CalcCorrelationDiagElements(lv, centers), rtol=1e-10,
                               atol=1e-12)


def test_sigma_c_points_through_scan_state(ctx, prob, monkeypatch):
    """Single (level, frequency) values after a compressed grid scan come from its moments (no slab traffic) and agree
    with the slab-streaming pair kernel and the oracle -- inside the grid, between grid points, far outside the binned
    core; anything that changes the tensor, the energies or the PPM parameters sends them back to the pair kernel."""
    sz = prob["sizes"]
    monkeypatch.setenv("XTPB_SIGMA_GRID", "compressed")
    gw, gwo, tc, _ = _gw_pair(ctx, prob)
    gwo.rpa.setRPAInputEnergies(prob["energies"][sz.rpamin:sz.rpamax + 1])
    gw.PrepareScreening()
    gwo.sigma.PrepareScreening()
    centers = prob["energies"][sz.qpmin:sz.qpmax + 1].copy()
    rng = np.random.default_rng(21)
    lv = rng.integers(0, sz.qptotal, 40)
    fr = centers[lv] + rng.uniform(-5.0, 5.0, 40) * rng.choice([1.0, 0.1, 3.0], 40)
    before = gw.point_eval_info()
    v_pairs = gw.CalcCorrelationDiagElements(lv, fr)                    # no scan yet: pair kernel
    assert gw.point_eval_info()["direct_calls"] == before["direct_calls"] + 1
    gw.CalcCorrelationGrid(centers)
    assert gw.grid_scan_info()["compressed"]
    mid = gw.point_eval_info()
    v_scan = gw.CalcCorrelationDiagElements(lv, fr)
    after = gw.point_eval_info()
    assert after["compressed_calls"] == mid["compressed_calls"] + 1 and after["direct_calls"] == mid["direct_calls"]
    scale = max(np.abs(v_pairs).max(), 1e-3)
    assert np.abs(v_scan - v_pairs).max() < 1e-10 * scale
    ref = np.array([gwo.sigma.CalcCorrelationDiagElement(int(l), float(f)) for l, f in zip(lv[:12], fr[:12])])
    np.testing.assert_allclose(v_scan[:12], ref, rtol=1e-9, atol=1e-11)
    # derivatives always take the pair kernel
    gw.CalcCorrelationDiagElements(lv, fr, True)
    assert gw.point_eval_info()["compressed_calls"] == after["compressed_calls"]
    # XTPB_SIGMA_POINTS=direct
    monkeypatch.setenv("XTPB_SIGMA_POINTS", "direct")
    np.testing.assert_array_equal(gw.CalcCorrelationDiagElements(lv, fr), v_pairs)
    assert gw.point_eval_info()["compressed_calls"] == after["compressed_calls"]
    monkeypatch.delenv("XTPB_SIGMA_POINTS")
    # new energies invalidate the state
    e2 = prob["energies"][sz.rpamin:sz.rpamax + 1].copy()
    e2[sz.n_occ:] += 0.01
    gw.setRPAInputEnergies(e2)
    n0 = gw.point_eval_info()["compressed_calls"]
    gw.CalcCorrelationDiagElements(lv, fr)
    assert gw.point_eval_info()["compressed_calls"] == n0
    # ... and so does a rotation of the tensor behind the GW object's back
    gw.CalcCorrelationGrid(centers)
    tc.MultiplyRightWithAuxMatrix(np.eye(sz.n_aux))
    gw.CalcCorrelationDiagElements(lv, fr)
    assert gw.point_eval_info()["compressed_calls"] == n0


@pytest.mark.parametrize("solver", ["grid", "fixedpoint"])
def test_g0w0_qp_energies_same_through_scan_state_and_pair_kernel(ctx, prob, solver, monkeypatch):
    """G0W0 quasiparticle energies with the bisection rounds served three levels at a time from the grid scan's moments
    (default) and one midpoint at a time by the pair kernel (XTPB_SIGMA_POINTS=direct): same roots."""
    out = {}
    for mode in ("scan", "direct"):
        if mode == "direct":
            monkeypatch.setenv("XTPB_SIGMA_POINTS", "direct")
        gw, _, _, _ = _gw_pair(ctx, prob, qp_solver=solver)
        gw.CalculateGWPerturbation()
        out[mode] = (gw.getGWAResults(), gw.point_eval_info())
    if solver == "grid":
        assert out["scan"][1]["compressed_calls"] > 0
    assert out["direct"][1]["compressed_calls"] == 0
    np.testing.assert_allclose(out["scan"][0], out["direct"][0], rtol=0, atol=1e-8)


def test_sigma_c_qp_grid_compressed_vs_direct_large(ctx, monkeypatch):
    """synth-500 shape (500 levels x 1500 aux functions, 100 QP levels, the default 1001-point grid): the compressed
    scan against the pole-by-pole kernel on every grid point, against a numpy pole sum over the device's own rotated
    slab on sampled points, with unsorted RPA energies (must fall back to pole by pole), and the saving it buys."""
    from xtp_b200 import api
    sz = synth.WORKLOADS["synth-500"]
    rng = np.random.default_rng(3)
    tc = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.set_raw(synth.make_M_direct(sz, rng))
    e = synth.make_energies(sz, rng)
    gw = api.GW(ctx, tc, synth.make_vxc(sz, rng), e)
    gw.configure(api.gw_options(homo=sz.homo, qpmin=sz.qpmin, qpmax=sz.qpmax, rpamin=sz.rpamin, rpamax=sz.rpamax))
    gw.PrepareScreening()
    centers = e[sz.qpmin:sz.qpmax + 1].copy()
    monkeypatch.setenv("XTPB_SIGMA_GRID", "direct")
    direct = gw.CalcCorrelationGrid(centers)
    assert not gw.grid_scan_info()["compressed"]
    monkeypatch.setenv("XTPB_SIGMA_GRID", "compressed")
    comp = gw.CalcCorrelationGrid(centers)
    info = gw.grid_scan_info()
    assert info["compressed"] and info["bins"] > 8
    assert info["direct_evaluations"] < 0.25 * info["equivalent_evaluations"]
    scale = np.abs(direct).max()
    assert np.abs(comp - direct).max() < 1e-11 * scale
    # numpy pole sum over the rotated slab the device holds
    weight, freq = gw.getPpm()
    fac = np.where(weight < 1e-9, 0.0, 0.5 * weight * freq)
    steps, spacing = 1001, 0.01
    for level in (0, sz.homo, sz.homo + 1, sz.qptotal - 1):
        slab = tc[level + sz.qpmin - sz.rpamin].T            # [P, m]
        z = np.where(np.arange(sz.ntotal)[None, :] < sz.n_occ, e[None, :] - freq[:, None], e[None, :] + freq[:, None])
        a = fac[:, None] * slab * slab
        for j in (0, 137, 500, 731, 1000):
            om = centers[level] + (j - (steps - 1) / 2) * spacing
            ref = (a * orc.ppm_stabilized_inverse(om - z)).sum()
            assert abs(comp[level, j] - ref) < 1e-10 * scale
    # unsorted energies: bins are no longer contiguous m-ranges -> the scan must decline the compressed path
    e2 = e[sz.rpamin:sz.rpamax + 1].copy()
    e2[[sz.n_occ + 3, sz.n_occ + 4]] = e2[[sz.n_occ + 4, sz.n_occ + 3]]
    gw.setRPAInputEnergies(e2)
    unsorted_grid = gw.CalcCorrelationGrid(centers)
    assert not gw.grid_scan_info()["compressed"]
    monkeypatch.setenv("XTPB_SIGMA_GRID", "direct")
    np.testing.assert_array_equal(unsorted_grid, gw.CalcCorrelationGrid(centers))


def test_sigma_exact(ctx, prob):
    """Sigma_Exact: RPA::Diagonalize_H2p + residues + diagonal / derivative / off-diagonal elements."""
    sz = prob["sizes"]
    gw, gwo, _, _ = _gw_pair(ctx, prob, sigma_integration="exact")
    gwo.rpa.setRPAInputEnergies(prob["energies"][sz.rpamin:sz.rpamax + 1])
    gw.PrepareScreening()
    gwo.sigma.PrepareScreening()
    rng = np.random.default_rng(14)
    levels = rng.integers(0, sz.qptotal, 24)
    freqs = rng.uniform(-1.5, 1.5, 24)
    val, der = gw.CalcCorrelationDiagElements(levels, freqs, derivative=True)
    ref = np.array([gwo.sigma.CalcCorrelationDiagElement(int(l), float(x)) for l, x in zip(levels, freqs)])
    refd = np.array([gwo.sigma.CalcCorrelationDiagElementDerivative(int(l), float(x)) for l, x in zip(levels, freqs)])
    np.testing.assert_allclose(val, ref, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(der, refd, rtol=1e-7, atol=1e-8)
    fr = prob["energies"][sz.qpmin:sz.qpmax + 1]
    np.testing.assert_allclose(gw.CalcCorrelationOffDiag(fr), gwo.sigma.CalcCorrelationOffDiag(fr), rtol=1e-8, atol=1e-10)


def test_g0w0_exact_qp_energies(ctx, prob):
    gw, gwo, _, _ = _gw_pair(ctx, prob, sigma_integration="exact", qp_grid_steps=201)
    gw.CalculateGWPerturbation()
    gwo.CalculateGWPerturbation()
    np.testing.assert_allclose(gw.getGWAResults(), gwo.getGWAResults(), rtol=0, atol=1e-6)
    gw.CalculateHQP()
    gwo.CalculateHQP()
    np.testing.assert_allclose(gw.getHQP(), gwo.getHQP(), rtol=0, atol=1e-6)


@pytest.mark.parametrize("scheme,order", [("legendre", 12), ("laguerre", 16), ("hermite", 10)])
def test_sigma_cda(ctx, prob, scheme, order):
    """Sigma_CDA: Gauss quadrature on the imaginary axis + residues of the enclosed poles + Gaussian tail."""
    sz = prob["sizes"]
    gw, gwo, _, _ = _gw_pair(ctx, prob, sigma_integration="cda", quadrature_scheme=scheme, order=order, alpha=1e-3)
    gwo.rpa.setRPAInputEnergies(prob["energies"][sz.rpamin:sz.rpamax + 1])
    gw.PrepareScreening()
    gwo.sigma.PrepareScreening()
    rng = np.random.default_rng(15)
    e = prob["energies"][sz.qpmin:sz.qpmax + 1]
    levels = rng.integers(0, sz.qptotal, 10)
    freqs = np.concatenate([rng.uniform(-1.2, 1.2, 8), e[levels[8:]] + 0.05])    # some far from, some near the poles
    val, der = gw.CalcCorrelationDiagElements(levels, freqs, derivative=True)
    ref = np.array([gwo.sigma.CalcCorrelationDiagElement(int(l), float(x)) for l, x in zip(levels, freqs)])
    np.testing.assert_allclose(val, ref, rtol=1e-8, atol=1e-10)
    refd = np.array([gwo.sigma.CalcCorrelationDiagElementDerivative(int(l), float(x)) for l, x in zip(levels[:4], freqs[:4])])
    np.testing.assert_allclose(der[:4], refd, rtol=1e-5, atol=1e-7)
    assert np.all(gw.CalcCorrelationOffDiag(e) == 0.0)       # Sigma_CDA has no off-diagonal correlation upstream


def test_g0w0_cda_qp_energies(ctx):
    """G0W0 with the CDA self-energy on the tiny problem (fixed-point QP solver; residues need one eps^-1 each)."""
    prob = synth.make_problem("tiny")
    sz = prob["sizes"]
    tc = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.Fill(prob["ao3c"], prob["C"], prob["aux_coulomb"])
    prob["tc_o"] = tc
    gw, gwo, _, _ = _gw_pair(ctx, prob, sigma_integration="cda", qp_solver="fixedpoint", qp_grid_steps=41)
    gw.CalculateGWPerturbation()
    gwo.CalculateGWPerturbation()
    np.testing.assert_allclose(gw.getGWAResults(), gwo.getGWAResults(), rtol=0, atol=1e-6)


@pytest.mark.parametrize("solver", ["grid", "fixedpoint"])
def test_g0w0_qp_energies(ctx, prob, solver):
    gw, gwo, _, _ = _gw_pair(ctx, prob, qp_solver=solver, qp_grid_steps=401)
    gw.CalculateGWPerturbation()
    gwo.CalculateGWPerturbation()
    np.testing.assert_allclose(gw.getGWAResults(), gwo.getGWAResults(), rtol=0, atol=1e-6)
    gw.CalculateHQP()
    gwo.CalculateHQP()
    np.testing.assert_allclose(gw.getHQP(), gwo.getHQP(), rtol=0, atol=1e-6)
    wq, _ = gw.DiagonalizeQPHamiltonian()
    np.testing.assert_allclose(wq, gwo.DiagonalizeQPHamiltonian()[0], rtol=0, atol=1e-6)


@pytest.mark.parametrize("mixing_order", [0, 1, 3])
def test_evgw(ctx, prob, mixing_order):
    """evGW iterates: plain update, linear mixing and Anderson mixing (upstream anderson_mixing.cc)."""
    gw, gwo, _, _ = _gw_pair(ctx, prob, gw_sc_max_iterations=5, qp_grid_steps=201, gw_mixing_order=mixing_order,
                             gw_mixing_alpha=0.7)
    gw.CalculateGWPerturbation()
    gwo.Mmn._fill_args = (prob["ao3c"], prob["C"], prob["aux_coulomb"], None, 5e-7)
    gwo.CalculateGWPerturbation()
    np.testing.assert_allclose(gw.getGWAResults(), gwo.getGWAResults(), rtol=0, atol=1e-6)
    np.testing.assert_allclose(gw.RPAInputEnergies(), gwo.RPAInputEnergies(), rtol=0, atol=1e-6)


@pytest.mark.parametrize("mode", ["dense", "dense-hx", "factorised"])
@pytest.mark.parametrize("name", list(orc.OPERATOR_TYPES))
def test_bse_operator_matmul_diagonal(ctx, prob, name, mode, monkeypatch):
    """BSE_OPERATOR<...>::matmul / diagonal / get_full_matrix vs the dense element-wise Hamiltonian, for the three
    device strategies: screened direct term materialised once in HBM with the rank-N_aux exchange term kept
    factorised (default when H fits), everything in the dense H (XTPB_BSE_HX_DENSE=1), and all products factorised."""
    from xtp_b200 import api
    monkeypatch.setenv("XTPB_BSE_DENSE_MAX_GB", "0" if mode == "factorised" else "32")
    monkeypatch.setenv("XTPB_BSE_HX_DENSE", "1" if mode == "dense-hx" else "0")
    sz = prob["sizes"]
    rng = np.random.default_rng(3)
    hq = rng.standard_normal((sz.vtotal + sz.ctotal,) * 2)
    hq = 0.5 * (hq + hq.T)
    eps_inv = rng.uniform(0.2, 1.0, sz.n_aux)
    cqp, cx, cd, cd2 = orc.OPERATOR_TYPES[name]
    tc = gpu_tc(ctx, prob)
    op = api.BSE_OPERATOR(ctx, cqp, cx, cd, cd2, eps_inv, tc, hq, sz.homo, sz.rpamin, sz.vmin, sz.cmax)
    ref = orc.BSE_OPERATOR(cqp, cx, cd, cd2, eps_inv, prob["tc_o"], hq)
    ref.configure(orc.BSEOperator_Options(sz.homo, sz.rpamin, sz.qpmin, sz.vmin, sz.cmax))
    H = ref.get_full_matrix()
    scale = np.abs(H).max()
    for k in (1, 5, 37):
        X = rng.standard_normal((op.rows(), k))
        assert np.abs(op.matmul(X) - H @ X).max() < 1e-11 * scale * np.sqrt(op.rows())
    assert np.abs(op.diagonal() - np.diag(H)).max() < 1e-12 * scale
    assert np.abs(op.get_full_matrix() - H).max() < 1e-12 * scale


@pytest.mark.parametrize("corr", ["DPR", "OLSEN"])
def test_davidson_dense_vs_eigh(ctx, corr):
    """test_davidson pattern of the reference: random diagonally dominant matrix vs a dense eigensolver."""
    from xtp_b200 import api
    rng = np.random.default_rng(7)
    n = 501
    A = rng.standard_normal((n, n)) * 0.01
    A = 0.5 * (A + A.T) + np.diag(np.sort(rng.uniform(0, 10, n)))
    op = api.DenseOperator(ctx, A)
    ds = api.DavidsonSolver()
    ds.set_tolerance("lapack"); ds.set_correction(corr); ds.set_max_search_space(80)
    ds.solve(op, 7)
    assert ds.info() == "Success"
    w, Z = np.linalg.eigh(A)
    np.testing.assert_allclose(ds.eigenvalues(), w[:7], atol=1e-9)
    ov = np.abs(np.einsum('ij,ij->j', ds.eigenvectors(), Z[:, :7]))
    np.testing.assert_allclose(ov, 1.0, atol=1e-7)
    # the oracle's Davidson on the same problem converges to the same values
    dso = orc.DavidsonSolver(); dso.set_tolerance("lapack"); dso.set_correction(corr); dso.set_max_search_space(80)

    class D:
        def rows(self): return n
        def diagonal(self): return np.diag(A).copy()
        def matmul(self, X): return A @ X
    dso.solve(D(), 7)
    np.testing.assert_allclose(ds.eigenvalues(), dso.eigenvalues(), atol=1e-9)


def test_davidson_restart_and_nonconvergence(ctx):
    from xtp_b200 import api
    rng = np.random.default_rng(8)
    n = 300
    A = rng.standard_normal((n, n)) * 0.05
    A = 0.5 * (A + A.T) + np.diag(np.sort(rng.uniform(0, 5, n)))
    op = api.DenseOperator(ctx, A)
    ds = api.DavidsonSolver()
    ds.set_tolerance("strict"); ds.set_max_search_space(24)      # forces restarts
    ds.solve(op, 4)
    assert ds.info() == "Success"
    np.testing.assert_allclose(ds.eigenvalues(), np.linalg.eigvalsh(A)[:4], atol=1e-6)
    ds2 = api.DavidsonSolver(); ds2.set_tolerance("lapack"); ds2.set_iter_max(2)
    ds2.solve(op, 4)
    assert ds2.info() == "NoConvergence" and ds2.num_iterations() == 2


def test_full_gwbse_step(ctx, prob):
    """Whole hot path on the GPU vs the oracle: Fill -> G0W0 (PPM) -> Hqp -> BSE singlets + triplets."""
    from xtp_b200 import api
    sz = prob["sizes"]
    nmax = 4
    gwopt = orc.GWOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax, qp_grid_steps=401)
    bseopt = orc.BSEOptions(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, nmax=nmax,
                            davidson_tolerance="lapack")
    ref = orc.run_gwbse(prob["ao3c"], prob["C"], prob["energies"], prob["vxc"], prob["aux_coulomb"], gwopt, bseopt,
                        triplets=True)
    tc = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.Fill(prob["ao3c"], prob["C"], prob["aux_coulomb"])
    gw = api.GW(ctx, tc, prob["vxc"], prob["energies"])
    gw.configure(api.gw_options(homo=sz.homo, qpmin=sz.qpmin, qpmax=sz.qpmax, rpamin=sz.rpamin, rpamax=sz.rpamax,
                                qp_grid_steps=401))
    gw.CalculateGWPerturbation()
    np.testing.assert_allclose(gw.getGWAResults(), ref["qp_pert"], rtol=0, atol=1e-6)
    gw.CalculateHQP()
    bse = api.BSE(ctx, tc)
    bse.configure(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, nmax, gw.RPAInputEnergies(),
                  gw.getHQP(), davidson_tolerance="lapack")
    np.testing.assert_allclose(np.sort(bse.epsilon_0_inv()), np.sort(ref["eps0_inv"]), rtol=1e-8)
    es, vs = bse.Solve_singlets_TDA()
    assert bse.last_davidson.info() == "Success"
    np.testing.assert_allclose(es, ref["singlet_energies"], rtol=0, atol=1e-6)
    et, _ = bse.Solve_triplets_TDA()
    np.testing.assert_allclose(et, ref["triplet_energies"], rtol=0, atol=1e-6)
    # eigenvectors: same subspace (sign/phase free)
    ov = np.abs(np.einsum('ij,ij->j', vs, ref["singlet_vectors"]))
    gaps = np.diff(ref["singlet_energies"])
    if gaps.min() > 1e-4:
        np.testing.assert_allclose(ov, 1.0, atol=1e-5)


def _bse_pair(ctx, prob, nmax=3):
    from xtp_b200 import api
    sz = prob["sizes"]
    gw, gwo, tc, tco = _gw_pair(ctx, prob, qp_grid_steps=201)
    gw.CalculateGWPerturbation()
    gw.CalculateHQP()
    gwo.CalculateGWPerturbation()
    gwo.CalculateHQP()
    bse = api.BSE(ctx, tc)
    bse.configure(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, nmax, gw.RPAInputEnergies(),
                  gw.getHQP(), davidson_tolerance="lapack", davidson_maxiter=200)
    bseo = orc.BSE(tco)
    bseo.configure(orc.BSEOptions(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, nmax=nmax,
                                  davidson_tolerance="lapack"), gwo.RPAInputEnergies(), gwo.getHQP())
    bseo._rpa_energies_for_tests = gwo.RPAInputEnergies().copy()
    return bse, bseo


@pytest.mark.parametrize("singlet", [True, False])
def test_full_bse_btda(ctx, prob, singlet):
    """Full (non-TDA) BSE: subspace solver on the device vs the dense symmetric reduction of [[A,B],[-B,-A]]."""
    bse, bseo = _bse_pair(ctx, prob)
    e, X, Y = bse.Solve_singlets_BTDA() if singlet else bse.Solve_triplets_BTDA()
    eo, Xo, Yo = bseo.solve_btda_dense(singlet)
    assert bse.last_btda["info"] == "Success"
    np.testing.assert_allclose(e, eo, rtol=0, atol=1e-7)
    np.testing.assert_allclose(np.diag(X.T @ X - Y.T @ Y), 1.0, atol=1e-8)
    # eigenvectors up to sign (energies of the tiny problems are non-degenerate)
    for j in range(len(e)):
        sgn = np.sign(X[:, j] @ Xo[:, j])
        np.testing.assert_allclose(sgn * X[:, j], Xo[:, j], atol=1e-5)
        np.testing.assert_allclose(sgn * Y[:, j], Yo[:, j], atol=1e-5)
    # the TDA energies lie above the full-BSE ones
    etda, _ = bse.Solve_singlets_TDA() if singlet else bse.Solve_triplets_TDA()
    assert np.all(e <= etda + 1e-9)


def test_transition_dipoles_and_oscillator_strengths(ctx, prob):
    """BSE::CalcCoupledTransition_Dipoles / Orbitals::Oscillatorstrengths (north_star: 1e-5 relative)."""
    from xtp_b200 import api
    sz = prob["sizes"]
    bse, bseo = _bse_pair(ctx, prob)
    rng = np.random.default_rng(21)
    r = rng.standard_normal((3, sz.n_basis, sz.n_basis))
    r = 0.5 * (r + np.transpose(r, (0, 2, 1)))
    e, X = bse.Solve_singlets_TDA()
    d = bse.transition_dipoles(r, prob["C"], X)
    dref = orc.BSE.transition_dipoles(r, prob["C"], sz.homo, sz.vmin, sz.cmax, X)
    np.testing.assert_allclose(d, dref, rtol=1e-10, atol=1e-12)
    f = api.oscillator_strengths(e, d)
    np.testing.assert_allclose(f, orc.BSE.oscillator_strengths(e, dref), rtol=1e-10, atol=1e-14)
    # against the oracle's own eigenvectors: oscillator strengths within 1e-5 relative (sign-invariant)
    eo, Xo = bseo.Solve_singlets_TDA()
    fo = orc.BSE.oscillator_strengths(eo, orc.BSE.transition_dipoles(r, prob["C"], sz.homo, sz.vmin, sz.cmax, Xo))
    np.testing.assert_allclose(f, fo, rtol=1e-5, atol=1e-9)
    # full BSE: X + Y enters
    eb, Xb, Yb = bse.Solve_singlets_BTDA()
    db = bse.transition_dipoles(r, prob["C"], Xb, Yb)
    np.testing.assert_allclose(db, orc.BSE.transition_dipoles(r, prob["C"], sz.homo, sz.vmin, sz.cmax, Xb, Yb),
                               rtol=1e-10, atol=1e-12)


def test_plot_sigma(ctx, prob):
    """GW::PlotSigma table (frequency, Sigma_c + e_KS + Sigma_x - Vxc per state) against the oracle."""
    gw, gwo, _, _ = _gw_pair(ctx, prob, qp_grid_steps=201)
    gw.CalculateGWPerturbation()
    gwo.CalculateGWPerturbation()
    sz = prob["sizes"]
    states = [0, sz.homo - sz.qpmin, sz.homo + 1 - sz.qpmin]
    tab = gw.PlotSigma(41, 0.02, states)
    ref = gwo.PlotSigma(41, 0.02, states)
    np.testing.assert_allclose(tab[:, 0::2], ref[:, 0::2], rtol=0, atol=1e-12)
    np.testing.assert_allclose(tab[:, 1::2], ref[:, 1::2], rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize("tda", [True, False])
def test_perturbative_dynamical_screening(ctx, prob, tda):
    """BSE::Perturbative_DynamicalScreening against the oracle, TDA and full BSE singlets."""
    bse, bseo = _bse_pair(ctx, prob, nmax=3)
    sz = prob["sizes"]
    if tda:
        e, X = bse.Solve_singlets_TDA()
        Y = None
    else:
        e, X, Y = bse.Solve_singlets_BTDA()
    dyn = bse.Perturbative_DynamicalScreening(e, X, Y)
    ref, its = bseo.Perturbative_DynamicalScreening(bseo._rpa_energies_for_tests, e, X, Y)
    np.testing.assert_allclose(dyn, ref, rtol=0, atol=2e-6)
    np.testing.assert_array_equal(bse.dynamical_iterations, its)
    assert np.abs(dyn - e).max() > 1e-7


@pytest.mark.parametrize("ranges", ["default", "explicit", "explicit-low", "explicit-high"])
def test_gwbse_driver_evaluate(ctx, ranges):
    """GWBSE::Initialize + Evaluate (upstream gwbse/gwbse.cc): the driver mirror runs Fill -> G0W0 -> Hqp -> BSE singlets
    and triplets in one call, with the level ranges coming from the `ranges` option, against the oracle's whole step.
    `explicit` makes the BSE window stick out of the QP window on both sides (BSE::AdjustHqpSize extends Hqp with the
    RPA input energies)."""
    from xtp_b200 import api
    prob = synth.make_problem("ch4-svp-shape")
    sz = prob["sizes"]
    # explicit-low: vmin < qpmin AND cmax < qpmax (only part of the QP block lies inside the BSE window -- the case the
    # round-1 AdjustHqpSize wrote out of bounds for); explicit-high: vmin > qpmin and cmax > qpmax
    kw = {"default": {}, "explicit": dict(rpamax=sz.n_basis - 1, qpmin=1, qpmax=8, bsemin=0, bsemax=10),
          "explicit-low": dict(rpamax=sz.n_basis - 1, qpmin=2, qpmax=9, bsemin=0, bsemax=7),
          "explicit-high": dict(rpamax=sz.n_basis - 1, qpmin=0, qpmax=7, bsemin=2, bsemax=11)}[ranges]
    ranges = ranges.split("-")[0]
    drv = api.GWBSE(ctx).Initialize(sz.n_basis, sz.homo + 1, ranges=ranges, tasks=("gw", "singlets", "triplets"), nmax=3,
                                    davidson_tolerance="lapack", **kw)
    r = orc.gwbse_level_ranges(ranges, sz.n_basis, sz.homo + 1, **kw)
    vxc = np.ascontiguousarray(prob["vxc"][r["qpmin"]:r["qpmax"] + 1, r["qpmin"]:r["qpmax"] + 1])   # QP-window block
    out = drv.Evaluate(prob["ao3c"], prob["C"], prob["energies"], vxc, prob["aux_coulomb"])
    assert {k: out["ranges"][k] for k in r} == r
    gwopt = orc.GWOptions(r["homo"], r["qpmin"], r["qpmax"], r["rpamin"], r["rpamax"])
    bseopt = orc.BSEOptions(r["homo"], r["rpamin"], r["rpamax"], r["qpmin"], r["qpmax"], r["vmin"], r["cmax"], nmax=3,
                            davidson_tolerance="lapack")
    ref = orc.run_gwbse(prob["ao3c"], prob["C"], prob["energies"], vxc, prob["aux_coulomb"], gwopt, bseopt,
                        triplets=True)
    np.testing.assert_allclose(out["QPpert_energies"], ref["qp_pert"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["Hqp"], ref["Hqp"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["QPdiag_energies"], ref["qp_diag"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["BSE_singlet_energies"], ref["singlet_energies"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["BSE_triplet_energies"], ref["triplet_energies"], rtol=0, atol=1e-6)
    assert out["BSE_singlet_coefficients"].shape == ((r["homo"] - r["vmin"] + 1) * (r["cmax"] - r["homo"]), 3)


def test_bse_reuses_ppm_eigenbasis_of_epsilon0(ctx, prob, monkeypatch):
    """G0W0 + PPM leaves the tensor in the eigenbasis of eps(0) at the RPA input energies BSE::configure is then given:
    the library reads the eigenvalues instead of recomputing eps(0) + eigensolver + rotation.  Both ways must give the
    same screening and the same excitation energies (and the oracle's, which recomputes like upstream)."""
    from xtp_b200 import api
    sz = prob["sizes"]
    out = {}
    for reuse in ("1", "0"):
        monkeypatch.setenv("XTPB_BSE_REUSE_EPS0", reuse)
        gw, gwo, tc, tco = _gw_pair(ctx, prob, qp_grid_steps=201)
        gw.CalculateGWPerturbation()
        gw.CalculateHQP()
        bse = api.BSE(ctx, tc)
        bse.configure(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, 3, gw.RPAInputEnergies(),
                      gw.getHQP(), davidson_tolerance="lapack", davidson_maxiter=200)
        assert bse.eps0_reused() == (reuse == "1")
        out[reuse] = (bse.epsilon_0_inv(), bse.Solve_singlets_TDA()[0], bse.Solve_triplets_TDA()[0])
        bse.close()
    np.testing.assert_allclose(out["1"][0], out["0"][0], rtol=1e-9)
    np.testing.assert_allclose(out["1"][1], out["0"][1], rtol=0, atol=1e-9)
    np.testing.assert_allclose(out["1"][2], out["0"][2], rtol=0, atol=1e-9)
    # a rotation of the tensor, or other energies, invalidate the shortcut
    gw, gwo, tc, tco = _gw_pair(ctx, prob, qp_grid_steps=201)
    monkeypatch.setenv("XTPB_BSE_REUSE_EPS0", "1")
    gw.CalculateGWPerturbation()
    gw.CalculateHQP()
    e = gw.RPAInputEnergies().copy()
    e[-1] += 1e-3
    bse = api.BSE(ctx, tc)
    bse.configure(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, 3, e, gw.getHQP())
    assert not bse.eps0_reused()
    bse.close()
    tc.MultiplyRightWithAuxMatrix(np.eye(sz.n_aux))
    bse.configure(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, 3, gw.RPAInputEnergies(),
                  gw.getHQP())
    assert not bse.eps0_reused()
    bse.close()


def test_coulomb_metric_prefetch_matches_synchronous_path(ctx, prob):
    """xtpb_tc_coulomb_metric_begin starts the first eigendecomposition of the metric step on the helper thread before
    Fill3cMO (what TCMatrix_gwbse.Fill does); the tensor must equal the one of the synchronous call sequence, with and
    without an aux overlap, and mismatched matrices must be refused."""
    from xtp_b200 import api
    sz = prob["sizes"]
    rng = np.random.default_rng(9)
    B = rng.standard_normal((sz.n_aux, sz.n_aux))
    S = B @ B.T / sz.n_aux + 0.5 * np.eye(sz.n_aux)
    for overlap in (None, S):
        a = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
        a.Fill(prob["ao3c"], prob["C"], prob["aux_coulomb"], overlap)           # prefetch inside
        b = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
        b.Fill3cMO(prob["ao3c"], prob["C"])
        b.apply_coulomb_metric(prob["aux_coulomb"], overlap)                    # no prefetch
        ref = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
        ref.Fill(prob["ao3c"], prob["C"], prob["aux_coulomb"], overlap)
        assert rel(a.get_raw(), b.get_raw()) < 1e-11
        assert rel(a.get_raw(), ref.M) < 1e-9
    c = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    c.coulomb_metric_begin(prob["aux_coulomb"])
    with pytest.raises(ValueError):
        c.apply_coulomb_metric(S)


def test_block_cache_reuses_scratch_without_changing_results(tmp_path):
    """The exact-size block cache (library default; XTPB_ALLOC_CACHE=0 turns it off): released scratch blocks are handed out again instead of
    going back to the driver.  The switch is read when the library loads, so the check runs in a child process: two
    identical G0W0+BSE steps; the second must be served from the cache and give the same energies (1e-12), equal to the
    oracle's within the usual bounds."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import json, sys
sys.path.insert(0, %r)
import numpy as np
from xtp_b200 import api, synth
prob = synth.make_problem("ch4-svp-shape")
sz = prob["sizes"]
ctx = api.Context(0)
res = []
for step in range(2):
    api.alloc_stats(reset=True)
    drv = api.GWBSE(ctx).Initialize(sz.n_basis, sz.homo + 1, tasks=("gw", "singlets"), nmax=3, davidson_tolerance="lapack")
    out = drv.Evaluate(prob["ao3c"], prob["C"], prob["energies"], prob["vxc"], prob["aux_coulomb"])
    st = api.alloc_stats()
    res.append({"qp": out["QPpert_energies"].tolist(), "s": out["BSE_singlet_energies"].tolist(), "calls": st["calls"],
                "hits": st["cache_hits"]})
print("RESULT " + json.dumps(res))
""" % root
    env = dict(os.environ, XTPB_ALLOC_CACHE="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    np.testing.assert_allclose(res[1]["qp"], res[0]["qp"], rtol=0, atol=1e-12)     # fresh blocks vs recycled blocks
    np.testing.assert_allclose(res[1]["s"], res[0]["s"], rtol=0, atol=1e-12)
    assert res[1]["calls"] > 0 and res[1]["hits"] >= 0.5 * res[1]["calls"]
    prob = synth.make_problem("ch4-svp-shape")
    sz = prob["sizes"]
    gwopt = orc.GWOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax)
    bseopt = orc.BSEOptions(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, nmax=3,
                            davidson_tolerance="lapack")
    ref = orc.run_gwbse(prob["ao3c"], prob["C"], prob["energies"], prob["vxc"], prob["aux_coulomb"], gwopt, bseopt)
    np.testing.assert_allclose(res[1]["qp"], ref["qp_pert"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(res[1]["s"], ref["singlet_energies"], rtol=0, atol=1e-6)


def test_bse_operator_properties_at_scale(ctx, monkeypatch):
    """Size-independent properties at a size the numpy oracle would need minutes for (synth-500 shape: 1500 aux
    functions, BSE size 2500): the three device strategies of BSE_OPERATOR::matmul (screened direct term dense +
    factorised exchange, fully dense H, all factorised) agree with each other, and the operator is linear and symmetric."""
    from xtp_b200 import api
    sz = synth.WORKLOADS["synth-500"]
    rng = np.random.default_rng(12)
    tc = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.set_raw(synth.make_M_direct(sz, rng))
    hs = sz.vtotal + sz.ctotal
    hq = rng.standard_normal((hs, hs)) * 0.05
    hq = 0.5 * (hq + hq.T) + np.diag(np.sort(rng.uniform(-1.0, 2.0, hs)))
    eps_inv = rng.uniform(0.2, 1.0, sz.n_aux)
    n = sz.bse_size
    X, Y = rng.standard_normal((n, 7)), rng.standard_normal((n, 7))
    results = {}
    for mode in ("dense", "dense-hx", "factorised"):
        monkeypatch.setenv("XTPB_BSE_DENSE_MAX_GB", "0" if mode == "factorised" else "32")
        monkeypatch.setenv("XTPB_BSE_HX_DENSE", "1" if mode == "dense-hx" else "0")
        op = api.BSE_OPERATOR(ctx, 1, 2, 1, 0, eps_inv, tc, hq, sz.homo, sz.rpamin, sz.vmin, sz.cmax)   # singlet TDA
        hx, hy = op.matmul(X), op.matmul(Y)
        results[mode] = (hx, op.diagonal())
        scale = np.abs(hx).max()
        assert np.abs(op.matmul(X - 0.7 * Y) - (hx - 0.7 * hy)).max() < 1e-10 * scale          # linear
        assert np.abs(Y.T @ hx - hy.T @ X).max() < 1e-9 * scale * np.sqrt(n)                  # symmetric
        op.close()
    scale = np.abs(results["dense"][0]).max()
    for mode in ("dense-hx", "factorised"):
        assert np.abs(results[mode][0] - results["dense"][0]).max() < 1e-10 * scale * np.sqrt(n)
        assert np.abs(results[mode][1] - results["dense"][1]).max() < 1e-11 * np.abs(results["dense"][1]).max()
