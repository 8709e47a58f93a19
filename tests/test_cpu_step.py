"""The complete CPU step behind bench.py's CPU baseline (oracle/cpu_step.py: reference-structure and same-algorithm
variants, OpenMP C kernels of oracle/cpu_gw.c) against the numpy oracle's whole step, and the sampled estimate of
oracle/cpu_reference.py against a full run of the same shape."""
import numpy as np
import pytest

from oracle import cpu_reference as cr
from oracle import cpu_step
from oracle import gwbse_oracle as orc
from xtp_b200 import synth


@pytest.fixture(scope="module", params=["tiny", "ch4-svp-shape"])
def case(request):
    prob = synth.make_problem(request.param)
    sz = prob["sizes"]
    gwopt = orc.GWOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax, qp_grid_steps=401)
    bseopt = orc.BSEOptions(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, nmax=3,
                            davidson_tolerance="lapack")
    ref = orc.run_gwbse(prob["ao3c"], prob["C"], prob["energies"], prob["vxc"], prob["aux_coulomb"], gwopt, bseopt)
    return prob, ref


@pytest.mark.parametrize("algorithm", ["reference", "factorised"])
def test_full_cpu_step_matches_oracle(case, algorithm):
    prob, ref = case
    out = cpu_step.run_step(prob, algorithm=algorithm, nmax=3, grid_steps=401, davidson_tolerance="lapack")
    np.testing.assert_allclose(out["qp"], ref["qp_pert"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["singlets"], ref["singlet_energies"], rtol=0, atol=2e-6)
    assert set(out["stage_seconds"]) == {"fill", "metric", "epsilon", "ppm", "sigma_x", "sigma_c", "offdiag",
                                         "bse_setup", "davidson"}
    assert out["matmul_calls"] >= 1 and out["threads"] >= 1
    if algorithm == "reference":
        assert out["sigma_c_evaluations"] >= prob["sizes"].qptotal * 401


def test_batched_grid_scan_equals_pointwise_kernel():
    """cpu_gw.c: sigma_ppm_grid_batched (one pass per slab, frequencies in registers) == sigma_ppm_diag point by point,
    including grid points inside the damped window of a pole."""
    import ctypes as C
    rng = np.random.default_rng(2)
    q, na, nt, nocc, steps = 3, 17, 23, 9, 101
    M = rng.standard_normal((q, na, nt)) * 0.3
    e = np.concatenate([np.sort(rng.uniform(-1.0, -0.3, nocc)), np.sort(rng.uniform(0.0, 2.0, nt - nocc))])
    freq = rng.uniform(0.3, 1.5, na)
    fac = rng.uniform(0.05, 0.5, na)
    fac[3] = 0.0
    om0 = np.array([-1.2, -0.4, 0.3])
    vals = np.empty((q, steps))
    lib = cpu_step._lib()
    lib.sigma_ppm_grid_batched(cpu_step._p(M), na * nt, nt, nt, na, nocc, cpu_step._p(e), cpu_step._p(freq),
                               cpu_step._p(fac), q, cpu_step._p(om0), 0.02, steps, cpu_step._p(vals))
    for l in range(q):
        want = cr.sigma_ppm_diag(M[l], nocc, e, freq, fac, om0[l] + 0.02 * np.arange(steps))
        np.testing.assert_allclose(vals[l], want, rtol=1e-11, atol=1e-12)
