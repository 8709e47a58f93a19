"""Property tests of the CPU oracle (hypothesis): invariances that hold for ANY input of
the right shape -- permutations and scalings of the tensor, linearity and symmetry of the BSE operators, ordering of
the dielectric matrix on the imaginary axis.  They pin the oracle where no reference vectors exist (parity unpinned)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import gwbse_oracle as orc
from xtp_b200 import synth


def _random_problem(seed, n_basis, n_aux, homo):
    sz = synth.Sizes(n_basis=n_basis, n_aux=n_aux, homo=homo)
    rng = np.random.default_rng(seed)
    M = synth.make_M_direct(sz, rng) * 3.0
    e = synth.make_energies(sz, rng)
    tc = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.set_raw(M)
    return sz, tc, e, rng


shapes = st.tuples(st.integers(0, 10 ** 6), st.integers(8, 14), st.integers(5, 12), st.integers(1, 3))


def _rpa(sz, tc, e):
    rpa = orc.RPA(tc)
    rpa.configure(sz.homo, sz.rpamin, sz.rpamax)
    rpa.setRPAInputEnergies(e[sz.rpamin:sz.rpamax + 1])
    return rpa


@settings(max_examples=25, deadline=None)
@given(shapes, st.floats(0.0, 3.0), st.floats(0.05, 3.0))
def test_epsilon_imaginary_axis_ordering(shape, w, dw):
    sz, tc, e, _ = _random_problem(*shape)
    rpa = _rpa(sz, tc, e)
    a, b = rpa.calculate_epsilon_i(w), rpa.calculate_epsilon_i(w + dw)
    assert np.abs(a - a.T).max() < 1e-12
    assert np.linalg.eigvalsh(b).min() >= 1.0 - 1e-10           # eps(i w) >= 1
    assert np.linalg.eigvalsh(a - b).min() >= -1e-10            # and decreases with w


@settings(max_examples=20, deadline=None)
@given(shapes, st.floats(0.0, 2.0))
def test_epsilon_invariant_under_permutation_of_unoccupied_levels(shape, w):
    """The sum over (occupied, unoccupied) pairs does not care about the order of the levels: permute the unoccupied
    columns of every slab together with their energies."""
    sz, tc, e, rng = _random_problem(*shape)
    ref = _rpa(sz, tc, e).calculate_epsilon_i(w)
    perm = np.arange(sz.ntotal)
    perm[sz.n_occ:] = sz.n_occ + rng.permutation(sz.ntotal - sz.n_occ)
    tc2 = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc2.set_raw(np.ascontiguousarray(tc.M[:, :, perm]))
    got = _rpa(sz, tc2, e[perm]).calculate_epsilon_i(w)
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-13)


@settings(max_examples=20, deadline=None)
@given(shapes, st.floats(0.3, 3.0))
def test_sigma_x_scales_quadratically_and_is_negative_semidefinite(shape, s):
    sz, tc, e, _ = _random_problem(*shape)
    rpa = _rpa(sz, tc, e)
    sig = orc.Sigma_PPM(tc, rpa)
    sig.configure(orc.SigmaOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax))
    sx = sig.CalcExchangeMatrix()
    assert np.linalg.eigvalsh(0.5 * (sx + sx.T)).max() <= 1e-12
    tc2 = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc2.set_raw(tc.M * s)
    sig2 = orc.Sigma_PPM(tc2, _rpa(sz, tc2, e))
    sig2.configure(orc.SigmaOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax))
    np.testing.assert_allclose(sig2.CalcExchangeMatrix(), s * s * sx, rtol=1e-12, atol=1e-14)


@settings(max_examples=15, deadline=None)
@given(shapes, st.sampled_from(sorted(orc.OPERATOR_TYPES)), st.floats(-2.0, 2.0))
def test_bse_operator_is_linear_and_symmetric(shape, name, alpha):
    sz, tc, e, rng = _random_problem(*shape)
    hs = sz.vtotal + sz.ctotal
    hq = rng.standard_normal((hs, hs))
    hq = 0.5 * (hq + hq.T)
    cqp, cx, cd, cd2 = orc.OPERATOR_TYPES[name]
    op = orc.BSE_OPERATOR(cqp, cx, cd, cd2, rng.uniform(0.2, 1.0, sz.n_aux), tc, hq)
    op.configure(orc.BSEOperator_Options(sz.homo, sz.rpamin, sz.qpmin, sz.vmin, sz.cmax))
    n = op.rows()
    X, Y = rng.standard_normal((n, 3)), rng.standard_normal((n, 3))
    np.testing.assert_allclose(op.matmul(X + alpha * Y), op.matmul(X) + alpha * op.matmul(Y), rtol=1e-10, atol=1e-11)
    # <Y, H X> == <H Y, X>: every operator type is symmetric (M real, Hqp symmetric)
    np.testing.assert_allclose(Y.T @ op.matmul(X), op.matmul(Y).T @ X, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(op.diagonal(), np.diag(op.get_full_matrix()), rtol=1e-10, atol=1e-12)


@settings(max_examples=15, deadline=None)
@given(shapes, st.floats(-1.5, 1.5))
def test_sigma_c_ppm_invariant_under_aux_sign_flips(shape, w):
    """M[:, P, :] -> -M[:, P, :] for any subset of aux functions is an orthogonal aux rotation: epsilon's spectrum, the
    plasmon-pole parameters (as sets) and every Sigma_c element must not change."""
    sz, tc, e, rng = _random_problem(*shape)

    def sigma(t):
        s = orc.Sigma_PPM(t, _rpa(sz, t, e))
        s.configure(orc.SigmaOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax))
        s.PrepareScreening()
        return np.array([s.CalcCorrelationDiagElement(l, w) for l in range(sz.qptotal)]), np.sort(s.ppm.ppm_freq)

    flips = rng.choice([-1.0, 1.0], sz.n_aux)
    tc2 = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc2.set_raw(tc.M * flips[None, :, None])
    v1, f1 = sigma(tc)
    v2, f2 = sigma(tc2)
    np.testing.assert_allclose(f2, f1, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(v2, v1, rtol=1e-7, atol=1e-9)
