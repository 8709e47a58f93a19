"""Closed-form GW-BSE results that do not depend on the (absent) reference source: the pins of the oracle and, in
the -m gpu suite, of the CUDA path itself.

1. Two-level system (one occupied level v, one empty level c, any number of aux functions).  chi0 has a single
   transition D = e_c - e_v, so with m_P = M_vc^P:
       eps(i w) = 1 + m m^T 4 D / (D^2 + w^2)                       (rank one)
       eps^-1(i w) - 1 = - mhat mhat^T 4 D |m|^2 / (w^2 + W^2),      W^2 = D^2 + 4 D |m|^2
   i.e. the screened interaction has exactly ONE plasmon pole at W.  Consequences:
     * the plasmon-pole model is exact: weight 1 - 1/lambda = 4 D |m|^2 / W^2, frequency W;
     * Sigma_c,nn(w) = (2 D / W) sum_{l in {v,c}} (M_nl . m)^2 / (w - e_l +- W)    (+ occupied, - empty),
       the same number from Sigma_PPM, Sigma_Exact (eta -> 0) and Sigma_CDA (order -> inf);
     * BSE (size 1): A = Hqp_cc - Hqp_vv + cx (m.m) - [a.b - d0 (a.m)(b.m) / (1 + d0 |m|^2)],  d0 = 4 D / (D^2 + eta^2),
       a = M_vv, b = M_cc;  B = cx (m.m) - [m.m - d0 |m|^4 / (1 + d0 |m|^2)];  TDA energy A, full BSE sqrt(A^2 - B^2);
       singlet cx = 2, triplet cx = 0.
2. For ANY system and a frequency w inside the HOMO-LUMO gap (no pole of G enclosed), the correlation self-energy is
   the imaginary-axis integral
       Sigma_c,nn(w) = -(1/pi) sum_m int_0^inf dw' (w - e_m) / ((w - e_m)^2 + w'^2)  M_nm^T [eps^-1(i w') - 1] M_nm
   (contour deformation of (i/2pi) int G W_c; e.g. Golze et al., JCTC 14, 4856 (2018), eq. 21-23 without residues),
   evaluated here with scipy's adaptive quadrature and a dense inverse per node -- no plasmon poles, no RPA
   eigenmodes, no prefactor shared with the code under test.
"""
import math

import numpy as np
from scipy import integrate


def two_level_system(n_aux=7, seed=5, hqp_gap=0.83):
    """Random two-level problem in the oracle's conventions: M[m, P, n] with levels (v, c) = (0, 1)."""
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((2, n_aux, 2)) * 0.35
    M = 0.5 * (M + M.transpose(2, 1, 0))                 # (mn|P) = (nm|P)
    energies = np.array([-0.45, 0.20])
    hqp = np.array([[-0.52, 0.0], [0.0, -0.52 + hqp_gap]])
    return {"M": np.ascontiguousarray(M), "energies": energies, "hqp": hqp, "n_aux": n_aux}


def two_level_sigma_c(sysm, level, omega):
    """Analytic Sigma_c,nn(omega) of the two-level system (eta = 0)."""
    M, e = sysm["M"], sysm["energies"]
    m = M[0, :, 1]
    D = e[1] - e[0]
    m2 = float(m @ m)
    W = math.sqrt(D * D + 4.0 * D * m2)
    pref = 2.0 * D / W
    s = 0.0
    for l, sign in ((0, +1.0), (1, -1.0)):
        proj = float(M[level, :, l] @ m)
        s += proj * proj / (omega - e[l] + sign * W)
    return pref * s


def two_level_ppm_parameters(sysm):
    """(weight, frequency) of the only plasmon pole with non-zero weight."""
    M, e = sysm["M"], sysm["energies"]
    m = M[0, :, 1]
    D = e[1] - e[0]
    m2 = float(m @ m)
    W2 = D * D + 4.0 * D * m2
    return 4.0 * D * m2 / W2, math.sqrt(W2)


def two_level_bse(sysm, eta=1e-3):
    """{'singlet_tda', 'triplet_tda', 'singlet_full', 'triplet_full'} excitation energies."""
    M, e, h = sysm["M"], sysm["energies"], sysm["hqp"]
    m, a, b = M[0, :, 1], M[0, :, 0], M[1, :, 1]
    D = e[1] - e[0]
    d0 = 4.0 * D / (D * D + eta * eta)
    m2 = float(m @ m)
    scr = d0 / (1.0 + d0 * m2)
    hd = float(a @ b) - scr * float(a @ m) * float(b @ m)       # sum_PQ M_vv eps^-1 M_cc
    hd2 = m2 - scr * m2 * m2                                    # sum_PQ M_vc eps^-1 M_cv
    out = {}
    for name, cx in (("singlet", 2.0), ("triplet", 0.0)):
        A = h[1, 1] - h[0, 0] + cx * m2 - hd
        B = cx * m2 - hd2
        out[name + "_tda"] = A
        out[name + "_full"] = math.sqrt(A * A - B * B)
    return out


def sigma_c_imaginary_axis(M_level, M_occ_unocc, energies, n_occ, omega, limit=400):
    """Imaginary-axis integral of the module docstring for one level.
    M_level[P, m]: slab of the level; M_occ_unocc[i, P, a]: M_ia^P for i occupied, a empty; omega inside the gap."""
    e = np.asarray(energies)
    naux = M_level.shape[0]
    dE = e[n_occ:][None, :] - e[:n_occ][:, None]
    A = M_occ_unocc.transpose(1, 0, 2).reshape(naux, -1)        # [P, (i,a)]
    dEf = dE.reshape(-1)
    a_m = omega - e

    def integrand(wp):
        eps = np.eye(naux) + (A * (4.0 * dEf / (dEf * dEf + wp * wp))) @ A.T
        K = np.linalg.inv(eps) - np.eye(naux)
        q = np.einsum("pm,pm->m", M_level, K @ M_level)
        return float((a_m / (a_m * a_m + wp * wp) * q).sum())

    val, err = integrate.quad(integrand, 0.0, np.inf, limit=limit, epsabs=1e-12, epsrel=1e-11)
    return -val / math.pi, err / math.pi


def two_level_dynamical_screening(sysm, e_static, eta=1e-3, max_iter=10, tol=1e-5):
    """Perturbative dynamical screening of the two-level TDA exciton (coefficient vector = [1]):
    E <- E_static - [hd(E) - hd(0)],  hd(w) = a.b - d(w) (a.m)(b.m) / (1 + d(w) |m|^2),
    d(w) = 2 [(D - w)/((D - w)^2 + eta^2) + (D + w)/((D + w)^2 + eta^2)]   (chi0 weight on the real axis)."""
    M, en = sysm["M"], sysm["energies"]
    m, a, b = M[0, :, 1], M[0, :, 0], M[1, :, 1]
    D = en[1] - en[0]
    m2, am, bm, ab = float(m @ m), float(a @ m), float(b @ m), float(a @ b)

    def hd(w):
        d = 2.0 * ((D - w) / ((D - w) ** 2 + eta * eta) + (D + w) / ((D + w) ** 2 + eta * eta))
        lam = 1.0 + d * m2                      # the one eigenvalue of eps(w) that is not 1
        inv = 1.0 / lam if lam > 1e-8 else 0.0  # BSE::SetupDirectInteractionOperator drops eigenvalues <= 1e-8 (w > D)
        return ab - (1.0 - inv) * am * bm / m2

    e, its = e_static, 0
    for its in range(1, max_iter + 1):
        old = e
        e = e_static - (hd(old) - hd(0.0))
        if abs(e - old) < tol:
            break
    return e, its
