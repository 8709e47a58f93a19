"""Real-molecule inputs (xtp_b200/molecule.py: Gaussian integrals, STO-3G, even-tempered aux basis, RHF) -- the stand-in
for XTP's integral/SCF layers that feed the GW-BSE path.  Unlike the GW-BSE oracle these
numbers CAN be pinned to the literature: the STO-3G Hartree-Fock energies, orbital energies and dipole moment of H2 and
H2O are textbook values."""
import numpy as np
import pytest

from oracle import gwbse_oracle as orc
from xtp_b200 import molecule as ml


@pytest.fixture(scope="module")
def water_inputs():
    return ml.gwbse_inputs(ml.water())


def test_h2_sto3g_rhf_energy():
    h2 = ml.Molecule([("H", (0, 0, 0)), ("H", (0, 0, 1.4))])
    assert abs(ml.rhf(h2, h2.sto3g())["energy"] - (-1.1167)) < 5e-5          # Szabo & Ostlund, table 3.x: -1.1167 Ha


def test_water_sto3g_rhf_matches_literature(water_inputs):
    w = ml.water()
    assert abs(w.nuclear_repulsion() - 8.002367061810450) < 1e-10
    scf = water_inputs["scf"]
    assert abs(scf["energy"] - (-74.942079928192)) < 1e-8
    np.testing.assert_allclose(scf["eps"], [-20.262892, -1.209697, -0.547965, -0.436527, -0.387587, 0.477619, 0.588139],
                               atol=2e-6)
    C, n_occ = water_inputs["C"], water_inputs["n_occ"]
    np.testing.assert_allclose(C.T @ scf["S"] @ C, np.eye(7), atol=1e-10)
    D = 2.0 * C[:, :n_occ] @ C[:, :n_occ].T
    mu = -np.einsum("kmn,mn->k", water_inputs["ao_dipoles"], D) + sum(ml._Z[s] * r for s, r in w.atoms)
    np.testing.assert_allclose(mu, [0.0, 0.603521296525, 0.0], atol=1e-8)


def test_three_centre_integrals_are_an_ri_factorisation(water_inputs):
    """(mu nu|la si) ~ sum_PQ (mu nu|P) V^-1_PQ (Q|la si): what TCMatrix_gwbse::Fill relies on."""
    T, V, G = water_inputs["ao3c"], water_inputs["aux_coulomb"], water_inputs["scf"]["eri"]
    assert np.abs(T - np.transpose(T, (0, 2, 1))).max() == 0.0 and np.abs(V - V.T).max() == 0.0
    assert np.linalg.eigvalsh(V).min() > 5e-7                 # nothing for Pseudo_InvSqrt_GWBSE to drop
    ri = np.einsum("pmn,pq,qls->mnls", T, np.linalg.inv(V), T, optimize=True)
    assert np.abs(ri - G).max() < 2e-4
    # s-type aux functions against the closed form (P|Q) = 2 pi^2.5 / (a b sqrt(a+b)) N_a N_b for same-centre s functions
    a, b = 0.25, 0.65
    fa, fb = ml.BasisFunction((0, 0, 0), (0, 0, 0), [a], [1.0]), ml.BasisFunction((0, 0, 0), (0, 0, 0), [b], [1.0])
    ref = 2 * np.pi ** 2.5 / (a * b * np.sqrt(a + b)) * fa.coefs[0] * fb.coefs[0]
    assert abs(ml.eri_two_center([fa, fb])[0, 1] - ref) < 1e-12 * ref


def test_g0w0_at_hf_and_bse_on_water(water_inputs):
    """The oracle's whole step on a real molecule: G0W0@HF (ScaHFX = 1, Vxc = 0) + BSE.  Sanity of the physics, not
    parity: screening pushes occupied levels up and the gap shrinks relative to Hartree-Fock; triplets lie below
    singlets; oscillator strengths are non-negative; the exact and plasmon-pole self-energies agree to a few mHa."""
    inp = water_inputs
    n, nocc = inp["n_basis"], inp["n_occ"]
    r = orc.gwbse_level_ranges("full", n, nocc)
    vxc = np.zeros((r["qpmax"] - r["qpmin"] + 1,) * 2)
    bseopt = orc.BSEOptions(r["homo"], r["rpamin"], r["rpamax"], r["qpmin"], r["qpmax"], r["vmin"], r["cmax"], nmax=4,
                            davidson_tolerance="lapack")
    out = {}
    for sigma in ("ppm", "exact"):
        gwopt = orc.GWOptions(r["homo"], r["qpmin"], r["qpmax"], r["rpamin"], r["rpamax"], ScaHFX=1.0,
                              sigma_integration=sigma)
        out[sigma] = orc.run_gwbse(inp["ao3c"], inp["C"], inp["energies"], vxc, inp["aux_coulomb"], gwopt, bseopt,
                                   triplets=True)
    res, eps, h = out["ppm"], inp["energies"], r["homo"]
    qp = res["qp_pert"]
    assert qp[h] > eps[h] and (qp[h + 1] - qp[h]) < (eps[h + 1] - eps[h])
    assert np.all(res["triplet_energies"][:2] < res["singlet_energies"][:2]) and res["singlet_energies"][0] > 0.1
    assert np.abs(out["exact"]["qp_pert"][1:] - qp[1:]).max() < 0.02          # valence levels: PPM vs exact
    d = orc.BSE.transition_dipoles(inp["ao_dipoles"], inp["C"], r["homo"], r["vmin"], r["cmax"], res["singlet_vectors"])
    assert np.all(orc.BSE.oscillator_strengths(res["singlet_energies"], d) >= 0.0)
