"""Generates tests/golden/*.npz: outputs of the CPU oracle (oracle/gwbse_oracle.py) on the seeded synthetic problems.

    python tests/golden/make_golden.py

PROVENANCE -- read before trusting these numbers.  The mounted reference (/root/reference/README.md:1) is a one-line
redirect stub: there is no votca/xtp source, test fixture (upstream xtp/src/tests/DataFiles/*) or binary to generate
golden vectors FROM.  These files are therefore outputs of this repo's own restatement of the GW-BSE working
equations, not of votca/xtp: parity with the reference stays UNPINNED.  What they do pin:
  * the oracle against silent drift (tests/test_golden.py, CPU, recomputes and compares),
  * the CUDA path on the GPU box against numbers that were produced in this container and travelled with the repo
    (tests/test_golden.py -m gpu), independent of whatever numpy/scipy/BLAS the box runs the oracle with,
  * the synthetic input generators (a checksum of every input array is stored).
Every array is FP64; energies in Hartree.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import gwbse_oracle as orc  # noqa: E402
from xtp_b200 import synth  # noqa: E402

CASES = ("tiny", "ch4-svp-shape")
# BASELINE.json configs[1] at its full size (benzene/def2-TZVP shape: 222 basis functions, 1110 aux functions, G0W0 PPM
# with the default 1001-point QP grid, 10 singlets + 10 triplets, Davidson "normal").  The numpy oracle needs ~7 minutes
# for it, so only the generator and the GPU test touch this case (`python tests/golden/make_golden.py --large`).
LARGE_CASES = ("benzene-tzvp-shape",)
# a real molecule: water, RHF/STO-3G orbitals and integrals from xtp_b200/molecule.py (literature-pinned SCF), G0W0@HF
# (ScaHFX = 1, Vxc = 0) with all 7 levels in every window, 4 singlets and triplets  (`--real`)
REAL_CASES = ("h2o-sto3g",)
GRID_STEPS, GRID_SPACING = 65, 0.05


def ao_dipoles(n_basis, seed=5):
    """three symmetric pseudo AO dipole matrices (stand-in for AODipole::Fill, upstream aomatrices/aodipole.cc)"""
    rng = np.random.default_rng(seed)
    r = rng.standard_normal((3, n_basis, n_basis))
    return 0.5 * (r + np.transpose(r, (0, 2, 1)))


def options(sz, nmax=4):
    gwopt = orc.GWOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax, qp_grid_steps=401)
    bseopt = orc.BSEOptions(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, nmax=nmax,
                            davidson_tolerance="lapack")
    return gwopt, bseopt


def compute(name):
    prob = synth.make_problem(name)
    sz = prob["sizes"]
    gwopt, bseopt = options(sz)
    out = {"input_checksums": np.array([np.abs(prob[k]).sum() for k in ("C", "energies", "ao3c", "aux_coulomb", "vxc")])}
    # a-1: the tensor (three slabs are enough to pin the layout and the metric rotation)
    tc = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.Fill(prob["ao3c"], prob["C"], prob["aux_coulomb"])
    out["M_slabs"] = tc.M[[0, sz.homo - sz.rpamin, sz.mtotal - 1]].copy()
    out["M_frobenius"] = np.array([np.linalg.norm(tc.M)])
    # a-2: epsilon on both axes
    rpa = orc.RPA(tc)
    rpa.configure(sz.homo, sz.rpamin, sz.rpamax)
    rpa.setRPAInputEnergies(prob["energies"][sz.rpamin:sz.rpamax + 1])
    out["eps_i_0p5"] = rpa.calculate_epsilon_i(0.5)
    out["eps_r_0p3"] = rpa.calculate_epsilon_r(0.3)
    # a-3..a-5: PPM parameters, Sigma_x, Sigma_c on a grid that crosses poles
    sig = orc.Sigma_PPM(tc, rpa)
    sig.configure(orc.SigmaOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax))
    out["sigma_x"] = sig.CalcExchangeMatrix()
    sig.PrepareScreening()           # rotates tc in place, as the reference does
    out["ppm_freq"], out["ppm_weight"] = sig.ppm.getPpm_freq().copy(), sig.ppm.getPpm_weight().copy()
    centers = prob["energies"][sz.qpmin:sz.qpmax + 1]
    offs = (np.arange(GRID_STEPS) - (GRID_STEPS - 1) / 2) * GRID_SPACING
    out["sigma_c_grid"] = np.array([[sig.CalcCorrelationDiagElement(l, centers[l] + d) for d in offs]
                                    for l in range(sz.qptotal)])
    out["sigma_c_offdiag"] = sig.CalcCorrelationOffDiag(centers)
    # a-8..a-11: the whole step in the order GWBSE::Evaluate drives it
    res = orc.run_gwbse(prob["ao3c"], prob["C"], prob["energies"], prob["vxc"], prob["aux_coulomb"], gwopt, bseopt,
                        triplets=True)
    for k in ("qp_pert", "qp_diag", "Hqp", "eps0_inv", "singlet_energies", "triplet_energies"):
        out[k] = np.asarray(res[k])
    # section 8f: full BSE, transition dipoles, oscillator strengths
    tc2 = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc2.Fill(prob["ao3c"], prob["C"], prob["aux_coulomb"])
    bse = orc.BSE(tc2)
    bse.configure(bseopt, res["rpa_energies"], res["Hqp"])
    e, X, Y = bse.Solve_singlets_BTDA()
    r = ao_dipoles(sz.n_basis)
    d = orc.BSE.transition_dipoles(r, prob["C"], sz.homo, sz.vmin, sz.cmax, X, Y)
    out["btda_singlet_energies"] = e
    out["btda_oscillator_strengths"] = orc.BSE.oscillator_strengths(e, d)
    d_tda = orc.BSE.transition_dipoles(r, prob["C"], sz.homo, sz.vmin, sz.cmax, res["singlet_vectors"])
    out["tda_oscillator_strengths"] = orc.BSE.oscillator_strengths(res["singlet_energies"], d_tda)
    return out


def compute_large(name, nmax=10):
    """whole step only, small outputs only (the N_aux^2 matrices stay out of the repository)"""
    prob = synth.make_problem(name)
    sz = prob["sizes"]
    gwopt = orc.GWOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax)
    bseopt = orc.BSEOptions(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, nmax=nmax)
    res = orc.run_gwbse(prob["ao3c"], prob["C"], prob["energies"], prob["vxc"], prob["aux_coulomb"], gwopt, bseopt,
                        triplets=True)
    out = {"input_checksums": np.array([np.abs(prob[k]).sum() for k in ("C", "energies", "ao3c", "aux_coulomb", "vxc")])}
    for k in ("qp_pert", "qp_diag", "Hqp", "eps0_inv", "singlet_energies", "triplet_energies"):
        out[k] = np.asarray(res[k])
    out["davidson_iterations"] = np.array([res["davidson_iterations"]])
    return out


def real_inputs(name):
    from xtp_b200 import molecule as ml
    assert name == "h2o-sto3g"
    return ml.gwbse_inputs(ml.water())


def compute_real(name, nmax=4):
    inp = real_inputs(name)
    r = orc.gwbse_level_ranges("full", inp["n_basis"], inp["n_occ"])
    vxc = np.zeros((r["qpmax"] - r["qpmin"] + 1,) * 2)
    gwopt = orc.GWOptions(r["homo"], r["qpmin"], r["qpmax"], r["rpamin"], r["rpamax"], ScaHFX=1.0)
    bseopt = orc.BSEOptions(r["homo"], r["rpamin"], r["rpamax"], r["qpmin"], r["qpmax"], r["vmin"], r["cmax"], nmax=nmax,
                            davidson_tolerance="lapack")
    res = orc.run_gwbse(inp["ao3c"], inp["C"], inp["energies"], vxc, inp["aux_coulomb"], gwopt, bseopt, triplets=True)
    out = {"input_checksums": np.array([np.abs(inp[k]).sum() for k in ("C", "energies", "ao3c", "aux_coulomb")]),
           "rhf_energy": np.array([inp["scf"]["energy"]])}
    for k in ("qp_pert", "qp_diag", "Hqp", "singlet_energies", "triplet_energies"):
        out[k] = np.asarray(res[k])
    d = orc.BSE.transition_dipoles(inp["ao_dipoles"], inp["C"], r["homo"], r["vmin"], r["cmax"], res["singlet_vectors"])
    out["tda_oscillator_strengths"] = orc.BSE.oscillator_strengths(res["singlet_energies"], d)
    return out


def main():
    if "--real" in sys.argv:
        for name in REAL_CASES:
            out = compute_real(name)
            path = os.path.join(HERE, f"{name}.npz")
            np.savez_compressed(path, **out)
            print(path, os.path.getsize(path), "bytes;", ", ".join(f"{k}{list(v.shape)}" for k, v in out.items()))
        return
    if "--large" in sys.argv:
        for name in LARGE_CASES:
            out = compute_large(name)
            path = os.path.join(HERE, f"{name}.npz")
            np.savez_compressed(path, **out)
            print(path, os.path.getsize(path), "bytes;", ", ".join(f"{k}{list(v.shape)}" for k, v in out.items()))
        return
    for name in CASES:
        out = compute(name)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes;", ", ".join(f"{k}{list(v.shape)}" for k, v in out.items()))


if __name__ == "__main__":
    main()
