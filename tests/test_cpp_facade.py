"""The header-only C++ facade (include/xtpb200/xtp_facade.hpp: TCMatrix_gwbse, RPA, GW, BSE, BSE_OPERATOR<...>,
DavidsonSolver with the reference's names) compiles with g++ against the C ABI without Eigen (CPU test) and, on a
GPU box, reproduces the oracle's QP and BSE singlet energies for a tiny molecule (gpu test)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "facade_test.cpp")
LIBDIR = os.path.join(ROOT, "xtp_b200")


def build(tmp_path):
    exe = str(tmp_path / "facade_test")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC, "-o", exe,
                    "-L" + LIBDIR, "-lxtpb200", "-Wl,-rpath," + LIBDIR], check=True)
    return exe


def test_facade_compiles_and_links(tmp_path):
    exe = build(tmp_path)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "facade compiled" in out


@pytest.mark.gpu
def test_facade_full_step_matches_oracle(tmp_path):
    from oracle import gwbse_oracle as orc
    from xtp_b200 import synth
    exe = build(tmp_path)
    prob = synth.make_problem("tiny")
    sz = prob["sizes"]
    nmax, grid = 3, 201
    gwopt = orc.GWOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax, qp_grid_steps=grid)
    bseopt = orc.BSEOptions(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, nmax=nmax)
    ref = orc.run_gwbse(prob["ao3c"], prob["C"], prob["energies"], prob["vxc"], prob["aux_coulomb"], gwopt, bseopt)
    head = np.array([sz.n_basis, sz.n_aux, sz.homo, sz.qpmax, sz.cmax, nmax, grid], dtype=np.float64)
    blob = np.concatenate([head, np.asfortranarray(prob["C"]).ravel(order="F"), prob["energies"],
                           np.asfortranarray(prob["vxc"]).ravel(order="F"),
                           np.asfortranarray(prob["aux_coulomb"]).ravel(order="F"), prob["ao3c"].ravel()])
    inp, outp = tmp_path / "in.bin", tmp_path / "out.txt"
    blob.tofile(inp)
    subprocess.run([exe, str(inp), str(outp)], check=True)
    vals = {}
    for line in open(outp):
        k, v = line.split()
        vals.setdefault(k, []).append(float(v))
    np.testing.assert_allclose(vals["qp"], ref["qp_pert"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(vals["singlet"], ref["singlet_energies"], rtol=0, atol=1e-4)   # Davidson tol 'normal'
    assert vals["davidson_info"] == [0.0]
    assert len(vals["btda"]) == nmax and len(vals["fosc"]) == nmax
    assert np.all(np.array(vals["btda"]) <= np.array(vals["singlet"]) + 1e-6)      # full BSE lies below TDA
    assert np.all(np.array(vals["fosc"]) >= 0.0)
    assert vals["plot_rows"] == [11.0] and len(vals["dynamic"]) == nmax
    assert np.abs(np.array(vals["dynamic"]) - np.array(vals["singlet"])).max() < 0.05
