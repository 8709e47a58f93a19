"""C-ABI surface checks that need no GPU: libxtpb200.so loads, exports every function that
include/xtpb200/xtpb200.h declares, the ctypes table covers the same set, and (on a box without a CUDA device)
the library refuses to create a context instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "xtpb200", "xtpb200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xtpb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_reference_surface():
    names = declared_functions()
    for needed in ["xtpb_tc_create", "xtpb_tc_fill_begin", "xtpb_tc_fill_block", "xtpb_tc_multiply_right_with_aux_matrix",
                   "xtpb_rpa_epsilon", "xtpb_gw_sigma_exchange", "xtpb_gw_prepare_screening",
                   "xtpb_gw_sigma_c_diag_elements", "xtpb_gw_sigma_c_offdiag", "xtpb_gw_calculate_gw_perturbation",
                   "xtpb_bse_operator_create", "xtpb_op_matmul", "xtpb_op_diagonal", "xtpb_davidson_solve"]:
        assert needed in names


def test_library_exports_every_declared_symbol():
    from xtp_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libxtpb200.so missing: run __graft_entry__.build()"
    handle = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(handle, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_ctypes_table_matches_header():
    from xtp_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_functions()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from xtp_b200 import api
    with pytest.raises(RuntimeError, match="no CUDA device"):
        api.Context(0)


def test_gaussian_quadrature_matches_numpy():
    """GaussianQuadrature is host code inside the library (no device needed): nodes/weights against numpy's rules
    through the oracle's restatement."""
    import numpy as np
    from oracle import gwbse_oracle as orc
    from xtp_b200 import api
    for scheme in ("legendre", "laguerre", "hermite"):
        for order in (8, 12, 16, 20, 40):
            a, b = api.GaussianQuadrature(scheme, order), orc.GaussianQuadrature(scheme, order)
            assert a.Order() == b.Order()
            np.testing.assert_allclose(a.points, b.points, rtol=1e-11)
            np.testing.assert_allclose(a.weights, b.weights, rtol=1e-10)


def test_host_eigh_matches_numpy():
    """The host eigensolver behind the Davidson subspace problems (tred2 + implicit QL; upstream uses
    Eigen::SelfAdjointEigenSolver there): eigenvalues, orthonormality and the decomposition itself against numpy,
    including degenerate, diagonal, tiny and badly scaled matrices and a non-trivial leading dimension."""
    import numpy as np
    from xtp_b200 import _lib
    lib = _lib.lib()
    rng = np.random.default_rng(7)
    cases = []
    for n in (1, 2, 3, 7, 30, 111, 257):
        A = rng.standard_normal((n, n))
        cases.append(A + A.T)
    cases.append(np.diag(np.arange(6.0)))                                   # already diagonal
    Q = np.linalg.qr(rng.standard_normal((12, 12)))[0]
    cases.append((Q * np.array([1.0] * 5 + [2.0] * 4 + [3.0] * 3)) @ Q.T)   # degenerate clusters
    cases.append(1e-12 * cases[3])
    cases.append(1e+9 * cases[4])
    B = rng.standard_normal((40, 40)); B = B + B.T; B[20:, :20] = 0; B[:20, 20:] = 0   # block diagonal (e == 0 splits)
    cases.append(B)
    for A in cases:
        n = A.shape[0]
        lda = n + 3
        buf = np.zeros((lda, n), order="F")
        buf[:n, :] = np.tril(A)                  # only the lower triangle is read
        w = np.empty(n)
        _lib.check(lib.xtpb_host_eigh(n, buf.ctypes.data_as(_lib.dptr), lda, w.ctypes.data_as(_lib.dptr)))
        U = buf[:n, :]
        scale = max(1.0, np.abs(A).max())
        ref = np.linalg.eigvalsh(A)
        np.testing.assert_allclose(w, ref, rtol=0, atol=1e-13 * scale * n)
        np.testing.assert_allclose(U.T @ U, np.eye(n), atol=1e-13 * n)
        np.testing.assert_allclose(U @ np.diag(w) @ U.T, A, rtol=0, atol=1e-13 * scale * n)
        assert np.all(np.diff(w) >= 0)


def test_anderson_mixing_matches_oracle():
    """Anderson mixer of the evGW loop (host code): every order / history length against the oracle's least-squares
    restatement, including a rank-deficient history (two identical residual differences) and the order-1 linear mix."""
    import numpy as np
    from oracle import gwbse_oracle as orc
    from xtp_b200 import _lib
    lib = _lib.lib()
    rng = np.random.default_rng(3)
    n = 17
    for order in (1, 2, 3, 6):
        for nh in (1, 2, 3, 5, 8):
            ins = rng.standard_normal((nh, n))
            outs = ins + 0.1 * rng.standard_normal((nh, n))
            if nh >= 3:
                outs[1] = ins[1] + (outs[0] - ins[0])            # duplicate residual: singular normal equations
            ref = orc.Anderson(order, 0.6)
            for a, b in zip(ins, outs):
                ref.UpdateInput(a)
                ref.UpdateOutput(b)
            want = ref.MixHistory()
            got = np.empty(n)
            _lib.check(lib.xtpb_anderson_mix(order, 0.6, n, nh, ins.ctypes.data_as(_lib.dptr),
                                             outs.ctypes.data_as(_lib.dptr), got.ctypes.data_as(_lib.dptr)))
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)
            if order == 1:
                np.testing.assert_allclose(got, 0.6 * outs[-1] + 0.4 * ins[-1], rtol=0, atol=1e-14)
    # a linear fixed-point map x -> A x + b is solved exactly once the history spans the space
    A = 0.5 * np.linalg.qr(rng.standard_normal((4, 4)))[0]
    b = rng.standard_normal(4)
    fixed = np.linalg.solve(np.eye(4) - A, b)
    mix = orc.Anderson(6, 1.0)
    x = np.zeros(4)
    for _ in range(6):
        mix.UpdateInput(x)
        mix.UpdateOutput(A @ x + b)
        x = mix.MixHistory()
    np.testing.assert_allclose(x, fixed, atol=1e-9)
