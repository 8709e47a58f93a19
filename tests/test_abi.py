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
