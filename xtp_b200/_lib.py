"""ctypes binding of libxtpb200.so (C ABI: include/xtpb200/xtpb200.h).

The library is the product; there is no Python/CPU fallback.  ``lib()`` raises
``RuntimeError`` when the shared object is missing or cannot be loaded."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxtpb200.so")

idx = C.c_longlong
dptr = C.POINTER(C.c_double)
iptr = C.POINTER(idx)
vp = C.c_void_p


class GwOptions(C.Structure):
    _fields_ = [
        ("homo", idx), ("qpmin", idx), ("qpmax", idx), ("rpamin", idx), ("rpamax", idx),
        ("eta", C.c_double), ("g_sc_limit", C.c_double), ("g_sc_max_iterations", idx),
        ("gw_sc_limit", C.c_double), ("gw_sc_max_iterations", idx), ("shift", C.c_double),
        ("ScaHFX", C.c_double), ("sigma_integration", C.c_int), ("reset_3c", idx),
        ("qp_solver", C.c_int), ("qp_grid_steps", idx), ("qp_grid_spacing", C.c_double),
        ("gw_mixing_order", idx), ("gw_mixing_alpha", C.c_double), ("quadrature_scheme", C.c_int),
        ("order", idx), ("alpha", C.c_double),
    ]


class BseOptions(C.Structure):
    _fields_ = [
        ("homo", idx), ("rpamin", idx), ("rpamax", idx), ("qpmin", idx), ("qpmax", idx),
        ("vmin", idx), ("cmax", idx), ("nmax", idx), ("use_Hqp_offdiag", C.c_int),
    ]


class DavidsonOptions(C.Structure):
    _fields_ = [
        ("tolerance", C.c_double), ("correction", C.c_int), ("size_update", C.c_int),
        ("iter_max", idx), ("max_search_space", idx), ("size_initial_guess", idx),
    ]


class ContractDesc(C.Structure):
    _fields_ = [(n, idx) for n in (
        "M", "N", "K", "n_outer", "n_batch",
        "a_row", "a_k", "a_outer", "a_batch", "a_len",
        "b_row", "b_k", "b_outer", "b_batch", "b_len",
        "c_row", "c_col", "c_batch", "c_len", "c_col_inner", "c_col_outer",
        "d_outer", "d_batch", "d_len")] + [
        ("alpha", C.c_double), ("beta", C.c_double), ("lower", C.c_int), ("force_cfg", C.c_int),
        ("force_splits", C.c_int)]


# name -> (restype, argtypes); every symbol declared in include/xtpb200/xtpb200.h
class GwbseRangeOptions(C.Structure):
    _fields_ = [("mode", C.c_int), ("n_levels", idx), ("n_occ", idx), ("n_core_ignored", idx), ("rpamax", C.c_double),
                ("qpmin", C.c_double), ("qpmax", C.c_double), ("bsemin", C.c_double), ("bsemax", C.c_double)]


class GwbseRanges(C.Structure):
    _fields_ = [(k, idx) for k in ("homo", "rpamin", "rpamax", "qpmin", "qpmax", "vmin", "cmax", "qptotal", "rpatotal",
                                   "bse_vtotal", "bse_ctotal", "bse_size")]


PROTOTYPES = {
    "xtpb_gwbse_level_ranges": (C.c_int, [C.POINTER(GwbseRangeOptions), C.POINTER(GwbseRanges)]),
    "xtpb_last_error": (C.c_char_p, []),
    "xtpb_version": (C.c_int, []),
    "xtpb_launch_count": (C.c_longlong, []),
    "xtpb_tma_launch_count": (C.c_longlong, []),
    "xtpb_tma_single_box_launch_count": (C.c_longlong, []),
    "xtpb_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "xtpb_ctx_destroy": (C.c_int, [vp]),
    "xtpb_ctx_sync": (C.c_int, [vp]),
    "xtpb_ctx_solver_seconds": (C.c_int, [vp, dptr, C.c_int]),
    "xtpb_comm_unique_id": (C.c_int, [C.c_char_p]),
    "xtpb_ctx_comm_init": (C.c_int, [vp, C.c_char_p, C.c_int, C.c_int]),
    "xtpb_ctx_comm_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "xtpb_alloc_stats": (C.c_int, [dptr, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), dptr, C.c_int]),
    "xtpb_alloc_free_wait_seconds": (C.c_int, [dptr]),
    "xtpb_host_alloc": (C.c_int, [C.c_ulonglong, C.POINTER(vp)]),
    "xtpb_host_free": (C.c_int, [vp]),
    "xtpb_profile_enable": (C.c_int, [C.c_int]),
    "xtpb_profile_reset": (C.c_int, []),
    "xtpb_profile_get": (C.c_int, [C.c_int, dptr, dptr, iptr]),
    "xtpb_tc_create": (C.c_int, [vp, idx, idx, idx, idx, idx, C.POINTER(vp)]),
    "xtpb_tc_destroy": (C.c_int, [vp]),
    "xtpb_tc_sizes": (C.c_int, [vp, iptr, iptr, iptr]),
    "xtpb_tc_set_raw": (C.c_int, [vp, dptr]),
    "xtpb_tc_set_raw_dev": (C.c_int, [vp, vp]),
    "xtpb_tc_device_view": (C.c_int, [vp, C.POINTER(vp), iptr, iptr, iptr]),
    "xtpb_tc_get_slab": (C.c_int, [vp, idx, dptr]),
    "xtpb_tc_fill_begin": (C.c_int, [vp, idx, dptr, idx]),
    "xtpb_tc_fill_block": (C.c_int, [vp, idx, idx, dptr, idx]),
    "xtpb_tc_fill_block_dev": (C.c_int, [vp, idx, idx, vp, idx]),
    "xtpb_tc_fill_block_packed": (C.c_int, [vp, idx, idx, vp]),
    "xtpb_tc_fill_block_packed_dev": (C.c_int, [vp, idx, idx, vp]),
    "xtpb_tc_local_aux_range": (C.c_int, [vp, iptr, iptr]),
    "xtpb_tc_fill_sharded_packed": (C.c_int, [vp, vp, C.c_int]),
    "xtpb_tc_multiply_right_with_aux_matrix": (C.c_int, [vp, dptr, idx]),
    "xtpb_tc_apply_coulomb_metric": (C.c_int, [vp, dptr, idx, dptr, idx, C.c_double, iptr]),
    "xtpb_tc_coulomb_metric_begin": (C.c_int, [vp, dptr, idx, dptr, idx]),
    "xtpb_tc_metric_path_info": (C.c_int, [vp, iptr, iptr]),
    "xtpb_tc_ppm_prefetch_begin": (C.c_int, [vp, dptr, idx, C.c_double]),
    "xtpb_tc_ppm_prefetch_info": (C.c_int, [vp, C.POINTER(C.c_int), iptr, iptr]),
    "xtpb_rpa_epsilon": (C.c_int, [vp, dptr, idx, idx, idx, C.c_double, dptr, C.c_int, C.c_int, dptr]),
    "xtpb_gw_options_default": (None, [C.POINTER(GwOptions)]),
    "xtpb_gaussian_quadrature": (C.c_int, [C.c_int, idx, dptr, dptr, iptr]),
    "xtpb_gw_create": (C.c_int, [vp, vp, C.POINTER(GwOptions), dptr, idx, dptr, idx, C.POINTER(vp)]),
    "xtpb_gw_destroy": (C.c_int, [vp]),
    "xtpb_gw_sigma_exchange": (C.c_int, [vp, dptr]),
    "xtpb_gw_set_rpa_input_energies": (C.c_int, [vp, dptr]),
    "xtpb_gw_get_rpa_input_energies": (C.c_int, [vp, dptr]),
    "xtpb_gw_prepare_screening": (C.c_int, [vp]),
    "xtpb_gw_get_ppm": (C.c_int, [vp, dptr, dptr]),
    "xtpb_gw_sigma_c_diag_elements": (C.c_int, [vp, idx, iptr, dptr, dptr, dptr]),
    "xtpb_gw_sigma_c_diag": (C.c_int, [vp, dptr, dptr]),
    "xtpb_gw_sigma_c_grid": (C.c_int, [vp, dptr, dptr]),
    "xtpb_gw_plot_sigma": (C.c_int, [vp, idx, C.c_double, idx, iptr, dptr]),
    "xtpb_gw_sigma_c_offdiag": (C.c_int, [vp, dptr, dptr]),
    "xtpb_gw_grid_scan_info": (C.c_int, [vp, C.POINTER(C.c_int), iptr, dptr, dptr]),
    "xtpb_gw_point_eval_info": (C.c_int, [vp, iptr, iptr]),
    "xtpb_ppm_grid_chunk": (C.c_int, []),
    "xtpb_ppm_grid_plan": (C.c_int, [idx, dptr, C.c_double, idx, C.c_double, C.c_double, idx, dptr, iptr,
                                     C.POINTER(C.c_int), iptr, C.POINTER(C.c_int)]),
    "xtpb_gw_calculate_gw_perturbation": (C.c_int, [vp]),
    "xtpb_gw_calculate_hqp": (C.c_int, [vp]),
    "xtpb_gw_get_gwa_results": (C.c_int, [vp, dptr]),
    "xtpb_gw_get_hqp": (C.c_int, [vp, dptr]),
    "xtpb_gw_diagonalize_qp_hamiltonian": (C.c_int, [vp, dptr, dptr]),
    "xtpb_gw_unconverged_levels": (C.c_int, [vp, iptr]),
    "xtpb_bse_create": (C.c_int, [vp, vp, C.POINTER(BseOptions), dptr, dptr, idx, C.c_int, C.POINTER(vp)]),
    "xtpb_bse_destroy": (C.c_int, [vp]),
    "xtpb_bse_screening_info": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "xtpb_bse_get_epsilon_0_inv": (C.c_int, [vp, dptr]),
    "xtpb_bse_operator_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "xtpb_bse_operator_create_raw": (C.c_int, [vp, vp, idx, idx, idx, idx, dptr, dptr, idx, C.c_int, C.c_int,
                                               C.c_int, C.c_int, C.POINTER(vp)]),
    "xtpb_bse_solve_btda": (C.c_int, [vp, C.c_int, C.POINTER(DavidsonOptions), dptr, dptr, dptr, idx, C.POINTER(C.c_int), iptr]),
    "xtpb_bse_perturbative_dynamical_screening": (C.c_int, [vp, idx, dptr, dptr, dptr, idx, idx, C.c_double, dptr,
                                                            iptr]),
    "xtpb_bse_transition_dipoles": (C.c_int, [vp, idx, dptr, idx, dptr, idx, dptr, dptr, idx, dptr]),
    "xtpb_oscillator_strengths": (C.c_int, [idx, dptr, dptr, dptr]),
    "xtpb_dense_operator_create": (C.c_int, [vp, dptr, idx, idx, C.POINTER(vp)]),
    "xtpb_op_destroy": (C.c_int, [vp]),
    "xtpb_op_size": (C.c_int, [vp, iptr]),
    "xtpb_op_matmul": (C.c_int, [vp, dptr, idx, idx, dptr, idx]),
    "xtpb_op_diagonal": (C.c_int, [vp, dptr]),
    "xtpb_op_get_full_matrix": (C.c_int, [vp, dptr, idx]),
    "xtpb_davidson_options_default": (None, [C.POINTER(DavidsonOptions)]),
    "xtpb_davidson_solve": (C.c_int, [vp, idx, C.POINTER(DavidsonOptions), dptr, dptr, idx, C.POINTER(C.c_int), iptr]),
    "xtpb_anderson_mix": (C.c_int, [idx, C.c_double, idx, idx, dptr, dptr, dptr]),
    "xtpb_host_eigh": (C.c_int, [idx, dptr, idx, dptr]),
    "xtpb_contract_host": (C.c_int, [vp, C.POINTER(ContractDesc), dptr, dptr, dptr, dptr]),
    "xtpb_contract_bench": (C.c_int, [vp, C.POINTER(ContractDesc), C.c_int, dptr]),
    "xtpb_contract_plan": (C.c_int, [C.POINTER(ContractDesc), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
}

_lib = None


def lib():
    """Load libxtpb200.so; raise loudly if it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make -C xtp_b200/csrc` (or __graft_entry__.build()). "
            "xtp_b200 has no CPU fallback.")
    handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(handle, name)       # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


class XtpbError(RuntimeError):
    pass


def check(status):
    if status != 0:
        raise XtpbError(lib().xtpb_last_error().decode("utf-8", "replace"))
