"""Host-side mirror of the reference's GW-BSE interface over libxtpb200's C ABI.

Class and method names follow upstream votca/xtp (TCMatrix_gwbse, RPA, GW,
Sigma via GW, BSE, BSE_OPERATOR, DavidsonSolver) so parity tests read like the
reference's own tests.  All arrays are numpy float64; matrices are passed to the
library column-major (Eigen's default).  Every call goes to CUDA; a missing or
unloadable library raises (see ``_lib.lib``)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, dptr, idx, iptr, vp


def _d(a):
    return a.ctypes.data_as(dptr)


def _f(a):
    """Column-major float64 copy/view."""
    return np.asfortranarray(a, dtype=np.float64)


class Context:
    """One CUDA device + stream (replaces OpenMP_CUDA / CudaPipeline)."""

    def __init__(self, device=0):
        self._h = vp()
        check(_lib.lib().xtpb_ctx_create(int(device), C.byref(self._h)))

    def close(self):
        if self._h:
            _lib.lib().xtpb_ctx_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(_lib.lib().xtpb_ctx_sync(self._h))

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """Join the NCCL communicator (one process per GPU); must precede TCMatrix_gwbse.Initialize."""
        assert len(unique_id) == 128
        check(_lib.lib().xtpb_ctx_comm_init(self._h, C.c_char_p(unique_id), int(rank), int(world)))

    def comm_info(self):
        r, w = C.c_int(0), C.c_int(1)
        check(_lib.lib().xtpb_ctx_comm_info(self._h, C.byref(r), C.byref(w)))
        return r.value, w.value

    def solver_seconds(self, reset=False):
        s = C.c_double()
        check(_lib.lib().xtpb_ctx_solver_seconds(self._h, C.byref(s), int(reset)))
        return s.value


def launch_count():
    return int(_lib.lib().xtpb_launch_count())


def tma_launch_count():
    return int(_lib.lib().xtpb_tma_launch_count())


def tma_single_box_launch_count():
    return int(_lib.lib().xtpb_tma_single_box_launch_count())


def comm_unique_id() -> bytes:
    """128-byte NCCL unique id (call on rank 0, distribute to every rank, then Context.comm_init)."""
    buf = C.create_string_buffer(128)
    check(_lib.lib().xtpb_comm_unique_id(buf))
    return buf.raw


class PinnedBuffer:
    """float64 numpy view of page-locked host memory (xtpb_host_alloc); keeps the allocation alive."""

    def __init__(self, count):
        self._p = vp()
        self.count = int(count)
        check(_lib.lib().xtpb_host_alloc(C.c_ulonglong(max(8, self.count * 8)), C.byref(self._p)))
        self.array = np.ctypeslib.as_array(C.cast(self._p, dptr), shape=(max(1, self.count),))[:self.count]

    def close(self):
        if self._p:
            self.array = None
            _lib.lib().xtpb_host_free(self._p)
            self._p = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


PROFILE_TAGS = {"other": 0, "fill": 1, "rotate": 2, "epsilon": 3, "sigma_x": 4, "sigma_offdiag": 5, "bse_matmul": 6,
                "davidson": 7, "dense_aux": 8, "sigma_ppm_grid": 9, "sigma_ppm_pairs": 10, "solver": 11,
                "unpack": 12, "cda": 13, "exact": 14, "comm": 15, "sigma_ppm_points": 16}
CONTRACTION_TAGS = ("other", "fill", "rotate", "epsilon", "sigma_x", "sigma_offdiag", "bse_matmul", "davidson",
                    "dense_aux", "cda", "exact")


def profile_enable(on=True):
    check(_lib.lib().xtpb_profile_enable(int(on)))


def profile_reset():
    check(_lib.lib().xtpb_profile_reset())


def profile_summary():
    """{tag: {"ms", "work", "launches"}} for every tag with at least one launch (synchronises the device)."""
    out = {}
    for name, tag in PROFILE_TAGS.items():
        ms, work, n = C.c_double(), C.c_double(), idx(0)
        check(_lib.lib().xtpb_profile_get(tag, C.byref(ms), C.byref(work), C.byref(n)))
        if n.value:
            out[name] = {"ms": ms.value, "work": work.value, "launches": int(n.value)}
    return out


def pack_lower(ao3c):
    """ao3c[P, mu, nu] symmetric -> packed lower triangles [P, n(n+1)/2] (row mu holds nu = 0..mu)."""
    n = ao3c.shape[-1]
    il = np.tril_indices(n)
    return np.ascontiguousarray(ao3c[:, il[0], il[1]])


class TCMatrix_gwbse:
    """upstream xtp/src/libxtp/threecenter_gwbse.cc"""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self._h = vp()

    def Initialize(self, auxsize, mmin, mmax, nmin, nmax):
        self.close()
        check(_lib.lib().xtpb_tc_create(self.ctx._h, auxsize, mmin, mmax, nmin, nmax, C.byref(self._h)))
        self._aux, self.mmin, self.mmax, self.nmin, self.nmax = int(auxsize), int(mmin), int(mmax), int(nmin), int(nmax)
        return self

    def close(self):
        if self._h:
            _lib.lib().xtpb_tc_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def auxsize(self): return self._aux
    def msize(self): return self.mmax - self.mmin + 1
    def nsize(self): return self.nmax - self.nmin + 1
    def get_mmin(self): return self.mmin
    def get_mmax(self): return self.mmax
    def get_nmin(self): return self.nmin
    def get_nmax(self): return self.nmax

    def set_raw(self, M):
        """M[m, P, n] C-contiguous == the reference's vector of column-major (n x aux) slabs."""
        M = np.ascontiguousarray(M, dtype=np.float64)
        assert M.shape == (self.msize(), self._aux, self.nsize())
        check(_lib.lib().xtpb_tc_set_raw(self._h, _d(M)))

    def set_raw_dev(self, dev_ptr):
        """Tensor from a contiguous device buffer [m][P][n] (e.g. a torch tensor's data_ptr())."""
        check(_lib.lib().xtpb_tc_set_raw_dev(self._h, vp(int(dev_ptr))))

    def device_view(self):
        """(pointer, ld_n, slab_stride, n_local) of the resident tensor: M[m](n, P) at ptr + 8*(m*slab + P*ld_n + n)."""
        p, ld, sl, nl = vp(), idx(0), idx(0), idx(0)
        check(_lib.lib().xtpb_tc_device_view(self._h, C.byref(p), C.byref(ld), C.byref(sl), C.byref(nl)))
        return int(p.value), int(ld.value), int(sl.value), int(nl.value)

    def torch_view(self):
        """The resident tensor as a torch tensor [m, P, ld_n] sharing the library's memory (read it, do not write it);
        columns >= n_local of the last axis are padding."""
        import torch
        ptr, ld, sl, nl = self.device_view()
        mt = self.mmax - self.mmin + 1

        class _Holder:
            __cuda_array_interface__ = {"shape": (mt, self._aux, ld), "typestr": "<f8", "data": (ptr, False),
                                        "version": 3, "strides": (sl * 8, ld * 8, 8)}
        return torch.as_tensor(_Holder(), device="cuda")[:, :, :nl]

    def __getitem__(self, m):
        """operator[]: slab m as an (nsize x auxsize) matrix."""
        out = np.empty((self._aux, self.nsize()))
        check(_lib.lib().xtpb_tc_get_slab(self._h, int(m), _d(out)))
        return out.T

    def get_raw(self):
        return np.stack([self[m].T for m in range(self.msize())])

    def Fill3cMO(self, ao3c, C_mo, block=64):
        """ao3c[P, mu, nu] symmetric slices (host), C_mo (n_basis x n_levels)."""
        Cf = _f(C_mo)
        nb = Cf.shape[0]
        check(_lib.lib().xtpb_tc_fill_begin(self._h, nb, _d(Cf), Cf.shape[0]))
        ao3c = np.ascontiguousarray(ao3c, dtype=np.float64)
        for p0 in range(0, self._aux, block):
            cnt = min(block, self._aux - p0)
            blk = ao3c[p0:p0 + cnt]
            check(_lib.lib().xtpb_tc_fill_block(self._h, p0, cnt, _d(blk), nb))

    def fill_begin(self, C_mo):
        Cf = _f(C_mo)
        check(_lib.lib().xtpb_tc_fill_begin(self._h, Cf.shape[0], _d(Cf), Cf.shape[0]))

    def fill_block_dev(self, P0, nP, dev_ptr, ld):
        check(_lib.lib().xtpb_tc_fill_block_dev(self._h, int(P0), int(nP), vp(int(dev_ptr)), int(ld)))

    def fill_block_packed(self, P0, packed):
        """packed[P, n(n+1)/2] host array (ideally a PinnedBuffer view)."""
        assert packed.dtype == np.float64 and packed.flags.c_contiguous
        check(_lib.lib().xtpb_tc_fill_block_packed(self._h, int(P0), packed.shape[0], vp(packed.ctypes.data)))

    def fill_block_packed_dev(self, P0, nP, dev_ptr):
        check(_lib.lib().xtpb_tc_fill_block_packed_dev(self._h, int(P0), int(nP), vp(int(dev_ptr))))

    def local_aux_range(self):
        """(P0, nP): the aux functions whose AO slices this rank contributes to the collective fill."""
        p0, n = idx(0), idx(0)
        check(_lib.lib().xtpb_tc_local_aux_range(self._h, C.byref(p0), C.byref(n)))
        return int(p0.value), int(n.value)

    def fill_sharded_packed(self, packed_local=None, dev_ptr=None):
        """Collective Fill3cMO: packed lower-triangular slices of this rank's aux range (host array or device pointer)."""
        if dev_ptr is not None:
            check(_lib.lib().xtpb_tc_fill_sharded_packed(self._h, vp(int(dev_ptr)), 1))
        else:
            assert packed_local.dtype == np.float64 and packed_local.flags.c_contiguous
            check(_lib.lib().xtpb_tc_fill_sharded_packed(self._h, vp(packed_local.ctypes.data), 0))

    def fill_block(self, P0, blk, ld=None):
        blk = np.ascontiguousarray(blk, dtype=np.float64)
        check(_lib.lib().xtpb_tc_fill_block(self._h, int(P0), blk.shape[0], _d(blk), int(ld or blk.shape[-1])))

    def MultiplyRightWithAuxMatrix(self, A):
        Af = _f(A)
        check(_lib.lib().xtpb_tc_multiply_right_with_aux_matrix(self._h, _d(Af), Af.shape[0]))

    def Fill(self, ao3c, C_mo, aux_coulomb, aux_overlap=None, etol=5e-7):
        self.coulomb_metric_begin(aux_coulomb, aux_overlap)    # eigensolver of the metric runs underneath Fill3cMO
        self.Fill3cMO(ao3c, C_mo)
        self.removedfunctions = self.apply_coulomb_metric(aux_coulomb, aux_overlap, etol)

    def ppm_prefetch_begin(self, rpa_energies, homo, eta=1e-3):
        """Hint (after fill_begin, before the fill blocks): accumulate the plasmon-pole model's two epsilon matrices
        for these RPA input energies while the aux blocks arrive (xtpb_tc_ppm_prefetch_begin)."""
        e = np.ascontiguousarray(rpa_energies, dtype=np.float64)
        check(_lib.lib().xtpb_tc_ppm_prefetch_begin(self._h, _d(e), int(homo), float(eta)))

    def ppm_prefetch_info(self):
        done, rows, used = C.c_int(0), idx(0), idx(0)
        check(_lib.lib().xtpb_tc_ppm_prefetch_info(self._h, C.byref(done), C.byref(rows), C.byref(used)))
        return {"complete": bool(done.value), "aux_functions_done": int(rows.value), "matrices_used": int(used.value)}

    def metric_path_info(self):
        """How apply_coulomb_metric obtained its factor so far (xtpb_tc_metric_path_info)."""
        ch, ei = idx(0), idx(0)
        check(_lib.lib().xtpb_tc_metric_path_info(self._h, C.byref(ch), C.byref(ei)))
        return {"cholesky_calls": int(ch.value), "eigensolver_calls": int(ei.value)}

    def coulomb_metric_begin(self, V, S=None):
        """Announce the metric matrices before the fill (xtpb_tc_coulomb_metric_begin: on the eigensolver path the first
        decomposition starts underneath the fill); the matching apply_coulomb_metric(V, S) call consumes the hint.  The
        column-major copies are kept alive here."""
        Vf = _f(V)
        Sf = _f(S) if S is not None else None
        self._metric_inputs = (V, S, Vf, Sf)
        check(_lib.lib().xtpb_tc_coulomb_metric_begin(self._h, _d(Vf), Vf.shape[0], _d(Sf) if Sf is not None else None,
                                                      Sf.shape[0] if Sf is not None else 0))

    def apply_coulomb_metric(self, V, S=None, etol=5e-7):
        pre = getattr(self, "_metric_inputs", None)
        self._metric_inputs = None
        if pre is not None and pre[0] is V and pre[1] is S:
            Vf, Sf = pre[2], pre[3]                # the very buffers the helper thread was given
        else:
            if pre is not None:
                raise ValueError("coulomb_metric_begin was called with different matrices")
            Vf = _f(V)
            Sf = _f(S) if S is not None else None
        removed = idx(0)
        check(_lib.lib().xtpb_tc_apply_coulomb_metric(
            self._h, _d(Vf), Vf.shape[0], _d(Sf) if Sf is not None else None,
            Sf.shape[0] if Sf is not None else 0, float(etol), C.byref(removed)))
        return int(removed.value)


class RPA:
    """upstream xtp/src/libxtp/gwbse/rpa.cc"""

    def __init__(self, Mmn: TCMatrix_gwbse):
        self.Mmn = Mmn
        self.eta = 1e-3

    def configure(self, homo, rpamin, rpamax):
        self.homo, self.rpamin, self.rpamax = int(homo), int(rpamin), int(rpamax)

    def setRPAInputEnergies(self, e):
        self.energies = np.ascontiguousarray(e, dtype=np.float64)

    def getRPAInputEnergies(self):
        return self.energies

    def _eps(self, omegas, imag):
        om = np.ascontiguousarray(np.atleast_1d(omegas), dtype=np.float64)
        na = self.Mmn.auxsize()
        out = np.empty((len(om), na, na))
        check(_lib.lib().xtpb_rpa_epsilon(self.Mmn._h, _d(self.energies), self.homo, self.rpamin, self.rpamax,
                                          float(self.eta), _d(om), len(om), int(imag), _d(out)))
        return out

    def calculate_epsilon_i(self, omega):
        return self._eps(omega, True)[0]

    def calculate_epsilon_r(self, omega):
        return self._eps(omega, False)[0]

    def calculate_epsilon_batch(self, omegas, imag=True):
        return self._eps(omegas, imag)


SIGMA = {"ppm": 0, "exact": 1, "cda": 2}
QPSOLVER = {"grid": 0, "fixedpoint": 1}
QUAD = {"legendre": 0, "laguerre": 1, "hermite": 2}


class GaussianQuadrature:
    """upstream xtp/src/libxtp/gwbse/gaussian_quadrature.cc (host code inside libxtpb200)"""

    def __init__(self, scheme="legendre", order=12):
        pts, wts, n = np.empty(2 * order + 2), np.empty(2 * order + 2), idx(0)
        check(_lib.lib().xtpb_gaussian_quadrature(QUAD[scheme] if isinstance(scheme, str) else int(scheme), int(order),
                                                  _d(pts), _d(wts), C.byref(n)))
        self.points, self.weights = pts[:n.value].copy(), wts[:n.value].copy()

    def Order(self): return len(self.points)
    def ScaledPoint(self, j): return self.points[j]
    def ScaledWeight(self, j): return self.weights[j]


def gw_options(**kw):
    o = _lib.GwOptions()
    _lib.lib().xtpb_gw_options_default(C.byref(o))
    for k, v in kw.items():
        if k == "sigma_integration" and isinstance(v, str):
            v = SIGMA[v]
        if k == "qp_solver" and isinstance(v, str):
            v = QPSOLVER[v]
        if k == "quadrature_scheme" and isinstance(v, str):
            v = QUAD[v]
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


class GW:
    """upstream xtp/src/libxtp/gwbse/gw.cc (+ sigma_base / sigma_ppm / ppm)"""

    def __init__(self, ctx: Context, Mmn: TCMatrix_gwbse, vxc, dft_energies):
        self.ctx, self.Mmn = ctx, Mmn
        self.vxc = _f(vxc)
        self.dft_energies = np.ascontiguousarray(dft_energies, dtype=np.float64)
        self._h = vp()

    def configure(self, opt):
        self.close()
        self.opt = opt
        self.qptotal = int(opt.qpmax - opt.qpmin + 1)
        self.rpatotal = int(opt.rpamax - opt.rpamin + 1)
        check(_lib.lib().xtpb_gw_create(self.ctx._h, self.Mmn._h, C.byref(opt), _d(self.vxc), self.vxc.shape[0],
                                        _d(self.dft_energies), len(self.dft_energies), C.byref(self._h)))

    def close(self):
        if self._h:
            _lib.lib().xtpb_gw_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def CalcExchangeMatrix(self):
        out = np.empty((self.qptotal, self.qptotal), order="F")
        check(_lib.lib().xtpb_gw_sigma_exchange(self._h, _d(out)))
        return out

    def setRPAInputEnergies(self, e):
        e = np.ascontiguousarray(e, dtype=np.float64)
        assert len(e) == self.rpatotal
        check(_lib.lib().xtpb_gw_set_rpa_input_energies(self._h, _d(e)))

    def RPAInputEnergies(self):
        out = np.empty(self.rpatotal)
        check(_lib.lib().xtpb_gw_get_rpa_input_energies(self._h, _d(out)))
        return out

    def PrepareScreening(self):
        check(_lib.lib().xtpb_gw_prepare_screening(self._h))

    def getPpm(self):
        na = self.Mmn.auxsize()
        w, f = np.empty(na), np.empty(na)
        check(_lib.lib().xtpb_gw_get_ppm(self._h, _d(w), _d(f)))
        return w, f

    def CalcCorrelationDiagElements(self, levels, frequencies, derivative=False):
        lv = np.ascontiguousarray(levels, dtype=np.int64)
        fr = np.ascontiguousarray(frequencies, dtype=np.float64)
        val = np.empty(len(lv))
        der = np.empty(len(lv)) if derivative else None
        check(_lib.lib().xtpb_gw_sigma_c_diag_elements(self._h, len(lv), lv.ctypes.data_as(iptr), _d(fr), _d(val),
                                                       _d(der) if derivative else None))
        return (val, der) if derivative else val

    def CalcCorrelationDiagElement(self, level, frequency):
        return float(self.CalcCorrelationDiagElements([level], [frequency])[0])

    def CalcCorrelationDiagElementDerivative(self, level, frequency):
        return float(self.CalcCorrelationDiagElements([level], [frequency], True)[1][0])

    def CalcCorrelationDiag(self, frequencies):
        fr = np.ascontiguousarray(frequencies, dtype=np.float64)
        out = np.empty(self.qptotal)
        check(_lib.lib().xtpb_gw_sigma_c_diag(self._h, _d(fr), _d(out)))
        return out

    def CalcCorrelationGrid(self, center_frequencies):
        """Sigma_c of every gw level on its QP grid (GW::SolveQP_Grid's scan / GW::PlotSigma): (qptotal, steps)."""
        fr = np.ascontiguousarray(center_frequencies, dtype=np.float64)
        assert len(fr) == self.qptotal
        out = np.empty((self.qptotal, int(self.opt.qp_grid_steps)))
        check(_lib.lib().xtpb_gw_sigma_c_grid(self._h, _d(fr), _d(out)))
        return out

    def grid_scan_info(self):
        """How the last QP-grid scan ran on this rank: compressed (far poles through Chebyshev moments) or pole by pole,
        the bins of the plan, and evaluated / equivalent (pole, frequency) pairs (xtpb_gw_grid_scan_info)."""
        comp, nb, direct, equiv = C.c_int(0), idx(0), C.c_double(0.0), C.c_double(0.0)
        check(_lib.lib().xtpb_gw_grid_scan_info(self._h, C.byref(comp), C.byref(nb), C.byref(direct), C.byref(equiv)))
        return {"compressed": bool(comp.value), "bins": nb.value, "direct_evaluations": direct.value,
                "equivalent_evaluations": equiv.value}

    def point_eval_info(self):
        """Calls that evaluated single (level, frequency) Sigma_c values through the moments of the last compressed
        grid scan / by streaming the slabs (xtpb_gw_point_eval_info)."""
        comp, direct = idx(0), idx(0)
        check(_lib.lib().xtpb_gw_point_eval_info(self._h, C.byref(comp), C.byref(direct)))
        return {"compressed_calls": int(comp.value), "direct_calls": int(direct.value)}

    def PlotSigma(self, steps, spacing, states, filename=None):
        """GW::PlotSigma: (steps, 2*len(states)) table, columns (frequency, Sigma_c + e_KS + Sigma_x - Vxc) per state;
        written to `filename` in upstream's text layout when one is given."""
        st = np.ascontiguousarray(states, dtype=np.int64)
        tab = np.empty((int(steps), 2 * len(st)), order="F")
        check(_lib.lib().xtpb_gw_plot_sigma(self._h, int(steps), float(spacing), len(st), st.ctypes.data_as(iptr), _d(tab)))
        if filename is not None:
            qpmin = int(self.opt.qpmin)
            head = "\t".join(f"#frequency_{qpmin + int(l)}\tSigma_c_{qpmin + int(l)}" for l in st)
            np.savetxt(filename, tab, delimiter="\t", header=head, comments="")
        return tab

    def CalcCorrelationOffDiag(self, frequencies):
        fr = np.ascontiguousarray(frequencies, dtype=np.float64)
        out = np.empty((self.qptotal, self.qptotal), order="F")
        check(_lib.lib().xtpb_gw_sigma_c_offdiag(self._h, _d(fr), _d(out)))
        return out

    def CalculateGWPerturbation(self):
        check(_lib.lib().xtpb_gw_calculate_gw_perturbation(self._h))

    def CalculateHQP(self):
        check(_lib.lib().xtpb_gw_calculate_hqp(self._h))

    def getGWAResults(self):
        out = np.empty(self.qptotal)
        check(_lib.lib().xtpb_gw_get_gwa_results(self._h, _d(out)))
        return out

    def getHQP(self):
        out = np.empty((self.qptotal, self.qptotal), order="F")
        check(_lib.lib().xtpb_gw_get_hqp(self._h, _d(out)))
        return out

    def DiagonalizeQPHamiltonian(self):
        w = np.empty(self.qptotal)
        v = np.empty((self.qptotal, self.qptotal), order="F")
        check(_lib.lib().xtpb_gw_diagonalize_qp_hamiltonian(self._h, _d(w), _d(v)))
        return w, v

    def unconverged_levels(self):
        n = idx(0)
        check(_lib.lib().xtpb_gw_unconverged_levels(self._h, C.byref(n)))
        return int(n.value)


class _Operator:
    def __init__(self, handle, ctx):
        self._h, self.ctx = handle, ctx
        n = idx(0)
        check(_lib.lib().xtpb_op_size(self._h, C.byref(n)))
        self._n = int(n.value)

    def close(self):
        if self._h:
            _lib.lib().xtpb_op_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rows(self): return self._n
    def cols(self): return self._n
    def size(self): return self._n

    def matmul(self, X):
        Xf = _f(X)
        if Xf.ndim == 1:
            Xf = Xf.reshape(-1, 1, order="F")
        k = Xf.shape[1]
        Y = np.empty((self._n, k), order="F")
        check(_lib.lib().xtpb_op_matmul(self._h, _d(Xf), Xf.shape[0], k, _d(Y), self._n))
        return Y

    def diagonal(self):
        d = np.empty(self._n)
        check(_lib.lib().xtpb_op_diagonal(self._h, _d(d)))
        return d

    def get_full_matrix(self):
        H = np.empty((self._n, self._n), order="F")
        check(_lib.lib().xtpb_op_get_full_matrix(self._h, _d(H), self._n))
        return H


OPERATOR_TYPES = {
    "SingletOperator_TDA": (1, 2, 1, 0),
    "TripletOperator_TDA": (1, 0, 1, 0),
    "SingletOperator_BTDA_B": (0, 2, 0, 1),
    "TripletOperator_BTDA_B": (0, 0, 0, 1),
    "HxOperator": (0, 1, 0, 0),
    "HdOperator": (0, 0, 1, 0),
    "Hd2Operator": (0, 0, 0, 1),
    "HqpOperator": (1, 0, 0, 0),
}


class BSE_OPERATOR(_Operator):
    """BSE_OPERATOR<cqp,cx,cd,cd2>(epsilon_0_inv, Mmn, Hqp) + configure(opt); upstream bse_operator.{h,cc}"""

    def __init__(self, ctx: Context, cqp, cx, cd, cd2, epsilon_0_inv, Mmn: TCMatrix_gwbse, Hqp, homo, rpamin, vmin, cmax):
        e = np.ascontiguousarray(epsilon_0_inv, dtype=np.float64)
        H = _f(Hqp)
        h = vp()
        check(_lib.lib().xtpb_bse_operator_create_raw(ctx._h, Mmn._h, homo, rpamin, vmin, cmax, _d(e), _d(H),
                                                      H.shape[0], cqp, cx, cd, cd2, C.byref(h)))
        super().__init__(h, ctx)


class DenseOperator(_Operator):
    def __init__(self, ctx: Context, A):
        Af = _f(A)
        h = vp()
        check(_lib.lib().xtpb_dense_operator_create(ctx._h, _d(Af), Af.shape[0], Af.shape[0], C.byref(h)))
        super().__init__(h, ctx)


class DavidsonSolver:
    """upstream xtp/src/libxtp/davidsonsolver.cc"""
    TOL = {"loose": 1e-3, "normal": 1e-4, "strict": 1e-5, "lapack": 1e-9}

    def __init__(self):
        self.opt = _lib.DavidsonOptions()
        _lib.lib().xtpb_davidson_options_default(C.byref(self.opt))
        self._info, self._iters = 1, 0

    def set_iter_max(self, n): self.opt.iter_max = int(n)
    def set_max_search_space(self, n): self.opt.max_search_space = int(n)
    def set_tolerance(self, name): self.opt.tolerance = self.TOL[name]
    def set_correction(self, name): self.opt.correction = {"DPR": 0, "OLSEN": 1}[name.upper()]
    def set_size_update(self, name): self.opt.size_update = {"min": 0, "safe": 1, "max": 2}[name.lower()]

    def solve(self, A: _Operator, neigen, size_initial_guess=0):
        self.opt.size_initial_guess = int(size_initial_guess)
        n = A.rows()
        self._evals = np.empty(neigen)
        self._evecs = np.empty((n, neigen), order="F")
        info, iters = C.c_int(1), idx(0)
        check(_lib.lib().xtpb_davidson_solve(A._h, int(neigen), C.byref(self.opt), _d(self._evals), _d(self._evecs), n,
                                             C.byref(info), C.byref(iters)))
        self._info, self._iters = info.value, int(iters.value)
        return self

    def eigenvalues(self): return self._evals
    def eigenvectors(self): return self._evecs
    def info(self): return "Success" if self._info == 0 else "NoConvergence"
    def num_iterations(self): return self._iters


class BSE:
    """upstream xtp/src/libxtp/gwbse/bse.cc"""

    def __init__(self, ctx: Context, Mmn: TCMatrix_gwbse):
        self.ctx, self.Mmn = ctx, Mmn
        self._h = vp()

    def configure(self, homo, rpamin, rpamax, qpmin, qpmax, vmin, cmax, nmax, RPAInputEnergies, Hqp,
                  use_Hqp_offdiag=True, rotate_full_tc=False, davidson_correction="DPR", davidson_tolerance="normal",
                  davidson_update="safe", davidson_maxiter=50):
        self.close()
        o = _lib.BseOptions(homo, rpamin, rpamax, qpmin, qpmax, vmin, cmax, nmax, int(use_Hqp_offdiag))
        self.opt = o
        self.davidson = (davidson_correction, davidson_tolerance, davidson_update, davidson_maxiter)
        e = np.ascontiguousarray(RPAInputEnergies, dtype=np.float64)
        H = _f(Hqp)
        check(_lib.lib().xtpb_bse_create(self.ctx._h, self.Mmn._h, C.byref(o), _d(e), _d(H), H.shape[0],
                                         int(rotate_full_tc), C.byref(self._h)))

    def close(self):
        if self._h:
            _lib.lib().xtpb_bse_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def Perturbative_DynamicalScreening(self, energies, X, Y=None, max_dyn_iter=10, dyn_tolerance=1e-5):
        """BSE::Perturbative_DynamicalScreening: dynamically screened energies (and iterations per state)."""
        e = np.ascontiguousarray(energies, dtype=np.float64)
        Xf = _f(X)
        Yf = _f(Y) if Y is not None else None
        out = np.empty(len(e))
        its = np.zeros(len(e), dtype=np.int64)
        check(_lib.lib().xtpb_bse_perturbative_dynamical_screening(
            self._h, len(e), _d(e), _d(Xf), _d(Yf) if Yf is not None else None, Xf.shape[0], int(max_dyn_iter),
            float(dyn_tolerance), _d(out), its.ctypes.data_as(iptr)))
        self.dynamical_iterations = its
        return out

    def eps0_reused(self):
        """True when BSE::configure took the eps(0) eigenvalues from the PPM rotation (xtpb_bse_screening_info)."""
        flag = C.c_int(0)
        check(_lib.lib().xtpb_bse_screening_info(self._h, C.byref(flag)))
        return bool(flag.value)

    def epsilon_0_inv(self):
        out = np.empty(self.Mmn.auxsize())
        check(_lib.lib().xtpb_bse_get_epsilon_0_inv(self._h, _d(out)))
        return out

    def make_operator(self, name):
        cqp, cx, cd, cd2 = OPERATOR_TYPES[name]
        h = vp()
        check(_lib.lib().xtpb_bse_operator_create(self._h, cqp, cx, cd, cd2, C.byref(h)))
        return _Operator(h, self.ctx)

    def solve_hermitian(self, op):
        ds = DavidsonSolver()
        corr, tol, upd, maxiter = self.davidson
        ds.set_correction(corr)
        ds.set_tolerance(tol)
        ds.set_size_update(upd)
        ds.set_iter_max(maxiter)
        ds.set_max_search_space(10 * int(self.opt.nmax))
        ds.solve(op, int(self.opt.nmax))
        self.last_davidson = ds
        return ds.eigenvalues(), ds.eigenvectors()

    def Solve_singlets_TDA(self):
        op = self.make_operator("SingletOperator_TDA")
        try:
            return self.solve_hermitian(op)
        finally:
            op.close()

    def Solve_triplets_TDA(self):
        op = self.make_operator("TripletOperator_TDA")
        try:
            return self.solve_hermitian(op)
        finally:
            op.close()

    def _solve_btda(self, singlet):
        """Full BSE (useTDA = false): energies, X, Y with X^T X - Y^T Y = 1 (upstream Solve_nonhermitian_Davidson)."""
        ds = DavidsonSolver()
        corr, tol, upd, maxiter = self.davidson
        ds.set_tolerance(tol)
        ds.set_iter_max(maxiter)
        ds.set_max_search_space(10 * int(self.opt.nmax))
        n, k = self.size(), int(self.opt.nmax)
        e = np.empty(k)
        X, Y = np.empty((n, k), order="F"), np.empty((n, k), order="F")
        info, iters = C.c_int(0), idx(0)
        check(_lib.lib().xtpb_bse_solve_btda(self._h, int(singlet), C.byref(ds.opt), _d(e), _d(X), _d(Y), n,
                                             C.byref(info), C.byref(iters)))
        self.last_btda = {"info": "Success" if info.value == 0 else "NoConvergence", "iterations": int(iters.value)}
        return e, X, Y

    def Solve_singlets_BTDA(self):
        return self._solve_btda(True)

    def Solve_triplets_BTDA(self):
        return self._solve_btda(False)

    def size(self):
        return int((self.opt.homo - self.opt.vmin + 1) * (self.opt.cmax - self.opt.homo))

    def transition_dipoles(self, ao_dipoles, C_mo, X, Y=None):
        """BSE::CalcCoupledTransition_Dipoles: (n_states, 3) from AO dipole matrices (3, nb, nb) and MO coefficients."""
        Cf = _f(C_mo)
        R = np.ascontiguousarray(ao_dipoles, dtype=np.float64)
        Xf = _f(X)
        Yf = _f(Y) if Y is not None else None
        out = np.empty((Xf.shape[1], 3))
        check(_lib.lib().xtpb_bse_transition_dipoles(self._h, Cf.shape[0], _d(Cf), Cf.shape[0], _d(R), Xf.shape[1],
                                                     _d(Xf), _d(Yf) if Yf is not None else None, Xf.shape[0], _d(out)))
        return out


def alloc_stats(reset=False):
    """Host seconds inside cudaMalloc/cudaFree, allocation count, block-cache hits and cached GB (xtpb_alloc_stats)."""
    sec, cached = C.c_double(0.0), C.c_double(0.0)
    calls, hits = C.c_longlong(0), C.c_longlong(0)
    wait = C.c_double(0.0)
    check(_lib.lib().xtpb_alloc_free_wait_seconds(C.byref(wait)))        # before a reset zeroes it
    check(_lib.lib().xtpb_alloc_stats(C.byref(sec), C.byref(calls), C.byref(hits), C.byref(cached), int(reset)))
    return {"seconds": sec.value, "calls": calls.value, "cache_hits": hits.value, "cached_gb": cached.value * 1e-9,
            "free_wait_seconds": wait.value}


RANGES = {"default": 0, "factor": 1, "explicit": 2, "full": 3}


def gwbse_level_ranges(mode, n_levels, n_occ, rpamax=0.0, qpmin=0.0, qpmax=0.0, bsemin=0.0, bsemax=0.0,
                       n_core_ignored=0):
    """GWBSE::Initialize (upstream gwbse/gwbse.cc): level ranges from the ``ranges`` option (host code in libxtpb200)."""
    o = _lib.GwbseRangeOptions(RANGES[mode], int(n_levels), int(n_occ), int(n_core_ignored), float(rpamax), float(qpmin),
                               float(qpmax), float(bsemin), float(bsemax))
    r = _lib.GwbseRanges()
    check(_lib.lib().xtpb_gwbse_level_ranges(C.byref(o), C.byref(r)))
    return {k: int(getattr(r, k)) for k, _ in _lib.GwbseRanges._fields_}


class GWBSE:
    """The driver of the path, upstream ``GWBSE`` (gwbse/gwbse.cc): ``Initialize`` fixes the level ranges and the
    options, ``Evaluate`` runs Fill -> G0W0/evGW -> Hqp -> BSE in the reference's order and returns what the reference
    stores into ``Orbitals`` (QPpert energies, Hqp, RPA input energies, BSE singlet/triplet energies and coefficients)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def Initialize(self, n_levels, n_occ, ranges="default", tasks=("gw", "singlets"), nmax=5, useTDA=True,
                   sigma_integration="ppm", qp_solver="grid", davidson_tolerance="normal", gw_sc_max_iterations=1,
                   ScaHFX=0.0, **range_values):
        self.r = gwbse_level_ranges(ranges, n_levels, n_occ, **range_values)
        self.tasks, self.nmax, self.useTDA = tuple(tasks), int(nmax), bool(useTDA)
        self.gwopt = dict(sigma_integration=sigma_integration, qp_solver=qp_solver,
                          gw_sc_max_iterations=gw_sc_max_iterations, ScaHFX=ScaHFX)
        self.davidson_tolerance = davidson_tolerance
        return self

    def Evaluate(self, ao3c, C_mo, dft_energies, vxc, aux_coulomb, aux_overlap=None):
        r = self.r
        tc = TCMatrix_gwbse(self.ctx).Initialize(np.asarray(ao3c).shape[0], r["rpamin"], max(r["qpmax"], r["cmax"]),
                                                  r["rpamin"], r["rpamax"])
        tc.Fill(ao3c, C_mo, aux_coulomb, aux_overlap)
        out = {"ranges": dict(r)}
        gw = GW(self.ctx, tc, vxc, dft_energies)
        gw.configure(gw_options(homo=r["homo"], qpmin=r["qpmin"], qpmax=r["qpmax"], rpamin=r["rpamin"],
                                rpamax=r["rpamax"], **self.gwopt))
        gw.CalculateGWPerturbation()
        out["QPpert_energies"] = gw.getGWAResults()
        gw.CalculateHQP()
        out["Hqp"] = gw.getHQP()
        out["QPdiag_energies"], out["QPdiag_coefficients"] = gw.DiagonalizeQPHamiltonian()
        out["RPA_input_energies"] = gw.RPAInputEnergies()
        gw.close()
        if "singlets" in self.tasks or "triplets" in self.tasks:
            bse = BSE(self.ctx, tc)
            bse.configure(r["homo"], r["rpamin"], r["rpamax"], r["qpmin"], r["qpmax"], r["vmin"], r["cmax"], self.nmax,
                          out["RPA_input_energies"], out["Hqp"], davidson_tolerance=self.davidson_tolerance)
            for task, singlet in (("singlets", True), ("triplets", False)):
                if task not in self.tasks:
                    continue
                if self.useTDA:
                    e, X = bse.Solve_singlets_TDA() if singlet else bse.Solve_triplets_TDA()
                    out[f"BSE_{task[:-1]}_energies"], out[f"BSE_{task[:-1]}_coefficients"] = e, X
                else:
                    e, X, Y = bse.Solve_singlets_BTDA() if singlet else bse.Solve_triplets_BTDA()
                    out[f"BSE_{task[:-1]}_energies"], out[f"BSE_{task[:-1]}_coefficients"] = e, X
                    out[f"BSE_{task[:-1]}_coefficients_AR"] = Y
            bse.close()
        tc.close()
        return out


def oscillator_strengths(energies, dipoles):
    """Orbitals::Oscillatorstrengths: f = 2/3 E |d|^2."""
    e = np.ascontiguousarray(energies, dtype=np.float64)
    d = np.ascontiguousarray(dipoles, dtype=np.float64)
    out = np.empty(len(e))
    check(_lib.lib().xtpb_oscillator_strengths(len(e), _d(e), _d(d), _d(out)))
    return out


def contract_host(ctx: Context, desc: "_lib.ContractDesc", A, B, d, Cmat):
    """Engine-level test hook: raw strided contraction on host buffers (flat float64 arrays)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    Cmat = np.ascontiguousarray(Cmat, dtype=np.float64)
    desc.a_len, desc.b_len, desc.c_len = A.size, B.size, Cmat.size
    if d is not None:
        d = np.ascontiguousarray(d, dtype=np.float64)
        desc.d_len = d.size
    else:
        desc.d_len = 0
    check(_lib.lib().xtpb_contract_host(ctx._h, C.byref(desc), _d(A), _d(B), _d(d) if d is not None else None, _d(Cmat)))
    return Cmat


def contract_plan(desc: "_lib.ContractDesc", n_sms=148):
    """(tile_cfg, split_k) the launcher would choose for this shape on n_sms SMs (xtpb_contract_plan; no device needed)."""
    cfg, sp = C.c_int(0), C.c_int(0)
    check(_lib.lib().xtpb_contract_plan(C.byref(desc), int(n_sms), C.byref(cfg), C.byref(sp)))
    return int(cfg.value), int(sp.value)


def contract_bench(ctx: Context, desc: "_lib.ContractDesc", reps=10):
    ms = C.c_double()
    check(_lib.lib().xtpb_contract_bench(ctx._h, C.byref(desc), int(reps), C.byref(ms)))
    return ms.value
