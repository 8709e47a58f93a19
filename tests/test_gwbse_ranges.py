"""GWBSE::Initialize level-range logic (upstream gwbse/gwbse.cc `ranges` option): the host code inside libxtpb200
(xtpb_gwbse_level_ranges, no device needed) against the oracle's restatement, plus the invariants every consumer
relies on (windows contain HOMO and LUMO, BSE/QP windows inside the RPA window)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import gwbse_oracle as orc
from xtp_b200 import api, synth
from xtp_b200._lib import XtpbError

KEYS = ("homo", "rpamin", "rpamax", "qpmin", "qpmax", "vmin", "cmax")


def _same(got, ref):
    assert {k: got[k] for k in KEYS} == ref


def test_default_matches_synth_sizes_convention():
    for name in ("ch4-svp-shape", "benzene-tzvp-shape", "c60-tzvp-shape"):
        sz = synth.WORKLOADS[name]
        r = api.gwbse_level_ranges("default", sz.n_basis, sz.homo + 1)
        assert (r["rpamin"], r["rpamax"], r["qpmin"], r["qpmax"], r["vmin"], r["cmax"]) == \
            (sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax)
        assert (r["qptotal"], r["rpatotal"], r["bse_size"]) == (sz.qptotal, sz.ntotal, sz.bse_size)


def test_known_cases():
    # methane/3-21G-like: 17 levels, 5 occupied
    _same(api.gwbse_level_ranges("default", 17, 5), dict(homo=4, rpamin=0, rpamax=16, qpmin=0, qpmax=9, vmin=0, cmax=9))
    _same(api.gwbse_level_ranges("full", 17, 5), dict(homo=4, rpamin=0, rpamax=16, qpmin=0, qpmax=16, vmin=0, cmax=16))
    _same(api.gwbse_level_ranges("factor", 100, 10, rpamax=0.5, qpmin=0.5, qpmax=2.0, bsemin=0.3, bsemax=1.0),
          dict(homo=9, rpamin=0, rpamax=49, qpmin=4, qpmax=29, vmin=6, cmax=19))
    _same(api.gwbse_level_ranges("explicit", 100, 10, rpamax=80, qpmin=2, qpmax=30, bsemin=5, bsemax=25, n_core_ignored=2),
          dict(homo=9, rpamin=2, rpamax=80, qpmin=2, qpmax=30, vmin=5, cmax=25))
    # clamps: out-of-range requests shrink to the available levels and keep HOMO/LUMO inside
    _same(api.gwbse_level_ranges("explicit", 20, 4, rpamax=500, qpmin=-3, qpmax=1, bsemin=9, bsemax=400),
          dict(homo=3, rpamin=0, rpamax=19, qpmin=0, qpmax=4, vmin=3, cmax=19))


def test_bad_input_is_an_error():
    with pytest.raises(XtpbError):
        api.gwbse_level_ranges("default", 10, 10)          # no empty level
    with pytest.raises(XtpbError):
        api.gwbse_level_ranges("default", 10, 4, n_core_ignored=4)


@settings(max_examples=200, deadline=None)
@given(n_levels=st.integers(2, 400), frac_occ=st.floats(0.01, 0.99), mode=st.sampled_from(["default", "factor", "explicit", "full"]),
       f=st.lists(st.floats(0.0, 3.0), min_size=5, max_size=5), e=st.lists(st.integers(-5, 450), min_size=5, max_size=5),
       core=st.integers(0, 3))
def test_library_matches_oracle_and_invariants(n_levels, frac_occ, mode, f, e, core):
    n_occ = min(n_levels - 1, max(1, int(frac_occ * n_levels)))
    core = min(core, n_occ - 1)
    vals = f if mode == "factor" else [float(x) for x in e]
    kw = dict(rpamax=vals[0], qpmin=vals[1], qpmax=vals[2], bsemin=vals[3], bsemax=vals[4], n_core_ignored=core)
    got = api.gwbse_level_ranges(mode, n_levels, n_occ, **kw)
    _same(got, orc.gwbse_level_ranges(mode, n_levels, n_occ, **kw))
    assert got["rpamin"] <= got["qpmin"] <= got["homo"] < got["qpmax"] <= got["rpamax"] <= n_levels - 1
    assert got["rpamin"] <= got["vmin"] <= got["homo"] < got["cmax"] <= got["rpamax"]
    assert got["bse_size"] == (got["homo"] - got["vmin"] + 1) * (got["cmax"] - got["homo"])
