"""CPU tests of the compressed Sigma_c grid scan: the host plan exported by libxtpb200 (xtpb_ppm_grid_plan, no device
needed) and, through the numpy mirror of the CUDA kernels, the mathematics of the near/far split against the oracle's
direct pole sum (upstream Sigma_PPM::CalcCorrelationDiagElement on the GW::SolveQP_Grid grid)."""
import numpy as np
import pytest

import ppm_grid_mirror as mir

W = 0.25          # damping half-window of Sigma_PPM::Stabilize
BIN = 0.125       # kPpmGridBinWidth


def _check_plan(edges, near, grid_start, spacing, steps, zmin, zmax):
    nb = len(edges) - 1
    assert nb >= 1 and np.all(np.diff(edges) > 0)
    assert edges[0] <= zmin + 1e-12 or edges[1] > zmin      # the first kept bin reaches down to the lowest pole
    assert edges[-1] >= zmax - 1e-12 or edges[-2] <= zmax
    c, h = 0.5 * (edges[:-1] + edges[1:]), 0.5 * np.diff(edges)
    assert h.min() >= 0.5 * BIN - 1e-12                      # no bin is narrower than the core bin width
    ck = mir.CHUNK
    n_chunks = (steps + ck - 1) // ck
    assert near.shape == (len(grid_start), n_chunks, 4)
    for l, g0 in enumerate(grid_start):
        for ch in range(n_chunks):
            wa, wb = g0 + spacing * ck * ch, g0 + spacing * (ck * (ch + 1) - 1)
            lo, hi, ilo, ihi = near[l, ch]
            assert 0 <= lo <= nb and -1 <= hi < nb
            # inner bins: a sub-range of the near bins (or the empty range hi+1 .. hi) whose poles are damped for every
            # target of the chunk; the near bins next to it are not
            if ilo <= ihi:
                assert lo <= ilo and ihi <= hi
                assert edges[ilo] > wb - W and edges[ihi + 1] < wa + W
                for b in (ilo - 1, ihi + 1):
                    if lo <= b <= hi:
                        assert not (edges[b] >= wb - W + 1e-6 and edges[b + 1] <= wa + W - 1e-6)
            else:
                assert (ilo, ihi) == (hi + 1, hi)
            for b in list(range(0, lo)) + list(range(hi + 1, nb)):
                dist = max(wa - c[b], c[b] - wb, 0.0)
                assert dist >= 3.0 * h[b] and dist >= h[b] + W, (l, ch, b)
                # no pole of a far bin can fall inside the damping window of any target of the chunk
                assert min(abs(wa - edges[b]), abs(wa - edges[b + 1]), abs(wb - edges[b]), abs(wb - edges[b + 1])) >= W - 1e-12


@pytest.mark.parametrize("steps,spacing", [(1001, 0.01), (173, 0.013), (33, 0.2), (1, 0.01)])
def test_plan_invariants(steps, spacing):
    rng = np.random.default_rng(5)
    centers = np.sort(rng.uniform(-1.0, 0.6, 12))
    grid_start = centers - spacing * (steps - 1) / 2
    for zmin, zmax in [(-9.0, 14.0), (-0.3, 0.2), (40.0, 400.0), (-2.0, -2.0), (-1e4, 1e4)]:
        pl = mir.plan(grid_start, spacing, steps, zmin, zmax)
        assert pl is not None
        _check_plan(pl[0], pl[1], grid_start, spacing, steps, zmin, zmax)


def test_plan_declines_bad_input():
    g = np.array([0.0, 1.0])
    assert mir.plan(g, 0.01, 1001, 1.0, -1.0) is None            # no live poles
    assert mir.plan(g, 0.01, 1001, -np.inf, 1.0) is None
    assert mir.plan(np.array([0.0, 1e4]), 0.01, 1001, -1.0, 1.0) is None   # levels too far apart: too many bins


def _problem(seed, ntot, n_occ, naux, wide):
    rng = np.random.default_rng(seed)
    e = np.concatenate([np.sort(rng.uniform(-1.0, -0.3, n_occ)), np.sort(rng.uniform(0.0, 3.0, ntot - n_occ))])
    freq = rng.uniform(0.3, 40.0 if wide else 2.0, naux)
    weight = rng.uniform(0.05, 0.6, naux)
    weight[rng.integers(0, naux, 3)] = 0.0                    # dropped plasmon poles (fac = 0)
    fac = np.where(weight < 1e-9, 0.0, 0.5 * weight * freq)
    slab = rng.standard_normal((naux, ntot)) / np.sqrt(naux)
    return e, freq, fac, slab


@pytest.mark.parametrize("seed,ntot,n_occ,naux,wide,steps,spacing", [
    (1, 40, 6, 50, False, 1001, 0.01),
    (2, 64, 9, 30, True, 1001, 0.01),
    (3, 33, 0, 20, False, 173, 0.013),       # no occupied levels on this rank
    (4, 17, 17, 20, True, 257, 0.02),        # no unoccupied levels
    (5, 50, 10, 40, False, 31, 0.3),         # one ragged chunk, coarse grid: everything near or everything far
])
def test_compressed_scan_matches_direct_sum(seed, ntot, n_occ, naux, wide, steps, spacing):
    e, freq, fac, slab = _problem(seed, ntot, n_occ, naux, wide)
    zmin, zmax = mir.pole_range(e, n_occ, freq, fac)
    levels = [0, max(0, n_occ - 1), min(ntot - 1, n_occ), ntot - 1]
    grid_start = np.array([e[l] - spacing * (steps - 1) / 2 for l in levels])
    pl = mir.plan(grid_start, spacing, steps, zmin, zmax)
    assert pl is not None
    edges, near = pl
    counters = {"near": 0, "all": 0}
    for i, l in enumerate(levels):
        got = mir.grid_values(slab, e, n_occ, freq, fac, grid_start[i], spacing, steps, edges, near[i], counters)
        ref = mir.direct_values(slab, e, n_occ, freq, fac, grid_start[i], spacing, steps)
        np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-13)
    if steps == 1001 and not wide:
        assert counters["near"] < 0.35 * counters["all"]      # the point of the exercise


def test_compressed_scan_sharded_over_two_ranks():
    """Multi-GPU arithmetic of the scan: the second tensor index is distributed cyclically, every
    rank plans and scans ITS columns (local energies, local occupied count, its own pole range and bins) and the partial
    grids are summed (an all-reduce on the device).  Replayed here with the numpy mirror for world = 2."""
    e, freq, fac, slab = _problem(11, 61, 11, 36, False)
    n_occ, steps, spacing = 11, 1001, 0.01
    level_energy = e[n_occ]
    grid_start = np.array([level_energy - spacing * (steps - 1) / 2])
    ref = mir.direct_values(slab, e, n_occ, freq, fac, grid_start[0], spacing, steps)
    total = np.zeros(steps)
    for rank in range(2):
        cols = np.arange(rank, len(e), 2)                     # TCMatrix::nglob: n = rank + j * world
        e_loc, slab_loc = e[cols], slab[:, cols]
        n_occ_loc = int(np.count_nonzero(cols < n_occ))       # TCMatrix::nloc_below
        zmin, zmax = mir.pole_range(e_loc, n_occ_loc, freq, fac)
        pl = mir.plan(grid_start, spacing, steps, zmin, zmax)
        assert pl is not None
        total += mir.grid_values(slab_loc, e_loc, n_occ_loc, freq, fac, grid_start[0], spacing, steps, pl[0], pl[1][0])
    np.testing.assert_allclose(total, ref, rtol=1e-11, atol=1e-13)
