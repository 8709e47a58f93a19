"""The contraction launcher's plan (tile configuration, split-K) on a CPU box: xtpb_contract_plan is host arithmetic
only (xtp_b200/csrc/contract.cu: contract_plan).  Pins the decisions DESIGN.md section 3.1 describes for the shapes of
the C60-size step on a 148-SM B200."""
import pytest

from xtp_b200 import _lib, api


def desc(**kw):
    d = _lib.ContractDesc()
    d.n_outer = d.n_batch = 1
    d.alpha, d.beta, d.force_cfg, d.force_splits = 1.0, 0.0, -1, 0
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def test_tile_width_follows_the_column_count():
    assert api.contract_plan(desc(M=4096, N=20, K=4096))[0] == 2          # <= 32 columns: 128x32 tiles
    assert api.contract_plan(desc(M=4096, N=64, K=4096))[0] == 1          # <= 64: 128x64
    assert api.contract_plan(desc(M=4096, N=180, K=4096))[0] == 1         # 180 -> 192 with 64-wide, 256 with 128-wide
    assert api.contract_plan(desc(M=4096, N=5500, K=4096))[0] == 0
    assert api.contract_plan(desc(M=360, N=180, K=64, lower=1))[0] == 0   # triangular outputs: square tiles


def test_small_grids_are_split_to_about_three_waves():
    # Sigma_x at C60 size: 3 x 3 tiles, K = 180 x 5500 outer blocks
    cfg, s = api.contract_plan(desc(M=360, N=360, K=180, n_outer=5500, lower=1))
    assert cfg == 0 and 40 <= s <= 64
    # a product with fewer than 16 k-tiles is never split
    assert api.contract_plan(desc(M=128, N=128, K=128))[1] == 1


def test_tail_balancing_split_of_the_epsilon_syrk():
    """946 lower-triangle tiles on 148 SMs = 6.4 waves of 40 ms tiles: five splits turn the seven waves into 6.4."""
    eps = desc(M=5500, N=5500, K=1680, n_outer=180, lower=1)
    assert api.contract_plan(eps) == (0, 5)
    # per rank of eight (210 unoccupied levels each): same tile count, shorter tiles, still worth it
    assert api.contract_plan(desc(M=5500, N=5500, K=210, n_outer=180, lower=1))[1] == 5
    # forced values are kept
    eps.force_splits = 2
    assert api.contract_plan(eps) == (0, 2)


@pytest.mark.parametrize("shape", [
    dict(M=64 * 1860, N=360, K=1860),                 # Fill3cMO, first half of a 64-slice group (rows folded)
    dict(M=1860, N=64 * 360, K=1860),                 # second half (columns folded)
    dict(M=1860, N=5500, K=5500, n_batch=52),         # aux rotation chunk
    dict(M=32400, N=16290, K=5500),                   # dense BSE direct term over the occupied pairs
    dict(M=5500, N=5500, K=5500),                     # N_aux^3 products
])
def test_many_wave_or_short_contractions_are_left_alone(shape):
    assert api.contract_plan(desc(**shape))[1] == 1


def test_plan_depends_on_the_sm_count():
    eps = desc(M=5500, N=5500, K=1680, n_outer=180, lower=1)
    # 946 tiles on 132 SMs: 7.17 waves -> 8; two splits give 7.5, three 7.33, ... the planner must not pick 1
    assert api.contract_plan(eps, n_sms=132)[1] > 1
    # 946 = 2 x 11 x 43 tiles on 86 SMs: exactly 11 waves, nothing to gain
    assert api.contract_plan(eps, n_sms=86)[1] == 1
