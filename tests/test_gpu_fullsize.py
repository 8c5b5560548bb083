"""BASELINE.json's full-size synthetic workload (1024 x 1024 gridworld, config 4) through size-independent properties:
the oracle cannot finish these sizes in seconds, so the checks are the known answers of the algebra (log Z == 0 and
gap == 0 at V = nu, `test/runtests.jl`-style bounds), conservation (every path lands in exactly one bucket), determinism
(same seed -> same bits) and independence of the tiling.  -m gpu."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(engine):
    import bench
    walls, goal, params = bench.make_world("grid1024")
    engine.load_grid(walls, goal, **params)
    nu, st = engine.solve_nu(tol=0.0, max_sweeps=3000, precision="f64")     # 0.99^3000 ~ 1e-13
    return walls, goal, params, nu


def test_nu_solution_properties(engine, big):
    walls, goal, params, nu = big
    assert np.array_equal(np.isnan(nu), walls != 0)
    assert nu[goal] == 0.0
    free = walls == 0
    # soft value of a cost-per-step problem: bounded by the geometric sums of the extreme rewards (+ log-prior slack)
    lo = -params["step_cost"] / (1 - params["gamma"]) - 1e-9
    hi = params["reward"] + np.log(4.0) / (1 - params["gamma"])
    assert lo <= np.nanmin(nu) and np.nanmax(nu) <= hi
    # fixed point: one more sweep changes nothing
    engine.vi_prepare("nu", precision="f64")
    engine.vi_set_iterate(nu)
    engine.vi_sweeps(1)
    x, res = engine.vi_fetch()
    assert res < 1e-6 and float(np.nanmax(np.abs(x - nu))) < 1e-10
    # fp32 modes agree with Float64 to 1e-5 relative at full size
    for prec in ("f32", "f32fast"):
        n32, _ = engine.solve_nu(tol=0.0, max_sweeps=3000, precision=prec)
        rel = np.abs(n32[free] - nu[free]) / np.maximum(1.0, np.abs(nu[free]))
        assert rel.max() < 1e-5, (prec, rel.max())


def test_logz_and_gap_vanish_at_the_fixed_point(engine, big):
    walls, goal, params, nu = big
    V = np.nan_to_num(nu)
    logz, _ = engine.solve_logz(V, beta=0.3, tol=0.0, max_sweeps=2000, precision="f64")
    assert float(np.nanmax(np.abs(logz))) < 1e-8            # logz_updateV leaves nu unchanged
    gap, _ = engine.solve_gap(V, V, tol=0.0, max_sweeps=400, precision="f64")
    assert float(np.nanmax(np.abs(gap))) < 1e-8
    # reference runtests bounds (test/runtests.jl:33-36) from a weaker start, one update at beta = 1 and beta -> 0
    vlo, _ = engine.solve_nu(reward_scale=0.8, tol=0.0, max_sweeps=3000)
    up, _ = engine.update_v(np.nan_to_num(vlo), beta=1.0, tol=0.0, max_sweeps=3000, allow_unconverged=True)
    free = walls == 0
    assert np.all(up[free] >= nu[free] - 0.01)


def test_tiling_independence_at_full_size(engine, big):
    walls, goal, params, nu = big
    x0 = np.asfortranarray(np.random.default_rng(0).normal(size=walls.shape))
    outs = []
    try:
        for tj, nw in (("21", "9"), ("38", "16"), ("82", "15")):
            os.environ["PIQL_TILE_TJ"], os.environ["PIQL_TILE_NW"] = tj, nw
            engine.vi_prepare("nu", precision="f32fast")
            engine.vi_set_iterate(x0)
            engine.vi_sweeps(9)
            outs.append(np.nan_to_num(engine.vi_fetch()[0]))
    finally:
        os.environ.pop("PIQL_TILE_TJ", None)
        os.environ.pop("PIQL_TILE_NW", None)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_rollout_conservation_determinism_and_known_answer(engine, big):
    walls, goal, params, nu = big
    V = np.nan_to_num(nu)
    n = 1 << 24
    sa, cnt, ls, st = engine.mc_logz(V, beta=1.0, npaths=n, seed=11)
    assert int(cnt.sum()) == n                                 # every path lands in exactly one (start, action) bucket
    nfree = int((walls == 0).sum())
    per = cnt.reshape(walls.shape[1], walls.shape[0], 4).sum(axis=2).T
    assert per[walls != 0].sum() == 0 and per[walls == 0].min() >= n // nfree and per[walls == 0].max() <= n // nfree + 1
    assert 50 * n < st.transitions < 101 * n                   # geometric horizon 1/(1-gamma) = 100, cut short at the goal
    # same seed, same bits; other seed, other paths
    sa2, cnt2, ls2, _ = engine.mc_logz(V, beta=1.0, npaths=n, seed=11)
    assert np.array_equal(cnt, cnt2) and np.array_equal(np.nan_to_num(ls), np.nan_to_num(ls2))
    sa3, cnt3, ls3, _ = engine.mc_logz(V, beta=1.0, npaths=n, seed=12)
    assert not np.array_equal(cnt, cnt3)
    # known answer: at V = nu the exact log Z is 0 everywhere; the per-cell estimates (20 paths each) scatter around it
    free = walls == 0
    z = np.exp(ls[free])
    assert abs(z.mean() - 1.0) < 5e-3, z.mean()


def test_config5_4096_grid_properties(engine):
    """BASELINE config 5 (4096 x 4096, one GPU's share of the work): fp32-fast soft VI reaches a fixed point that one more
    Float64 sweep barely moves and that respects the value bounds; rollouts conserve paths and are deterministic."""
    import bench
    walls, goal, params = bench.make_world("grid4096")
    engine.load_grid(walls, goal, **params)
    free = walls == 0
    nu32, st = engine.solve_nu(tol=0.0, max_sweeps=2000, precision="f32fast")
    assert st.kernel == 2 and np.array_equal(np.isnan(nu32), walls != 0) and nu32[goal] == 0.0
    engine.vi_prepare("nu", precision="f64")
    engine.vi_set_iterate(nu32)
    engine.vi_sweeps(1)
    x, res = engine.vi_fetch()
    rel = np.abs(x[free] - nu32[free]) / np.maximum(1.0, np.abs(nu32[free]))
    assert rel.max() < 2e-6, rel.max()              # fp32 fixed point: a Float64 sweep moves it by fp32 rounding only
    lo = -params["step_cost"] / (1 - params["gamma"]) - 1e-5
    assert lo <= np.nanmin(nu32) and np.nanmax(nu32) <= params["reward"] + np.log(4.0) / (1 - params["gamma"])
    V = np.nan_to_num(nu32)
    n = 1 << 24
    sa, cnt, ls, st1 = engine.mc_logz(V, beta=1.0, npaths=n, seed=3, want=("count", "s"))
    assert int(cnt.sum()) == n and 50 * n < st1.transitions < 101 * n
    sa2, cnt2, ls2, _ = engine.mc_logz(V, beta=1.0, npaths=n, seed=3, want=("count", "s"))
    assert np.array_equal(cnt, cnt2) and np.array_equal(np.nan_to_num(ls), np.nan_to_num(ls2))
    started = cnt.reshape(walls.shape[1], walls.shape[0], 4).sum(axis=2).T > 0
    z = np.exp(ls[started])
    assert abs(z.mean() - 1.0) < 2e-2, z.mean()     # exact log Z = 0 at V = nu (one or two paths per cell here)
