"""The tiled multi-sweep soft-VI kernel (csrc/tile.cu) against the CPU oracle, through the C ABI.  -m gpu.

A tile runs 4 Jacobi sweeps in shared memory with a 4-cell halo, so the checks that matter are the ones
a halo / ghost-frame / goal-override bug would break: after exactly N sweeps from a *random* iterate the
result must equal N plain Jacobi sweeps of the oracle everywhere (1e-12 in Float64), for N that are and
are not multiples of 4, on grids whose sides are not multiples of the tile, with the goal near a tile
edge; values must not depend on the tiling; the stop-rule residual must match.
"""
import os

import numpy as np
import pytest

import oracle
from conftest import random_gridworld

pytestmark = pytest.mark.gpu

PARAMS = dict(randomness=0.05, step_cost=0.02, reward=0.2, gamma=0.97)


def nmax(a, b):
    return float(np.nanmax(np.abs(a - b)))


def world(shape, seed, goal=None):
    walls, g = random_gridworld(shape, 0.2, seed=seed)
    if goal is not None and not walls[goal]:
        g = goal
    return walls, tuple(int(v) for v in g)


def oracle_sweeps(org, kind, x, n, **kw):
    d = None
    for _ in range(n):
        x, d = org.sweep(kind, np.nan_to_num(x), **kw)
    return x, d


def gpu_sweeps(engine, kind, x, n, precision="f64", **kw):
    engine.vi_prepare(kind, precision=precision, **kw)
    engine.vi_set_iterate(x)
    engine.vi_sweeps(n)
    return engine.vi_fetch()


@pytest.mark.parametrize("shape,seed,goal", [((150, 70), 11, None), ((100, 100), 5, (4, 50)), ((260, 40), 12, (119, 20)),
                                              ((131, 129), 13, (120, 9)), ((40, 300), 14, (20, 33))])
def test_n_tiled_sweeps_equal_n_jacobi_sweeps(engine, shape, seed, goal):
    walls, goal = world(shape, seed, goal)
    engine.load_grid(walls, goal, **PARAMS)
    org = oracle.Grid(walls, goal, **PARAMS)
    rng = np.random.default_rng(seed)
    x0 = np.asfortranarray(rng.normal(size=walls.shape) * 0.5)
    V = np.asfortranarray(rng.normal(size=walls.shape) * 0.3)
    nu = np.asfortranarray(rng.normal(size=walls.shape) * 0.3)
    Q = rng.normal(size=(walls.size, 4)) * 0.2
    cases = [("nu", dict(reward_scale=0.7)), ("logz", dict(V=V, beta=0.4)), ("logz", dict(V=V, Q=Q, beta=1.0)),
             ("gap", dict(V=V, nu=nu)), ("gap", dict(V=V, nu=nu, Q=Q))]
    for kind, kw in cases:
        for n in (1, 4, 7):
            ref, dref = oracle_sweeps(org, kind, x0, n, **kw)
            out, d = gpu_sweeps(engine, kind, x0, n, **kw)
            assert np.array_equal(np.isnan(out), walls != 0)
            assert nmax(out, ref) < 1e-12, (kind, n, nmax(out, ref))
            assert abs(d - dref) < 1e-9 * max(1.0, dref), (kind, n, d, dref)


def test_values_do_not_depend_on_the_tiling(engine):
    walls, goal = world((200, 180), 21)
    engine.load_grid(walls, goal, **PARAMS)
    x0 = np.asfortranarray(np.random.default_rng(1).normal(size=walls.shape))
    outs = []
    try:
        for tj, nw in (("10", "8"), ("34", "8"), ("42", "16"), ("58", "8")):
            os.environ["PIQL_TILE_TJ"], os.environ["PIQL_TILE_NW"] = tj, nw
            for prec in ("f64", "f32", "f32fast"):
                out, _ = gpu_sweeps(engine, "nu", x0, 10, precision=prec)
                outs.append((prec, out))
    finally:
        os.environ.pop("PIQL_TILE_TJ", None)
        os.environ.pop("PIQL_TILE_NW", None)
    for prec in ("f64", "f32", "f32fast"):
        same = [o for p, o in outs if p == prec]
        for o in same[1:]:
            assert np.array_equal(np.nan_to_num(o).view(np.uint64), np.nan_to_num(same[0]).view(np.uint64)), prec


def test_tiled_solves_match_oracle(engine):
    walls, goal = world((140, 110), 31)
    p = dict(randomness=0.02, step_cost=0.02, reward=0.2, gamma=0.95)
    engine.load_grid(walls, goal, **p)
    org = oracle.Grid(walls, goal, **p)
    nu, st = engine.solve_nu(tol=1e-11, check_every=32, max_sweeps=100000)
    assert st.kernel == 2 and st.converged
    ref, _ = org.solve("nu", tol=1e-11, nthreads=8)
    assert nmax(nu, ref) < 1e-9
    V = np.nan_to_num(org.solve("nu", reward_scale=0.1, tol=1e-11, nthreads=8)[0])
    v1, st1 = engine.update_v(V, beta=0.5, tol=1e-10, max_sweeps=100000)
    assert st1.converged and nmax(v1, org.logz_updateV(V, beta=0.5, tol=1e-10, nthreads=8)) < 1e-8
    # (no converged gap solve here: the gap recursion is undiscounted and contracts only through the goal, which a
    #  near-uniform soft policy on 10^4 cells reaches after ~10^5 steps; its sweeps are covered by the N-sweep test)
    # bit-stable: the same solve twice gives identical bits and sweep counts
    nu2, st3 = engine.solve_nu(tol=1e-11, check_every=32, max_sweeps=100000)
    assert st3.sweeps == st.sweeps
    assert np.array_equal(nu.view(np.uint64), nu2.view(np.uint64))


@pytest.mark.parametrize("precision,bar", [("f32", 1e-5), ("f32fast", 1e-5)])
def test_fp32_tiled_fixed_points_within_1e5_relative(engine, precision, bar):
    """North star: fp32 value-iteration fixed points within 1e-5 relative of the Float64 ones (gamma = 0.99)."""
    walls, goal = world((160, 120), 41)
    p = dict(randomness=0.01, step_cost=0.02, reward=0.2, gamma=0.99)
    engine.load_grid(walls, goal, **p)
    nu64, _ = engine.solve_nu(tol=1e-10, max_sweeps=100000)
    nu32, _ = engine.solve_nu(precision=precision, tol=1e-4, max_sweeps=6000, allow_unconverged=True)
    free = ~np.isnan(nu64)
    rel = np.abs(nu32[free] - nu64[free]) / np.maximum(1.0, np.abs(nu64[free]))
    print(precision, "nu rel err max %.2e mean %.2e" % (rel.max(), rel.mean()))
    assert rel.max() < bar, rel.max()
    V = np.nan_to_num(engine.solve_nu(reward_scale=0.1, tol=1e-10, max_sweeps=100000)[0])
    l64, _ = engine.solve_logz(V, beta=1.0, tol=1e-10, max_sweeps=100000)
    l32, _ = engine.solve_logz(V, beta=1.0, precision=precision, tol=1e-4, max_sweeps=8000, allow_unconverged=True)
    rel = np.abs(l32[free] - l64[free]) / np.maximum(1.0, np.abs(l64[free]))
    print(precision, "logz rel err max %.2e mean %.2e" % (rel.max(), rel.mean()))
    assert rel.max() < bar, rel.max()


def test_fast_logz_reports_when_it_leaves_the_float_range(engine, piql):
    """PIQL_F32_FAST log Z sweeps iterate on Z = exp(log Z) (linear domain, no transcendentals in a sweep): an iterate beyond
    the float exponent range must be reported (PIQL_ERR_UNSUPPORTED), never returned as inf / NaN; the next solve on the context
    starts clean, an underflowing start recovers (the 1 - gamma term is positive), and PIQL_F32 has no such limit."""
    walls, goal = world((150, 70), 11)
    engine.load_grid(walls, goal, **PARAMS)
    V = np.asfortranarray(np.random.default_rng(2).normal(size=walls.shape) * 0.3)
    free = walls == 0
    engine.vi_prepare("logz", precision="f32fast", V=V, beta=0.7)
    engine.vi_set_iterate(np.asfortranarray(np.full(walls.shape, 400.0)))
    engine.vi_sweeps(4)
    with pytest.raises(piql.PiqlError) as e:
        engine.vi_fetch()
    assert e.value.code == piql._capi.ERR_UNSUPPORTED
    out, _ = gpu_sweeps(engine, "logz", np.full(walls.shape, 400.0), 4, precision="f32", V=V, beta=0.7)
    assert np.isfinite(out[free]).all()
    ref, _ = oracle_sweeps(oracle.Grid(walls, goal, **PARAMS), "logz", np.full(walls.shape, -400.0), 6, V=V, beta=0.7)
    out, _ = gpu_sweeps(engine, "logz", np.full(walls.shape, -400.0), 6, precision="f32fast", V=V, beta=0.7)
    assert np.isfinite(out[free]).all() and nmax(out, ref) < 1e-4


def test_values_do_not_depend_on_depth_or_plane_staging(engine):
    """The plan variants the library times -- 4 or 8 sweeps per round, constant planes in shared memory or read from global
    memory -- are schedules of the same per-cell Jacobi update: every bit must agree, for every kind and precision."""
    walls, goal = world((200, 180), 21)
    engine.load_grid(walls, goal, **PARAMS)
    rng = np.random.default_rng(1)
    x0 = np.asfortranarray(rng.normal(size=walls.shape) * 0.5)
    V = np.asfortranarray(rng.normal(size=walls.shape) * 0.3)
    nu = np.asfortranarray(rng.normal(size=walls.shape) * 0.3)
    cases = [("nu", {}, ("f64", "f32", "f32fast")), ("logz", dict(V=V, beta=0.6), ("f64", "f32fast")), ("gap", dict(V=V, nu=nu), ("f32", "f32fast"))]
    try:
        os.environ["PIQL_TILE_TJ"], os.environ["PIQL_TILE_NW"] = "26", "8"
        for kind, kw, precs in cases:
            for prec in precs:
                outs = []
                for kh, pm in (("4", "0"), ("8", "0"), ("4", "1"), ("8", "1")):
                    os.environ["PIQL_TILE_KH"], os.environ["PIQL_TILE_PLANES_MODE"] = kh, pm
                    out, _ = gpu_sweeps(engine, kind, x0, 13, precision=prec, **kw)
                    plan = engine.vi_tile_plan(kind, prec)
                    assert plan["sweeps_per_round"] == int(kh)
                    outs.append(np.nan_to_num(out))
                    assert plan["persistent"] == 1        # 14 tiles: all resident, interiors stay in shared memory, rings exchanged
                # one launch per round (no persistence) computes the same bits ...
                os.environ["PIQL_TILE_NO_PERSIST"] = "1"
                os.environ["PIQL_TILE_KH"], os.environ["PIQL_TILE_PLANES_MODE"] = "4", "0"
                out, _ = gpu_sweeps(engine, kind, x0, 13, precision=prec, **kw)
                assert engine.vi_tile_plan(kind, prec)["persistent"] == 0
                os.environ.pop("PIQL_TILE_NO_PERSIST")
                outs.append(np.nan_to_num(out))
                for o in outs[1:]:
                    assert np.array_equal(outs[0], o), (kind, prec)
                # ... and they are the oracle's 13 Jacobi sweeps
                if prec == "f64":
                    ref, _ = oracle_sweeps(oracle.Grid(walls, goal, **PARAMS), kind, x0, 13, **kw)
                    assert nmax(out, ref) < 1e-12, (kind, nmax(out, ref))
    finally:
        for k in ("PIQL_TILE_TJ", "PIQL_TILE_NW", "PIQL_TILE_KH", "PIQL_TILE_PLANES_MODE", "PIQL_TILE_NO_PERSIST"):
            os.environ.pop(k, None)
