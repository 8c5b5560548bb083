"""The column-resident schedule of the soft-VI sweeps (col_resident_kernel, csrc/tile.cu) against the CPU oracle and against the
temporally blocked schedule, through the C ABI.  -m gpu.

One CTA per SM keeps a slab of whole columns in shared memory for a whole batch of sweeps and exchanges one column per sweep
with each neighbour through L2 mailboxes ({word, sequence number} pairs).  What a bug there would break: the values after N
sweeps (N = 1: no exchange at all; N = 2: first mailbox round; odd / even N: both parities), the sequence numbers across
launches (many short calls in a row), the goal patch when the goal or one of its neighbours sits in a slab's edge column, the
strip seams of a column (grids taller than 128 rows), the residual.  Every bit must agree with the tiled schedule: both run
strip_compute on the same Jacobi iterate.
"""
import os

import numpy as np
import pytest

import oracle
from conftest import random_gridworld

pytestmark = pytest.mark.gpu

PARAMS = dict(randomness=0.05, step_cost=0.02, reward=0.2, gamma=0.97)
ENV = ("PIQL_TILE_COLRES", "PIQL_TILE_NW", "PIQL_TILE_PLANES_MODE", "PIQL_TILE_TJ", "PIQL_TILE_KH")


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


def gpu_sweeps(engine, kind, x, n, precision="f64", calls=1, **kw):
    engine.vi_prepare(kind, precision=precision, **kw)
    engine.vi_set_iterate(x)
    for _ in range(calls):
        engine.vi_sweeps(n)
    return engine.vi_fetch()


@pytest.fixture
def pins():
    saved = {k: os.environ.get(k) for k in ENV}
    yield os.environ
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def forced(engine, kind, prec):
    plan = engine.vi_tile_plan(kind, prec)
    assert plan["column_resident"] == 1, plan
    return plan


@pytest.mark.parametrize("shape,seed,goal", [((150, 70), 11, None), ((100, 100), 5, (4, 50)), ((260, 40), 12, (128, 20)),
                                              ((131, 329), 13, (120, 9)), ((40, 300), 14, (20, 33)), ((300, 200), 15, (129, 2))])
def test_column_resident_sweeps_equal_jacobi_sweeps(engine, pins, shape, seed, goal):
    walls, goal = world(shape, seed, goal)
    engine.load_grid(walls, goal, **PARAMS)
    org = oracle.Grid(walls, goal, **PARAMS)
    rng = np.random.default_rng(seed)
    x0 = np.asfortranarray(rng.normal(size=walls.shape) * 0.5)
    V = np.asfortranarray(rng.normal(size=walls.shape) * 0.3)
    nu = np.asfortranarray(rng.normal(size=walls.shape) * 0.3)
    Q = rng.normal(size=(walls.size, 4)) * 0.2
    pins["PIQL_TILE_COLRES"] = "1"
    cases = [("nu", dict(reward_scale=0.7)), ("logz", dict(V=V, beta=0.4)), ("logz", dict(V=V, Q=Q, beta=1.0)),
             ("gap", dict(V=V, nu=nu)), ("gap", dict(V=V, nu=nu, Q=Q))]
    for kind, kw in cases:
        for n in (1, 2, 7):
            ref, dref = oracle_sweeps(org, kind, x0, n, **kw)
            out, d = gpu_sweeps(engine, kind, x0, n, **kw)
            forced(engine, kind, "f64")
            assert np.array_equal(np.isnan(out), walls != 0)
            assert nmax(out, ref) < 1e-12, (kind, n, nmax(out, ref))
            assert abs(d - dref) < 1e-9 * max(1.0, dref), (kind, n, d, dref)


def test_column_resident_bits_equal_the_tiled_schedule(engine, pins):
    """All kinds and precisions, several warps per CTA, planes in shared memory or global memory, one long call or many short ones
    (the mailbox sequence numbers run on from launch to launch): the same bits as one launch per round of the tiled kernel."""
    walls, goal = world((200, 180), 21)
    engine.load_grid(walls, goal, **PARAMS)
    rng = np.random.default_rng(1)
    x0 = np.asfortranarray(rng.normal(size=walls.shape) * 0.5)
    V = np.asfortranarray(rng.normal(size=walls.shape) * 0.3)
    nu = np.asfortranarray(rng.normal(size=walls.shape) * 0.3)
    cases = [("nu", {}, ("f64", "f32", "f32fast")), ("logz", dict(V=V, beta=0.6), ("f64", "f32", "f32fast")),
             ("gap", dict(V=V, nu=nu), ("f64", "f32", "f32fast"))]
    for kind, kw, precs in cases:
        for prec in precs:
            pins["PIQL_TILE_COLRES"] = "0"
            pins["PIQL_TILE_TJ"], pins["PIQL_TILE_NW"], pins["PIQL_TILE_KH"] = "26", "8", "4"
            ref, dref = gpu_sweeps(engine, kind, x0, 12, precision=prec, **kw)
            assert engine.vi_tile_plan(kind, prec)["column_resident"] == 0
            pins.pop("PIQL_TILE_TJ"); pins.pop("PIQL_TILE_KH")
            pins["PIQL_TILE_COLRES"] = "1"
            for nw in ("8", "16") + (() if prec == "f64" else ("20", "28")):
                for pm in ("0", "1"):
                    pins["PIQL_TILE_NW"], pins["PIQL_TILE_PLANES_MODE"] = nw, pm
                    for n, calls in ((12, 1), (3, 4), (1, 12)):
                        out, d = gpu_sweeps(engine, kind, x0, n, precision=prec, calls=calls, **kw)
                        plan = forced(engine, kind, prec)
                        assert plan["warps"] == int(nw)
                        assert np.array_equal(np.nan_to_num(out).view(np.uint64), np.nan_to_num(ref).view(np.uint64)), (kind, prec, nw, pm, n)
                        assert d == dref or abs(d - dref) <= 1e-12 * max(1.0, abs(dref)), (kind, prec, d, dref)
            pins.pop("PIQL_TILE_NW"); pins.pop("PIQL_TILE_PLANES_MODE")


def test_column_resident_solve_matches_oracle_and_is_chosen_when_it_wins(engine, pins):
    """A whole solve to tolerance through the public call; the tuner may pick any schedule, a forced column-resident solve takes
    the same number of sweeps and returns the same bits."""
    walls, goal = world((140, 110), 31)
    p = dict(randomness=0.02, step_cost=0.02, reward=0.2, gamma=0.95)
    engine.load_grid(walls, goal, **p)
    org = oracle.Grid(walls, goal, **p)
    nu0, st0 = engine.solve_nu(tol=1e-11, check_every=32, max_sweeps=100000)
    pins["PIQL_TILE_COLRES"] = "1"
    nu1, st1 = engine.solve_nu(tol=1e-11, check_every=32, max_sweeps=100000)
    forced(engine, "nu", "f64")
    assert st1.converged and st1.sweeps == st0.sweeps
    assert np.array_equal(np.nan_to_num(nu0).view(np.uint64), np.nan_to_num(nu1).view(np.uint64))
    ref, _ = org.solve("nu", tol=1e-11, nthreads=8)
    assert nmax(nu1, ref) < 1e-9
    V = np.nan_to_num(org.solve("nu", reward_scale=0.1, tol=1e-11, nthreads=8)[0])
    v1, s1 = engine.update_v(V, beta=0.5, tol=1e-10, max_sweeps=100000)
    assert s1.converged and nmax(v1, org.logz_updateV(V, beta=0.5, tol=1e-10, nthreads=8)) < 1e-8
    l32, _ = engine.solve_logz(V, beta=1.0, precision="f32fast", tol=1e-4, max_sweeps=8000, allow_unconverged=True)
    l64, _ = engine.solve_logz(V, beta=1.0, tol=1e-10, max_sweeps=100000)
    free = ~np.isnan(l64)
    assert float(np.max(np.abs(l32[free] - l64[free]) / np.maximum(1.0, np.abs(l64[free])))) < 1e-5
