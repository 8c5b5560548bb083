"""The batched entry points (piql_learning_traj, piql_solve_gap_batch): the reference's whole experiment
(run_experiment, scripts/experiments.jl:20-39 -- 8 beta chains of 5 logz_updateV solves, then the optimality gap of all
48 iterates) in two launches, against the Float64 results the reference stored in data/*.jld2.  -m gpu."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def nmax(a, b):
    return float(np.nanmax(np.abs(a - b)))


def test_whole_experiment_reproduces_reference_data(engine, golden):
    g = golden
    engine.load_grid(g.walls, g.goal, **g.params)
    hard = g.params["gamma"] > 0.999
    betas = [float(b) for b in g.betas]
    t0 = time.perf_counter()
    traj, st = engine.learning_traj(np.nan_to_num(g.V(0, 0)), betas, n=5)
    t1 = time.perf_counter()
    assert traj.shape == (8, 6, g.H, g.W) and all(s.converged and s.kernel == 3 for s in st)
    for b, beta in enumerate(betas):
        assert nmax(traj[b, 0], g.V(b, 0)) == 0.0
        for k in range(1, 6):
            # errors of the chain: each update stops at sum|d| < 1e-12 and V + logz/beta amplifies that by 1/beta
            tol = (1e-7 if hard else 1e-9) + 2e-11 / beta
            assert np.array_equal(np.isnan(traj[b, k]), g.walls != 0)
            assert nmax(traj[b, k], g.V(b, k)) < tol, (g.name, b, k, nmax(traj[b, k], g.V(b, k)))
    Vs = np.nan_to_num(traj.reshape(48, g.H, g.W))
    t2 = time.perf_counter()
    gaps, gst = engine.solve_gap_batch(Vs, np.nan_to_num(g.nu))
    t3 = time.perf_counter()
    assert all(s.converged for s in gst)
    gaps = gaps.reshape(8, 6, g.H, g.W)
    for b in range(8):
        for k in range(6):
            tol = 1e-6 if hard else 2e-9
            assert nmax(gaps[b, k], g.G(b, k)) < tol + 1e-3 * float(np.nanmax(np.abs(g.G(b, k)))) * (2e-8 / betas[b] if k else 0), \
                (g.name, b, k, nmax(gaps[b, k], g.G(b, k)))
    print("%s: 40 logz_updateV solves in %.1f ms (%d sweeps), 48 gap solves in %.1f ms (%d sweeps)"
          % (g.name, 1e3 * (t1 - t0), sum(s.sweeps for s in st), 1e3 * (t3 - t2), sum(s.sweeps for s in gst)))


def test_batched_solves_equal_single_solves(engine, golden01):
    g = golden01
    engine.load_grid(g.walls, g.goal, **g.params)
    V0 = np.nan_to_num(g.V(0, 0))
    betas = [float(g.betas[2]), 1.0]
    traj, st = engine.learning_traj(V0, betas, n=2)
    for b, beta in enumerate(betas):
        v = V0
        for k in (1, 2):
            v1, s1 = engine.update_v(np.nan_to_num(v), beta=beta)
            assert s1.sweeps == st[b * 2 + k - 1].sweeps
            assert nmax(traj[b, k], v1) < 1e-13
            v = v1
    nu = np.nan_to_num(g.nu)
    Vs = np.nan_to_num(np.stack([g.V(7, 1), g.V(3, 2)]))
    gaps, gst = engine.solve_gap_batch(Vs, nu)
    for i in range(2):
        one, s1 = engine.solve_gap(Vs[i], nu)
        assert s1.sweeps == gst[i].sweeps and nmax(gaps[i], one) < 1e-13
    # fp32 chains stay within 1e-5 relative of the Float64 ones at beta = 1
    t32, _ = engine.learning_traj(V0, [1.0], n=2, precision="f32", tol=2e-5, max_sweeps=20000, allow_unconverged=True)
    free = ~np.isnan(traj[1, 2])
    rel = np.abs(t32[0, 2][free] - traj[1, 2][free]) / np.maximum(1.0, np.abs(traj[1, 2][free]))
    assert rel.max() < 1e-4, rel.max()


def test_run_experiment_through_the_reference_interface(piql):
    gw = piql.make_gridworld((10, 10), density=0.1, rng=np.random.default_rng(4))
    a = piql.run_experiment(gw, randomness=0.05, gamma=0.95, betarange=[0.05, 1.0], n=2, batched=True)
    b = piql.run_experiment(gw, randomness=0.05, gamma=0.95, betarange=[0.05, 1.0], n=2, batched=False)
    assert set(a["traj"]) == set(b["traj"]) == {0.05, 1.0}
    for beta in a["traj"]:
        for k in range(3):
            assert nmax(a["traj"][beta][k].array, b["traj"][beta][k].array) < 1e-12
            assert nmax(a["gap"][beta][k].array, b["gap"][beta][k].array) < 1e-12
    assert a["s0"] == b["s0"]


def test_monte_carlo_update_tracks_the_exact_update(engine, golden01):
    """piql_mc_update_v: V + log Z^/beta from rollouts approximates logz_updateV (exact log Z) within Monte-Carlo error,
    one iteration and a chain of three; same seed -> same bits."""
    g = golden01
    engine.load_grid(g.walls, g.goal, **g.params)
    V0 = np.nan_to_num(g.V(7, 0))
    free = g.walls == 0
    exact = V0
    for _ in range(3):
        exact, _ = engine.update_v(np.nan_to_num(exact), beta=1.0)
    one_exact, _ = engine.update_v(V0, beta=1.0)
    n = 1 << 24
    one, st = engine.mc_update_v(V0, beta=1.0, npaths=n, niter=1, seed=5)
    assert np.array_equal(np.isnan(one), g.walls != 0) and st.paths == n and st.transitions > n
    err = np.abs(one[free] - one_exact[free])
    assert err.max() < 0.03 and err.mean() < 4e-3, (err.max(), err.mean())
    again, _ = engine.mc_update_v(V0, beta=1.0, npaths=n, niter=1, seed=5)
    assert np.array_equal(one[free], again[free])
    three, st3 = engine.mc_update_v(V0, beta=1.0, npaths=n, niter=3, seed=5)
    assert st3.paths == 3 * n
    err3 = np.abs(three[free] - exact[free])
    assert err3.max() < 0.05 and err3.mean() < 8e-3, (err3.max(), err3.mean())
    # the goal: log Z(goal) = -V(goal) exactly, also by Monte-Carlo (every path from the goal stops at once)
    assert abs(one[g.goal] - one_exact[g.goal]) < 1e-6


def test_mc_update_through_the_reference_interface(piql):
    gw = piql.make_gridworld((10, 10), density=0.1, rng=np.random.default_rng(8))
    ctrl = piql.gridworld_ctrl(gw, randomness=0.05, gamma=0.95)
    weaker = piql.modify(ctrl, reward_function=lambda s, a, ss: 0.1 * ctrl.reward_function(s, a, ss))
    V_lo = piql.generate_nu(weaker)
    exact = piql.logz_updateV(ctrl, V_lo, beta=1.0)
    mc = piql.mc_logz_updateV(ctrl, V_lo, beta=1.0, npaths=1 << 22, seed=2)
    err = np.array([mc(s) - exact(s) for s in exact.dict])
    assert np.abs(err).max() < 0.05 and abs(err.mean()) < 3e-3
