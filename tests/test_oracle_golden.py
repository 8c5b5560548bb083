"""Pins the CPU oracle (oracle/piql_oracle.c) against the reference's own stored results
(data/*.jld2 decoded into tests/golden by tools/make_golden.py) and against the properties the
reference's test/runtests.jl asserts.  CPU only."""
import numpy as np
import pytest

import oracle
from conftest import Golden, random_gridworld


def nmax(a, b):
    return float(np.nanmax(np.abs(a - b)))


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert [hex(x) for x in oracle.philox4x32_10([0] * 4, [0] * 2)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)] == \
        ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                                                 [0xa4093822, 0x299f31d0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_anchor_numbers(golden01):
    # SURVEY App. B / BASELINE.md: the decoder reproduces the anchors verbatim
    g = golden01
    assert g.nu[g.s0] == -1.967650018144423
    assert abs(np.nansum(g.nu) + 409.085791800701) < 1e-9
    assert g.V(7, 5)[g.s0] == -1.9676500173601226
    assert g.G(7, 1)[g.s0] == 0.17038013696194754
    assert g.betas[0] == 0.0024787521766663585 and g.betas[7] == 1.0
    assert (g.walls == 0).sum() == 251 and g.goal == (14, 9) and g.s0 == (1, 16)


def test_nu_and_vlo_match_reference(golden):
    org = golden.oracle()
    nu, st = org.solve("nu")
    assert st.diff < 1e-12
    assert nmax(nu, golden.nu) < (1e-8 if golden.params["gamma"] > 0.999 else 1e-10)
    vlo, _ = org.solve("nu", reward_scale=0.1)   # modify(reward_function = 0.1 * r)
    assert nmax(vlo, golden.V(7, 0)) < (1e-8 if golden.params["gamma"] > 0.999 else 1e-10)
    assert np.array_equal(np.isnan(nu), golden.walls != 0)


@pytest.mark.parametrize("name,b", [("r0_01g0_99", 7), ("r0_01g0_99", 0), ("r0_05g0_95", 7), ("r0_05g0_95", 3),
                                    ("r0_00g0_99", 7)])
def test_learning_trajectory_matches_reference(name, b):
    g = Golden(name)
    org, beta = g.oracle(), float(g.betas[b])
    V = np.nan_to_num(g.V(b, 0))
    for k in range(1, 4):
        V = org.logz_updateV(V, beta=beta)
        # the reference stops at sum|d| < 1e-12, so V = V + logz/beta carries ~1e-12/beta
        assert nmax(V, g.V(b, k)) < 1e-10 + 2e-12 / beta, (k, nmax(V, g.V(b, k)))
        V = np.nan_to_num(V)


def test_long_horizon_update_matches_reference():
    g = Golden("r0_00g0_9999")
    V1 = g.oracle().logz_updateV(np.nan_to_num(g.V(7, 0)), beta=1.0)
    assert nmax(V1, g.V(7, 1)) < 1e-7


@pytest.mark.parametrize("name,b,k", [("r0_05g0_95", 7, 0), ("r0_05g0_95", 0, 1), ("r0_01g0_99", 7, 2)])
def test_optimality_gap_matches_reference(name, b, k):
    g = Golden(name)
    gap, st = g.oracle().solve("gap", V=np.nan_to_num(g.V(b, k)), nu=np.nan_to_num(g.nu), mode="gauss_seidel")
    assert st.diff < 1e-12
    assert nmax(gap, g.G(b, k)) < 1e-9


def test_sweep_orders_agree(golden01):
    org = golden01.oracle()
    a, _ = org.solve("nu", mode="jacobi")
    b, _ = org.solve("nu", mode="gauss_seidel", seed=3)
    c, st = org.solve("nu", mode="faithful", start=int(org.free_cells()[17]), seed=5)
    assert st.nstates == 251            # Growable discovery closes the reachable set
    assert nmax(a, b) < 1e-10 and nmax(a, c) < 1e-10
    V = np.nan_to_num(golden01.V(7, 0))
    la, _ = org.solve("logz", V=V)
    lc, _ = org.solve("logz", V=V, mode="faithful", start=int(org.free_cells()[3]))
    assert nmax(la, lc) < 1e-10
    lz, _ = org.solve("z", V=V, tol=1e-13)  # generate_z: the linear-space variant, Z = exp(logZ)
    assert nmax(np.log(lz), la) < 1e-9


def test_fixed_point_known_answers(golden01):
    # V = nu is a fixed point of logz_updateV: logZ == 0 and gap == 0
    org, nu = golden01.oracle(), np.nan_to_num(golden01.nu)
    logz, _ = org.solve("logz", V=nu, beta=0.3)
    assert float(np.nanmax(np.abs(logz))) < 1e-9
    gap, _ = org.solve("gap", V=nu, nu=nu, mode="gauss_seidel")
    assert float(np.nanmax(np.abs(gap))) < 1e-9


def test_reference_runtests_properties():
    # test/runtests.jl:4-40 on a 10x10 gridworld, density 0.1, randomness 0.1, gamma 0.99
    walls, goal = random_gridworld((10, 10), 0.1, seed=11)
    org = oracle.Grid(walls, goal, randomness=0.1, gamma=0.99)
    assert not walls[goal]
    nu, _ = org.solve("nu")
    lo, _ = org.solve("nu", reward_scale=0.8)
    hi, _ = org.solve("nu", reward_scale=1.2)
    wiggle = 1e-2
    for V in (lo, hi):
        V = np.nan_to_num(V)
        v1 = org.logz_updateV(V, beta=1.0)
        v0 = org.logz_updateV(V, beta=1e-4)
        free = ~np.isnan(nu)
        assert (v1[free] >= nu[free] - wiggle).all()      # upper bound  :33-34
        assert (v0[free] <= nu[free] + wiggle).all()      # lower bound  :35-36
        s = int(org.free_cells()[7])
        if s != org.goal_idx:
            assert org.kl_cost(V, s) > 0                  # :39-40


def test_monte_carlo_estimators_hit_exact_logz(golden01):
    g = golden01
    org = g.oracle()
    V = np.nan_to_num(g.V(7, 0))
    logz, _ = org.solve("logz", V=V, beta=1.0)
    exact = logz[g.s0]
    tab = org.policy_table(V, beta=1.0)
    for est, n, hor in (("A", 200000, 1 << 20), ("B", 20000, 3000)):
        lw, meta = org.rollout(tab, V[g.goal], [g.s0_idx], n, seed=1, estimator=est, horizon=hor, nthreads=4)
        mx, sm, cnt, tr = oracle.reduce_paths(lw, meta, 1)
        lsa, ls = oracle.logz_from_stats(mx, sm, cnt)
        w = np.exp(lw.astype(np.float64))
        sigma = w.std() / np.sqrt(n) / w.mean()
        assert abs(ls[0] - exact) < 4 * sigma, (est, ls[0], exact, sigma)
        assert cnt.sum() == n and tr == (meta >> 2).sum()


def test_table_sampler_follows_the_float64_walk(golden01):
    # the fp32-table rollout takes the same actions as choose_action/propagator evaluated in Float64
    g = golden01
    org = g.oracle()
    V = np.nan_to_num(g.V(7, 1))
    tab = org.policy_table(V, beta=0.5)
    inj = np.random.default_rng(0).random((32, 300, 4)).astype(np.float32)
    lw, meta, trace = org.rollout(tab, V[g.goal], [g.s0_idx], 32, injected=inj, trace_steps=300)
    for p in range(32):
        tr64, lw64, ln = org.rollout_f64(V, g.s0_idx, inj[p].astype(np.float64), beta=0.5)
        L = int(meta[p] >> 2)
        assert np.array_equal(tr64[:L], trace[p, :L])
        assert abs(lw[p] - lw64) < 1e-5 * max(1.0, abs(lw64))


def test_flood_fill_matches_host_generator():
    import piql_jl_b200 as piql
    rng = np.random.default_rng(5)
    raw = piql.make_walls((30, 30), density=0.35, rng=rng)
    goal = (15, 15)
    raw[goal] = False
    walled, nreach = oracle.wall_unreachable(raw, goal[0] + raw.shape[0] * goal[1])
    from piql_jl_b200 import gridworld as gwmod
    mask = gwmod.get_reachable(raw, goal)
    assert np.array_equal(walled != 0, ~mask) and nreach == mask.sum()


def test_path_estimates_restate_the_reference(golden01):
    # fitness_reward_estimate / direct_reward_estimate (path_functions.jl:34-50): fp32 records against the Float64 closures
    g = golden01
    org = g.oracle()
    V = np.nan_to_num(g.V(7, 1))
    tab, kl = org.policy_table(V, beta=1.0), org.kl_table(V)
    fit, dr, tr = org.path_estimates(tab, kl, V, [g.s0_idx], 64, nsteps=20, seed=3, trace=True)
    for p in range(0, 64, 9):
        f64, d64 = org.path_estimates_f64(V, tr[p])
        assert np.abs(fit[p] - f64).max() < 2e-6 and np.abs(dr[p] - d64).max() < 2e-6
    assert (tr[:, 1:, 0] == tr[:, :-1, 2]).all() and (tr[:, 0, 0] == g.s0_idx).all()
    # V = nu: fitness of the V-determined policy vanishes, the fitness total never moves
    nu = np.nan_to_num(g.nu)
    fit, _ = org.path_estimates(org.policy_table(nu, beta=1.0), org.kl_table(nu), nu, [g.s0_idx], 256, nsteps=20, seed=4)
    assert np.abs(fit - np.float32(nu[g.s0])).max() < 2e-5


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_restatements_agree_on_random_gridworlds(seed):
    """The oracle's independent restatements of one recursion -- the faithful one (hash-map Growable that discovers the
    states from a start cell, shuffled Gauss-Seidel, closure-style nested recomputation) and the dense Jacobi one that
    the stored data/*.jld2 pin -- agree on random small problems over the whole parameter range, including goals that
    touch walls, randomness 0 and 1 and short horizons."""
    from conftest import random_gridworld
    rng = np.random.default_rng(100 + seed)
    size = (int(rng.integers(3, 9)), int(rng.integers(3, 9)))
    walls, goal = random_gridworld(size, float(rng.choice([0.0, 0.15, 0.3])), seed=200 + seed)
    params = dict(randomness=float(rng.choice([0.0, 0.03, 0.5, 1.0])), gamma=float(rng.choice([0.5, 0.9, 0.99])),
                  reward_scale=float(rng.choice([0.2, 1.0])))
    org = oracle.Grid(walls, goal, **params)
    free = org.free_cells()
    start = int(free[int(rng.integers(len(free)))])
    nu_j, _ = org.solve("nu", mode="jacobi")
    nu_f, st = org.solve("nu", mode="faithful", start=start, seed=seed)
    assert st.nstates == len(free)                      # every free cell reaches the goal, so discovery finds them all
    assert np.array_equal(np.isnan(nu_j), walls != 0) and nmax(nu_j, nu_f) < 1e-9
    assert nu_j[goal] == 0.0
    V = np.nan_to_num(org.solve("nu", reward_scale=0.3)[0])
    beta = float(rng.choice([0.05, 0.5, 1.0]))
    lz_j, _ = org.solve("logz", V=V, beta=beta)
    lz_f, _ = org.solve("logz", V=V, beta=beta, mode="faithful", start=start, seed=seed + 1)
    assert nmax(lz_j, lz_f) < 1e-9 and abs(lz_j[goal] + V[goal]) < 1e-11
    gap_j, _ = org.solve("gap", V=V, nu=np.nan_to_num(nu_j))
    gap_g, _ = org.solve("gap", V=V, nu=np.nan_to_num(nu_j), mode="gauss_seidel", seed=seed)
    assert nmax(gap_j, gap_g) < 1e-8 and np.nanmin(gap_j) > -1e-9   # nu is the optimum: the gap of any policy is >= 0
    # one backup of the converged iterate is the iterate (the sweep and the solver share no loop)
    again, diff = org.sweep("nu", np.nan_to_num(nu_j))
    assert diff < 1e-10 and nmax(again, nu_j) < 1e-11
