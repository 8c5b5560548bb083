"""Parity of the CUDA path with the CPU oracle, through the C ABI.  Needs a B200: -m gpu.

Bars: bit-exact for integer work (trajectories, first actions, path lengths, counts) and for the
fp32 log-weight of estimator A (same additions in the same order); 1e-9 absolute for Float64
value iteration against the reference's stored results (SURVEY App. B tolerance note); 1e-5
relative for fp32 sweeps and for log-sum-exp reductions.
"""
import numpy as np
import pytest

import oracle
from conftest import Golden, random_gridworld

pytestmark = pytest.mark.gpu


def nmax(a, b):
    return float(np.nanmax(np.abs(a - b)))


def load(engine, g):
    engine.load_grid(g.walls, g.goal, **g.params)
    return g.oracle()


# ---- single backups -------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol", [("f64", 1e-13), ("f32", 2e-6)])
def test_single_sweep_matches_oracle(engine, golden01, precision, tol):
    g = golden01
    org = load(engine, g)
    rng = np.random.default_rng(0)
    x = np.asfortranarray(rng.normal(size=(g.H, g.W)) * 0.5)
    V = np.nan_to_num(g.V(7, 1))
    nu = np.nan_to_num(g.nu)
    for kind, kw in (("nu", dict(reward_scale=0.7)), ("logz", dict(V=V, beta=0.4)), ("gap", dict(V=V, nu=nu))):
        ref, dref = org.sweep(kind, x, **kw)
        out, d = engine.sweep_once(kind, x, precision=precision, **kw)
        assert np.array_equal(np.isnan(out), g.walls != 0)
        assert nmax(out, ref) < tol, (kind, nmax(out, ref))
        assert abs(d - dref) < tol * 400, (kind, d, dref)


def test_single_sweep_with_q_table(engine, golden01):
    g = golden01
    org = load(engine, g)
    rng = np.random.default_rng(1)
    x = np.asfortranarray(rng.normal(size=(g.H, g.W)) * 0.3)
    V = np.nan_to_num(g.V(5, 2))
    Q = rng.normal(size=(g.H * g.W, 4)) * 0.2
    for kind, kw in (("logz", dict(V=V, Q=Q, beta=0.25)), ("gap", dict(V=V, Q=Q, nu=np.nan_to_num(g.nu)))):
        ref, _ = org.sweep(kind, x, **kw)
        out, _ = engine.sweep_once(kind, x, **kw)
        assert nmax(out, ref) < 1e-13


def test_single_sweep_large_grid_generic_kernel(engine):
    walls, goal = random_gridworld((120, 90), 0.2, seed=2)
    engine.load_grid(walls, goal, 0.05, 0.02, 0.2, 0.98)
    org = oracle.Grid(walls, goal, randomness=0.05, step_cost=0.02, reward=0.2, gamma=0.98)
    x = np.asfortranarray(np.random.default_rng(3).normal(size=walls.shape))
    ref, dref = org.sweep("nu", x)
    out, d = engine.sweep_once("nu", x)
    assert nmax(out, ref) < 1e-13 and abs(d - dref) < 1e-9


# ---- the linear-space twin: update_z! / generate_z / z_updateV (value_iteration.jl:23-59) -----------------------
@pytest.mark.parametrize("precision,tol", [("f64", 1e-13), ("f32", 2e-6)])
def test_linear_space_z_sweep_matches_oracle(engine, golden01, precision, tol):
    g = golden01
    org = load(engine, g)
    x = np.asfortranarray(np.random.default_rng(4).uniform(0.2, 2.0, size=(g.H, g.W)))   # Z > 0
    V = np.nan_to_num(g.V(7, 1))
    ref, dref = org.sweep("z", x, V=V, beta=0.4)
    out, d = engine.sweep_once("z", x, V=V, beta=0.4, precision=precision)
    assert np.array_equal(np.isnan(out), g.walls != 0)
    assert nmax(out, ref) < tol and abs(d - dref) < tol * 4000, (nmax(out, ref), d, dref)
    Q = np.random.default_rng(5).normal(size=(g.H * g.W, 4)) * 0.2
    if precision == "f64":
        assert nmax(engine.sweep_once("z", x, V=V, Q=Q, beta=0.25)[0], org.sweep("z", x, V=V, Q=Q, beta=0.25)[0]) < 1e-13


def test_generate_z_and_z_updateV(engine, golden01, piql):
    g = golden01
    org = load(engine, g)
    V = np.nan_to_num(g.V(7, 0))
    for beta in (1.0, 0.1800923121479524):
        z, st = engine.solve_z(V, beta=beta)                       # the reference's stop rule: 1e-9 on sum |dlog Z|
        ref, rst = org.solve("z", V=V, beta=beta, tol=1e-9)
        assert st.kernel == 1 and st.converged and st.sweeps == rst.sweeps
        assert nmax(z, ref) < 1e-12
        logz, _ = engine.solve_logz(V, beta=beta)
        assert nmax(np.log(z), logz) < 2e-9                        # the two forms of the same fixed point
        v1, _ = engine.z_update_v(V, beta=beta)
        free = g.walls == 0
        assert nmax(v1[free], (V + np.log(ref) / beta)[free]) < 1e-10
        assert nmax(v1, engine.update_v(V, beta=beta)[0]) < 2e-8   # z_updateV == logz_updateV up to the stop rules
    # through the mirror of the reference interface
    gw = piql.Gridworld(g.walls != 0, g.goal)
    ctrl = piql.gridworld_ctrl(gw, randomness=g.params["randomness"], gamma=g.params["gamma"])
    Vg = piql.Growable.from_dense(V, g.walls)
    Z = piql.generate_z(ctrl, Vg, β=1.0)
    assert Z.default(None) == 1.0 and abs(Z(g.s0) - ref_at(org, V, g)) < 1e-12
    V1 = piql.z_updateV(ctrl, Vg, β=1.0)
    assert abs(V1(g.s0) - (Vg(g.s0) + np.log(Z(g.s0)))) < 1e-12


def ref_at(org, V, g):
    z, _ = org.solve("z", V=V, beta=1.0, tol=1e-9)
    return z[g.s0]


def test_linear_space_z_on_a_large_grid(engine):
    # > RESIDENT_MAX_CELLS: one launch per sweep (stats.kernel 5), stop rule checked every 32 sweeps
    walls, goal = random_gridworld((100, 100), 0.2, seed=5)
    engine.load_grid(walls, goal, 0.02, 0.02, 0.2, 0.95)
    org = oracle.Grid(walls, goal, randomness=0.02, step_cost=0.02, reward=0.2, gamma=0.95)
    V = np.nan_to_num(org.solve("nu", reward_scale=0.1, tol=1e-11, nthreads=8)[0])
    z, st = engine.solve_z(V, beta=0.5, check_every=32)
    assert st.kernel == 5 and st.converged and st.sweeps % 32 == 0
    ref, _ = org.solve("z", V=V, beta=0.5, tol=1e-9, nthreads=8)
    assert nmax(np.log(z), np.log(ref)) < 1e-8
    with pytest.raises(Exception):
        engine.vi_prepare("z", V=V)


# ---- fixed points against the reference's stored results ---------------------------------------
def test_nu_and_vlo_match_reference_data(engine, golden):
    g = golden
    load(engine, g)
    tol = 1e-7 if g.params["gamma"] > 0.999 else 1e-9
    nu, st = engine.solve_nu()
    assert st.converged and st.kernel == 1 and st.residual < 1e-12
    assert nmax(nu, g.nu) < tol
    assert np.array_equal(np.isnan(nu), g.walls != 0)
    vlo, _ = engine.solve_nu(reward_scale=0.1)
    assert nmax(vlo, g.V(7, 0)) < tol


def test_learning_trajectories_match_reference_data(engine, golden):
    g = golden
    load(engine, g)
    scale = 1e-7 if g.params["gamma"] > 0.999 else 1e-9
    for b in (0, 3, 7):
        beta = float(g.betas[b])
        for k in range(1, 6):
            # V_k = logz_updateV(V_{k-1}) from the stored V_{k-1}; both sides stop at sum|d| < 1e-12, which
            # V + logz/beta amplifies by 1/beta (SURVEY App. B tolerance note)
            V, st = engine.update_v(np.nan_to_num(g.V(b, k - 1)), beta=beta)
            assert st.converged
            assert nmax(V, g.V(b, k)) < scale + 5e-12 / beta, (b, k, nmax(V, g.V(b, k)))


def test_optimality_gaps_match_reference_data(engine, golden):
    g = golden
    load(engine, g)
    nu = np.nan_to_num(g.nu)
    for b, k in ((7, 0), (7, 1), (0, 1), (4, 2), (7, 5)):
        gap, st = engine.solve_gap(np.nan_to_num(g.V(b, k)), nu)
        assert st.converged
        assert nmax(gap, g.G(b, k)) < (1e-6 if g.params["gamma"] > 0.999 else 1e-9), (b, k, nmax(gap, g.G(b, k)))


def test_fixed_point_known_answers(engine, golden01):
    g = golden01
    load(engine, g)
    nu = np.nan_to_num(g.nu)
    logz, _ = engine.solve_logz(nu, beta=0.3)
    assert float(np.nanmax(np.abs(logz))) < 1e-9
    gap, _ = engine.solve_gap(nu, nu)
    assert float(np.nanmax(np.abs(gap))) < 1e-9


def test_reference_runtests_properties(engine, piql):
    # test/runtests.jl:4-36 through the reference-interface mirror
    gw = piql.make_gridworld((10, 10), density=0.1, rng=np.random.default_rng(11))
    ctrl = piql.gridworld_ctrl(gw, randomness=0.1, gamma=0.99, rng=np.random.default_rng(12))
    weaker = piql.modify(ctrl, reward_function=lambda s, a, ss: 0.8 * ctrl.reward_function(s, a, ss))
    stronger = piql.modify(ctrl, reward_function=lambda s, a, ss: 1.2 * ctrl.reward_function(s, a, ss))
    nu, V_lo, V_hi = piql.generate_nu(ctrl), piql.generate_nu(weaker), piql.generate_nu(stronger)
    vhi_1, vlo_1 = piql.logz_updateV(ctrl, V_hi, beta=1.0), piql.logz_updateV(ctrl, V_lo, beta=1.0)
    vhi_0, vlo_0 = piql.logz_updateV(ctrl, V_hi, beta=1e-4), piql.logz_updateV(ctrl, V_lo, beta=1e-4)
    wiggle = 1e-2
    assert all(vlo_1(k) >= nu(k) - wiggle for k in nu.dict)
    assert all(vhi_1(k) >= nu(k) - wiggle for k in nu.dict)
    assert all(vlo_0(k) <= nu(k) + wiggle for k in nu.dict)
    assert all(vhi_0(k) <= nu(k) + wiggle for k in nu.dict)
    s = ctrl.initial_state()
    if s != tuple(gw.goal):
        assert 0 < piql.kl_cost(ctrl, s, lambda s, a: piql.controlV(ctrl, V_lo, s, a))
    # and the whole problem agrees with the oracle
    org = oracle.Grid(np.asfortranarray(gw.walls, dtype=np.uint8), gw.goal, randomness=0.1, gamma=0.99)
    assert nmax(nu.array, org.solve("nu")[0]) < 1e-10
    assert nmax(vlo_1.array, org.logz_updateV(np.nan_to_num(V_lo.array), beta=1.0)) < 1e-10


def test_generic_kernel_solve_matches_oracle(engine):
    # > RESIDENT_MAX_CELLS cells: tiled multi-sweep kernel with the deterministic residual
    walls, goal = random_gridworld((100, 100), 0.2, seed=5)
    engine.load_grid(walls, goal, 0.02, 0.02, 0.2, 0.95)
    org = oracle.Grid(walls, goal, randomness=0.02, step_cost=0.02, reward=0.2, gamma=0.95)
    nu, st = engine.solve_nu(tol=1e-11, check_every=32)
    assert st.kernel == 2 and st.converged
    ref, _ = org.solve("nu", tol=1e-11, nthreads=8)
    assert nmax(nu, ref) < 1e-9
    V = np.nan_to_num(org.solve("nu", reward_scale=0.1, tol=1e-11, nthreads=8)[0])
    v1, st1 = engine.update_v(V, beta=0.5, tol=1e-10)
    assert st1.converged and nmax(v1, org.logz_updateV(V, beta=0.5, tol=1e-10, nthreads=8)) < 1e-8
    # bit-stable: the same solve twice gives identical bits and sweep counts
    nu2, st2 = engine.solve_nu(tol=1e-11, check_every=32)
    assert st2.sweeps == st.sweeps
    assert np.array_equal(nu.view(np.uint64), nu2.view(np.uint64))


def test_fp32_sweeps_within_1e5_relative(engine, golden01):
    g = golden01
    load(engine, g)
    nu64, _ = engine.solve_nu()
    nu32, st = engine.solve_nu(precision="f32", tol=2e-5, max_sweeps=20000, allow_unconverged=True)
    free = ~np.isnan(nu64)
    rel = np.abs(nu32[free] - nu64[free]) / np.maximum(1.0, np.abs(nu64[free]))
    assert rel.max() < 1e-5 * 10, rel.max()   # contraction 1/(1-gamma) = 100 amplifies fp32 rounding
    V = np.nan_to_num(g.V(7, 0))
    l64, _ = engine.solve_logz(V, beta=1.0)
    l32, _ = engine.solve_logz(V, beta=1.0, precision="f32", tol=2e-5, max_sweeps=20000, allow_unconverged=True)
    rel = np.abs(l32[free] - l64[free]) / np.maximum(1.0, np.abs(l64[free]))
    assert rel.max() < 1e-4, rel.max()


# ---- the reference's own path statistics: fitness_reward_estimate / direct_reward_estimate --------------------
@pytest.mark.parametrize("injected", [False, True])
def test_path_estimates_bit_exact(engine, golden, injected):
    g = golden
    org = load(engine, g)
    V = np.nan_to_num(g.V(5, 2))
    npaths, nsteps = 3000, 20                                         # path_transitions' nsteps = 20
    starts = [g.s0_idx, g.goal[0] + g.H * g.goal[1], int(org.free_cells()[7])]   # also a path that starts on the goal
    inj = np.random.default_rng(8).random((npaths, nsteps, 4), dtype=np.float32) if injected else None
    fit, dr, tr, tab, kl = engine.path_estimates(V, npaths, nsteps=nsteps, want_trace=True, want_tables=True,
                                                 seed=11, path0=5, starts=starts, injected=inj)
    # what the kernel read: the beta = 1 records and kl_cost per cell, against the oracle's Float64 evaluation
    ref_tab, ref_kl = org.policy_table(V, beta=1.0), org.kl_table(V)
    ulp = np.abs(tab.view(np.int32).astype(np.int64) - ref_tab.view(np.int32).astype(np.int64))
    ulp[:, 3] = 0
    assert ulp[np.abs(ref_tab) > 1e-30].max() <= 2 and np.abs(kl - ref_kl).max() < 1e-7
    # same records, same uniforms: identical trajectories and identical fp32 running totals
    rfit, rdr, rtr = org.path_estimates(tab, kl, V, starts, npaths, nsteps=nsteps, path0=5, seed=11, injected=inj, trace=True)
    assert np.array_equal(tr, rtr)
    assert np.array_equal(fit.view(np.uint32), rfit.view(np.uint32))
    assert np.array_equal(dr.view(np.uint32), rdr.view(np.uint32))
    # ... and they are the Float64 estimates the reference's closures give along those paths
    for p in (0, 1, 2, 17, npaths - 1):
        f64, d64 = org.path_estimates_f64(V, tr[p])
        assert np.abs(fit[p] - f64).max() < 1e-5 and np.abs(dr[p] - d64).max() < 1e-5
    # the walk goes on through the goal (no terminal handling in path_transitions)
    assert tr[2, 0, 0] == starts[(5 + 2) % 3] == starts[1] and (tr[:, 1:, 0] == tr[:, :-1, 2]).all()


def test_path_estimates_known_answers(engine, golden01, piql):
    g = golden01
    load(engine, g)
    nu = np.nan_to_num(g.nu)
    # at V = nu the V-determined policy is optimal and fitness vanishes identically: the fitness total stays at nu(s0)
    fit, dr = engine.path_estimates(nu, 4096, nsteps=20, seed=1, starts=[g.s0_idx])
    assert np.abs(fit - np.float32(nu[g.s0])).max() < 2e-5
    # ... while the direct (sampled-reward) totals are unbiased for the same quantity only through gamma-discounting,
    # which this undiscounted 20-step walk does not apply: they drift down by about step_cost + KL per step
    assert dr[:, -1].mean() < dr[:, 0].mean() < nu[g.s0] + 1e-3
    # the host-side generator mirrors give the same totals along the same path
    gw = piql.Gridworld(g.walls != 0, g.goal)
    ctrl = piql.gridworld_ctrl(gw, randomness=g.params["randomness"], gamma=g.params["gamma"])
    V = piql.Growable.from_dense(np.nan_to_num(g.V(7, 1)), g.walls)
    f2, d2, tr = engine.path_estimates(np.nan_to_num(g.V(7, 1)), 4, nsteps=20, seed=2, starts=[g.s0_idx], want_trace=True)
    Q = lambda s, a: piql.controlV(ctrl, V, s, a)
    tf = td = V(g.s0)
    for t in range(20):
        s, a, ss = (int(tr[0, t, 0]) % g.H, int(tr[0, t, 0]) // g.H), int(tr[0, t, 1]), (int(tr[0, t, 2]) % g.H, int(tr[0, t, 2]) // g.H)
        tf += piql.fitness(ctrl, V, Q, s, a)
        td += ctrl.reward_function(s, a, ss) - piql.kl_cost(ctrl, s, Q)
        assert abs(tf - f2[0, t]) < 1e-5 and abs(td - d2[0, t]) < 1e-5
    fit3, dr3 = piql.path_estimates(ctrl, V, s0=g.s0, npaths=4, seed=2)
    assert np.array_equal(fit3, f2) and np.array_equal(dr3, d2)


# ---- policy table and trajectories -----------------------------------------------------------------
def test_policy_table_matches_oracle(engine, golden01):
    g = golden01
    org = load(engine, g)
    V = np.nan_to_num(g.V(6, 1))
    for Q in (None, np.random.default_rng(2).normal(size=(g.H * g.W, 4)) * 0.1):
        tab = engine.policy_table(V, beta=0.42, Q=Q)
        ref = org.policy_table(V, beta=0.42, Q=Q)
        assert np.array_equal(tab[:, 3].view(np.uint32), ref[:, 3].view(np.uint32))      # wall masks
        ulp = np.abs(tab.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
        ulp[:, 3] = 0
        small = np.abs(ref) < 1e-30
        assert (ulp[~small] <= 1).all(), ulp.max()
        assert np.abs(tab - ref)[small].max(initial=0.0) < 1e-12


@pytest.mark.parametrize("estimator", ["A", "B"])
def test_trajectories_bit_exact_with_injected_uniforms(engine, golden, estimator):
    g = golden
    org = load(engine, g)
    V = np.nan_to_num(g.V(7, 0))
    table = engine.policy_table(V, beta=1.0)
    n, steps = 96, 600
    inj = np.random.default_rng(4).random((n, steps, 4)).astype(np.float32)
    starts = [g.s0_idx, int(org.free_cells()[40]), g.goal[0] + g.H * g.goal[1]]
    tr, logw, meta = engine.trace_paths(n, steps, starts=starts, injected=inj, estimator=estimator, horizon=500)
    rlw, rmeta, rtr = org.rollout(table, V[g.goal], starts, n, injected=inj, estimator=estimator, horizon=500,
                                  trace_steps=steps)
    assert np.array_equal(tr, rtr)
    assert np.array_equal(meta, rmeta)
    if estimator == "A":
        assert np.array_equal(logw.view(np.uint32), rlw.view(np.uint32))
    else:
        assert np.allclose(logw, rlw, rtol=1e-5, atol=1e-6)
    assert (meta[2::3] >> 2 == 0).all()    # paths that start at the goal have no transitions


def test_philox_trajectories_bit_exact(engine, golden01):
    g = golden01
    org = load(engine, g)
    V = np.nan_to_num(g.V(7, 0))
    table = engine.policy_table(V, beta=1.0)
    n = 20000
    free = org.free_cells()
    for path0, seed in ((0, 0), (123456789012, 0xDEADBEEFCAFE)):
        tr, logw, meta = engine.trace_paths(256, 200, path0=path0, seed=seed)
        rlw, rmeta, rtr = org.rollout(table, V[g.goal], free, 256, path0=path0, seed=seed, trace_steps=200)
        assert np.array_equal(tr, rtr) and np.array_equal(meta, rmeta)
        engine.mc_reset()
        engine.mc_rollout(n, path0=path0, seed=seed)
        logw, meta = engine.mc_records(n)
        rlw, rmeta = org.rollout(table, V[g.goal], free, n, path0=path0, seed=seed, nthreads=4)
        assert np.array_equal(meta, rmeta)
        assert np.array_equal(logw.view(np.uint32), rlw.view(np.uint32))


def test_table_set_round_trip(engine, golden01):
    g = golden01
    org = load(engine, g)
    V = np.nan_to_num(g.V(3, 2))
    ref = org.policy_table(V, beta=0.7)
    engine.policy_table_set(ref, V[g.goal])
    engine.mc_reset(starts=[g.s0_idx])
    engine.mc_rollout(5000, seed=9, starts=[g.s0_idx], estimator="B", horizon=400)
    logw, meta = engine.mc_records(5000)
    rlw, rmeta = org.rollout(ref, V[g.goal], [g.s0_idx], 5000, seed=9, estimator="B", horizon=400)
    assert np.array_equal(meta, rmeta) and np.allclose(logw, rlw, rtol=1e-5, atol=1e-6)


# ---- reduction -------------------------------------------------------------------------------------
@pytest.mark.parametrize("nstarts_kind", ["one", "few", "all"])
def test_reduction_matches_oracle(engine, golden01, nstarts_kind):
    g = golden01
    org = load(engine, g)
    V = np.nan_to_num(g.V(7, 0))
    engine.policy_table(V, beta=1.0, download=False)
    free = org.free_cells()
    starts = {"one": [g.s0_idx], "few": [int(c) for c in free[::50]], "all": None}[nstarts_kind]
    slot_cells = free if starts is None else np.asarray(starts)
    n, path0 = 300007, 1000003
    engine.mc_reset(starts=starts)
    engine.mc_rollout(n, path0=path0, seed=5, starts=starts)
    logw, meta = engine.mc_records(n)
    sa, cnt, ls, st = engine.mc_finalize()
    mx, sm, c, tr = oracle.reduce_paths(logw, meta, len(slot_cells), path0=path0)
    lsa_ref, ls_ref = oracle.logz_from_stats(mx, sm, c)
    assert st.transitions == tr and st.paths == n
    assert np.array_equal(cnt[slot_cells], c)
    other = np.setdiff1d(np.arange(g.H * g.W), slot_cells)
    assert (cnt[other] == 0).all() and np.isnan(sa[other]).all()
    ok = c > 0
    assert np.allclose(sa[slot_cells][ok], lsa_ref[ok], rtol=1e-5, atol=1e-6)
    assert np.allclose(ls.ravel(order="F")[slot_cells], ls_ref, rtol=1e-5, atol=1e-6)


def test_reduction_accumulates_across_calls_and_shards(engine, golden01):
    # two shards of the path-id range, reduced separately then merged == one call
    g = golden01
    org = load(engine, g)
    V = np.nan_to_num(g.V(7, 0))
    engine.policy_table(V, beta=1.0, download=False)
    n = 100000
    engine.mc_reset()
    engine.mc_rollout(n, seed=3)
    sa1, cnt1, ls1, st1 = engine.mc_finalize()
    engine.mc_reset()
    engine.mc_rollout(n // 4, path0=0, seed=3)
    engine.mc_rollout(n - n // 4, path0=n // 4, seed=3)
    sa2, cnt2, ls2, st2 = engine.mc_finalize()
    assert np.array_equal(cnt1, cnt2) and st1.transitions == st2.transitions
    assert np.allclose(sa1, sa2, rtol=1e-6, atol=1e-6, equal_nan=True)


# ---- the estimator -----------------------------------------------------------------------------------
@pytest.mark.parametrize("name,estimator,n", [("r0_00g0_99", "A", 1 << 22), ("r0_01g0_99", "B", 1 << 17),
                                              ("r0_00g0_9999", "A", 1 << 19)])
def test_mc_logz_hits_exact_value(engine, piql, name, estimator, n):
    g = Golden(name)
    load(engine, g)
    for b, V in ((7, np.nan_to_num(g.V(7, 0))), (5, np.nan_to_num(g.nu))):
        beta = float(g.betas[b])
        exact, _ = engine.solve_logz(V, beta=beta)
        hor = 3000 if estimator == "B" else 1 << 20
        sa, cnt, ls, st = engine.mc_logz(V, beta=beta, npaths=n, seed=b, estimator=estimator, horizon=hor,
                                         starts=[g.s0_idx])
        logw, _ = engine.mc_records(n)
        w = np.exp(logw.astype(np.float64) - logw.max())
        sigma = w.std() / np.sqrt(n) / w.mean()
        est = ls[g.s0]
        assert abs(est - exact[g.s0]) < 4 * sigma + 2e-6, (name, b, est, exact[g.s0], sigma)
        assert cnt.sum() == n and st.transitions > 0


def test_mc_result_through_reference_interface(engine, piql):
    gw = piql.make_gridworld((12, 12), density=0.15, rng=np.random.default_rng(21))
    ctrl = piql.gridworld_ctrl(gw, randomness=0.05, gamma=0.95)
    weaker = piql.modify(ctrl, reward_function=lambda s, a, ss: 0.1 * ctrl.reward_function(s, a, ss))
    V_lo = piql.generate_nu(weaker)
    exact = piql.generate_logz(ctrl, V_lo, beta=1.0)
    res = piql.mc_logz(ctrl, V_lo, beta=1.0, npaths=1 << 22, seed=1)
    err = np.array([res.logz(s) - exact(s) for s in exact.dict])
    assert np.abs(err).max() < 0.05 and abs(err.mean()) < 2e-3
    assert res.count_sa.sum() == 1 << 22 and res.transitions > 1 << 22


# ---- error behaviour -----------------------------------------------------------------------------------
def test_error_codes(piql, golden01):
    K = piql._capi
    eng = piql.Engine(0)
    with pytest.raises(piql.PiqlError) as e:
        eng.solve_nu()
    assert e.value.code == K.ERR_STATE
    w = golden01.walls.copy(order="F")
    w[0, 3] = 0
    with pytest.raises(piql.PiqlError) as e:
        eng.load_grid(w, golden01.goal, **golden01.params)
    assert e.value.code == K.ERR_ARG and "border" in str(e.value)
    with pytest.raises(piql.PiqlError) as e:
        eng.load_grid(golden01.walls, (0, 0), **golden01.params)
    assert e.value.code == K.ERR_ARG
    eng.load_grid(golden01.walls, golden01.goal, **golden01.params)
    with pytest.raises(piql.PiqlError) as e:
        eng.mc_rollout(10)
    assert e.value.code == K.ERR_STATE
    with pytest.raises(piql.PiqlError) as e:
        eng.solve_nu(max_sweeps=3)
    assert e.value.code == K.ERR_NOT_CONVERGED
    out, st = eng.solve_nu(max_sweeps=3, allow_unconverged=True)
    assert st.sweeps == 3 and not st.converged
    with pytest.raises(piql.PiqlError) as e:
        eng.mc_logz(np.zeros((golden01.H, golden01.W)), npaths=10, starts=[10 ** 6])
    assert e.value.code == K.ERR_ARG
    eng.close()


def test_config1_monte_carlo_protocol(engine):
    """golden r=0.00 gamma=0.99 grid, 2^24 paths from s0 = (2,17), seeds {0,1,2}, V = V_lo and V = nu,
    beta in {1, 0.18009...}: every estimate within 4 sigma of the exact log Z(s0)."""
    g = Golden("r0_00g0_99")
    load(engine, g)
    n = 1 << 24
    for V in (np.nan_to_num(g.V(7, 0)), np.nan_to_num(g.nu)):
        for beta in (1.0, float(g.betas[5])):
            exact, _ = engine.solve_logz(V, beta=beta)
            for seed in (0, 1, 2):
                sa, cnt, ls, st = engine.mc_logz(V, beta=beta, npaths=n, seed=seed, starts=[g.s0_idx])
                logw, _ = engine.mc_records(n)
                w = np.exp(logw.astype(np.float64) - logw.max())
                sigma = w.std() / np.sqrt(n) / w.mean()
                assert abs(ls[g.s0] - exact[g.s0]) < 4 * sigma + 2e-6, (beta, seed, ls[g.s0], exact[g.s0], sigma)
                assert int(cnt.sum()) == n


@pytest.mark.parametrize("name", ["r0_01g0_99", "r0_05g0_95"])
def test_config2_path_integral_vs_value_iteration_sweep(engine, name):
    """for every beta of the stored experiment and the first two iterates, the path-integral estimate of
    log Z(s0) against the value-iteration one (exact), and the resulting V_{k+1}(s0) against the reference's stored value."""
    g = Golden(name)
    load(engine, g)
    n = 1 << 22
    for b, beta in enumerate(float(x) for x in g.betas):
        for k in (0, 1):
            V = np.nan_to_num(g.V(b, k))
            exact, _ = engine.solve_logz(V, beta=beta)
            sa, cnt, ls, st = engine.mc_logz(V, beta=beta, npaths=n, seed=100 + 8 * k + b, starts=[g.s0_idx])
            logw, _ = engine.mc_records(n)
            w = np.exp(logw.astype(np.float64) - logw.max())
            sigma = w.std() / np.sqrt(n) / w.mean()
            assert abs(ls[g.s0] - exact[g.s0]) < 4 * sigma + 2e-6, (name, b, k, ls[g.s0], exact[g.s0], sigma)
            # V_{k+1}(s0) = V_k(s0) + log Z(s0) / beta   (value_iteration.jl:119-124), against data/*.jld2
            v_next = V[g.s0] + ls[g.s0] / beta
            assert abs(v_next - g.V(b, k + 1)[g.s0]) < (4 * sigma + 2e-6) / beta + 1e-8, (name, b, k)


def test_config3_long_horizon_rollouts(engine):
    """gamma = 0.9999.  Paths under the V_lo policy end at the goal long before the geometric horizon
    (1e4): lengths average ~630 steps with a tail of thousands, log-weights span tens of nats -- the reduction must stay in
    the log domain, and lanes are refilled as paths of very different length finish."""
    g = Golden("r0_00g0_9999")
    load(engine, g)
    V = np.nan_to_num(g.V(7, 0))
    n = 1 << 20
    exact, _ = engine.solve_logz(V, beta=1.0)
    sa, cnt, ls, st = engine.mc_logz(V, beta=1.0, npaths=n, seed=31, starts=[g.s0_idx])
    logw, meta = engine.mc_records(n)
    length = (meta >> 2).astype(np.int64)
    assert 560 < length.mean() < 700 and 1500 < length.max() < 20000 and length.min() >= 0
    assert int(length.sum()) == st.transitions
    assert np.isfinite(logw).all() and logw.max() <= 1e-3 and -200 < logw.min() < -20
    w = np.exp(logw.astype(np.float64) - logw.max())
    sigma = w.std() / np.sqrt(n) / w.mean()
    assert abs(ls[g.s0] - exact[g.s0]) < 4 * sigma + 2e-6, (ls[g.s0], exact[g.s0], sigma)
    # a direct float64 log-mean-exp of the records agrees with the kernel's (max, sum, count) reduction
    direct = logw.max() + np.log(np.mean(np.exp(logw.astype(np.float64) - logw.max())))
    assert abs(ls[g.s0] - direct) < 1e-6
