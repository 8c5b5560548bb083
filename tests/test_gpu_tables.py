"""The sparse-table path on the GPU (csrc/table.cu, piql_table_*): any ControlProblem the host can tabulate.  Checked
against the oracle's table restatement on the BinaryBandit and on a tabulated gridworld, against the grid kernels on the same
gridworld (two independent CUDA paths for one problem), and through the reference-style interface.  -m gpu."""
import numpy as np
import pytest

import oracle
from conftest import Golden

pytestmark = pytest.mark.gpu


def test_bandit_solves_match_the_table_oracle(engine, piql):
    bb = piql.binary_bandit_problem(rounds=8, ncoins=3, payoff=0.6, gamma=0.97)
    tab = piql.tabulate_problem(bb)
    engine.load_table(tab)
    nu, st = engine.table_solve("nu")
    ref, sw, _ = oracle.table_solve(tab, "nu")
    assert st.converged and st.kernel == 4 and np.abs(nu - ref).max() < 1e-12
    vlo, _ = engine.table_solve("nu", reward_scale=0.3)
    assert np.abs(vlo - oracle.table_solve(tab, "nu", reward_scale=0.3)[0]).max() < 1e-12
    for beta in (1.0, 0.2):
        v1, logz, st1 = engine.table_update_v(vlo, beta=beta)
        lref, _, _ = oracle.table_solve(tab, "logz", V=vlo, beta=beta)
        assert st1.converged and np.abs(logz - lref).max() < 1e-10 and np.abs(v1 - (vlo + lref / beta)).max() < 1e-9
    gap, st2 = engine.table_solve("gap", V=vlo, nu=nu)
    gref, _, _ = oracle.table_solve(tab, "gap", V=vlo, nu=ref)
    assert st2.converged and np.abs(gap - gref).max() < 1e-10
    # known answers of the algebra: log Z = 0 and gap = 0 at V = nu
    lz0, _ = engine.table_solve("logz", V=nu, beta=0.5)
    g0, _ = engine.table_solve("gap", V=nu, nu=nu)
    assert np.abs(lz0).max() < 1e-10 and np.abs(g0).max() < 1e-10


def test_finite_horizon_bandit_at_gamma_one(engine, piql):
    bb = piql.binary_bandit_problem(rounds=10, ncoins=2)      # the reference's defaults: gamma = 1
    tab = piql.tabulate_problem(bb)
    engine.load_table(tab)
    nu, st = engine.table_solve("nu", check_every=4)
    ref, _, _ = oracle.table_solve(tab, "nu")
    assert st.converged and st.sweeps <= 16 and np.abs(nu - ref).max() < 1e-12
    v1, logz, _ = engine.table_update_v(0.5 * nu, beta=1.0, check_every=4)
    lref, _, _ = oracle.table_solve(tab, "logz", V=0.5 * nu, beta=1.0)
    assert np.abs(logz - lref).max() < 1e-10


def test_table_path_equals_grid_path_on_the_reference_gridworld(engine, piql):
    g = Golden("r0_01g0_99")
    gw = piql.Gridworld(g.walls != 0, g.goal)
    ctrl = piql.modify(piql.gridworld_ctrl(gw, **{k: g.params[k] for k in ("randomness", "step_cost", "reward", "gamma")}), grid=None)
    tab = piql.tabulate_problem(ctrl, roots=[g.goal] + piql.state_iterator(gw))
    engine.load_table(tab)
    engine.load_grid(g.walls, g.goal, **g.params)
    pick = lambda A: np.array([A[s] for s in tab.states])
    nu_t, _ = engine.table_solve("nu")
    nu_g, _ = engine.solve_nu()
    assert np.abs(nu_t - pick(nu_g)).max() < 1e-11 and np.abs(nu_t - pick(g.nu)).max() < 1e-9
    V = np.nan_to_num(g.V(3, 1))
    beta = float(g.betas[3])
    v_t, _, _ = engine.table_update_v(pick(V), beta=beta)
    v_g, _ = engine.update_v(V, beta=beta)
    assert np.abs(v_t - pick(v_g)).max() < 1e-9 / beta and np.abs(v_t - pick(g.V(3, 2))).max() < 1e-9 + 5e-12 / beta
    gp_t, _ = engine.table_solve("gap", V=pick(V), nu=pick(np.nan_to_num(g.nu)))
    assert np.abs(gp_t - pick(g.G(3, 1))).max() < 2e-9


def test_bandit_through_the_reference_interface(piql):
    bb = piql.binary_bandit_problem(rounds=6, ncoins=2, gamma=0.95)
    nu = piql.generate_nu(bb)
    weaker = piql.modify(bb, reward_function=lambda s, a, ss: 0.5 * bb.reward_function(s, a, ss))
    v_lo = piql.generate_nu(weaker)
    s0 = bb.initial_state()
    assert nu(s0) > v_lo(s0) > 0
    v1 = piql.logz_updateV(bb, v_lo, beta=1.0)
    v0 = piql.logz_updateV(bb, v_lo, beta=1e-3)
    # the bounds of test/runtests.jl:33-36, on the bandit
    assert all(v1(s) >= nu(s) - 1e-2 for s in nu.dict) and all(v0(s) <= nu(s) + 1e-2 for s in nu.dict)
    gap = piql.generate_optimality_gap(bb, v_lo, nu=nu)
    assert gap(s0) > 0 and abs(piql.generate_optimality_gap(bb, nu, nu=nu)(s0)) < 1e-10
    lz = piql.generate_logz(bb, nu, beta=0.7)
    assert max(abs(v) for v in lz.dict.values()) < 1e-10


def test_table_rollouts_bit_exact_and_estimate_hits_exact_logz(engine, piql):
    """Monte-Carlo path sampling on tabulated problems: per-path records equal the oracle's bit for bit (same float tables,
    same Philox stream), every path lands in one bucket, and the estimate sits within its error of the exact log Z."""
    for ctrl, roots in ((piql.binary_bandit_problem(rounds=7, ncoins=2, payoff=0.8, gamma=0.96), None),
                        (None, "grid")):
        if ctrl is None:
            g = Golden("r0_01g0_99")
            gw = piql.Gridworld(g.walls != 0, g.goal)
            ctrl = piql.modify(piql.gridworld_ctrl(gw, **{k: g.params[k] for k in ("randomness", "step_cost", "reward", "gamma")}), grid=None)
            tab = piql.tabulate_problem(ctrl, roots=[g.goal] + piql.state_iterator(gw))
        else:
            tab = piql.tabulate_problem(ctrl)
        engine.load_table(tab)
        nu, _ = engine.table_solve("nu")
        V = 0.4 * nu
        exact, _ = engine.table_solve("logz", V=V, beta=1.0)
        s0 = 1 if roots == "grid" else 0
        n = 1 << 20
        sa, cnt, ls, st, (cdf, bf, cum) = engine.table_mc_logz(V, beta=1.0, npaths=n, seed=9, starts=[s0], nnz=len(tab.succ),
                                                               want_tables=True)
        logw, meta = engine.mc_records(n)
        lw_ref, meta_ref = oracle.table_rollout(tab, cdf, bf, cum, V.astype(np.float32), n, starts=[s0], seed=9)
        assert np.array_equal(meta, meta_ref) and np.array_equal(logw.view(np.uint32), lw_ref.view(np.uint32))
        assert int(cnt.sum()) == n and int(cnt[s0].sum()) == n and st.transitions == int((meta >> 4).sum())
        w = np.exp(logw.astype(np.float64) - logw.max())
        sigma = w.std() / np.sqrt(n) / w.mean()
        assert abs(ls[s0] - exact[s0]) < 4 * sigma + 2e-6, (ls[s0], exact[s0], sigma)
        ref_lse = np.log(np.mean(np.exp(lw_ref.astype(np.float64))))
        assert abs(ls[s0] - ref_lse) < 1e-6
        # paths from every state: counts tile the states; first-action buckets hold the right paths
        sa2, cnt2, ls2, _ = engine.table_mc_logz(V, beta=1.0, npaths=len(tab.states) * 64, seed=3)
        assert np.all(cnt2.sum(axis=1) == 64)
        term = tab.terminal != 0
        assert np.allclose(ls2[term], -V[term], atol=1e-6)          # Z(terminal) = exp(-V)


def test_bandit_mc_through_the_reference_interface(piql):
    bb = piql.binary_bandit_problem(rounds=6, ncoins=2, gamma=0.95)
    nu = piql.generate_nu(bb)
    v = piql.Growable.from_table(piql.solvers.table_for(bb)[1], 0.5 * np.array(list(nu.dict.values())))
    exact = piql.generate_logz(bb, v, beta=1.0)
    s0 = bb.initial_state()
    res = piql.mc_logz(bb, v, beta=1.0, npaths=1 << 20, seed=4, starts=[s0])
    assert abs(res.logz(s0) - exact(s0)) < 0.01 and res.transitions > 1 << 20
    assert res.count_sa.sum() == 1 << 20


def test_table_solves_and_sampler_with_a_q_table(engine, piql):
    """piql_table_set_q: a tabulated Q enters Qnormalization / controlQ / action_probability / fitness of the table path --
    log Z and gap solves equal the table oracle (1e-12) and the GPU grid path with the same Q (1e-10); the sampler's records
    stay bit-exact against the oracle on the tables built from that Q; clearing Q restores the V-determined solve."""
    g = Golden("r0_01g0_99")
    gw = piql.Gridworld(g.walls != 0, g.goal)
    ctrl = piql.modify(piql.gridworld_ctrl(gw, **g.params), grid=None)
    tab = piql.tabulate_problem(ctrl, roots=[g.goal] + piql.state_iterator(gw))
    rng = np.random.default_rng(3)
    Qd = rng.normal(size=(g.H * g.W, 4)) * 0.2
    idx = np.array([s[0] + g.H * s[1] for s in tab.states])
    Qt, V = Qd[idx], np.nan_to_num(g.V(5, 2))
    Vt = V.ravel(order="F")[idx]
    engine.load_table(tab)
    base, _ = engine.table_solve("logz", V=Vt, beta=0.6)
    engine.table_set_q(Qt)
    out, _ = engine.table_solve("logz", V=Vt, beta=0.6)
    ref, _, _ = oracle.table_solve(tab, "logz", V=Vt, beta=0.6, Q=Qt)
    assert np.abs(out - ref).max() < 1e-12 and np.abs(out - base).max() > 1e-3
    nut = np.nan_to_num(g.nu).ravel(order="F")[idx]
    gap, _ = engine.table_solve("gap", V=Vt, nu=nut)
    refg, _, _ = oracle.table_solve(tab, "gap", V=Vt, nu=nut, Q=Qt)
    assert np.abs(gap - refg).max() < 1e-11
    n = 1 << 18
    sa, cnt, ls, st, (cdf, bf, cum) = engine.table_mc_logz(Vt, beta=0.6, npaths=n, seed=5, starts=[3], nnz=len(tab.succ), want_tables=True)
    logw, meta = engine.mc_records(n)
    lw_ref, meta_ref = oracle.table_rollout(tab, cdf, bf, cum, Vt.astype(np.float32), n, starts=[3], seed=5)
    assert np.array_equal(meta, meta_ref) and np.array_equal(logw.view(np.uint32), lw_ref.view(np.uint32))
    w = np.exp(logw.astype(np.float64) - logw.max())
    assert abs(ls[3] - out[3]) < 5 * w.std() / np.sqrt(n) / w.mean() + 1e-5       # the estimate targets the Q solve's fixed point
    engine.table_set_q(None)
    again, _ = engine.table_solve("logz", V=Vt, beta=0.6)
    assert np.array_equal(again, base)
    # the grid path with the same Q
    engine.load_grid(g.walls, g.goal, **g.params)
    grid, _ = engine.solve_logz(V, beta=0.6, Q=Qd)
    assert np.abs(grid.ravel(order="F")[idx] - out).max() < 1e-10
    # through the reference interface: Q as a callable (s, a) -> value
    lz = piql.generate_logz(ctrl, lambda s: float(V[s]), beta=0.6, Q=lambda s, a: float(Qd[s[0] + g.H * s[1], a]))
    assert max(abs(lz(s) - o) for s, o in zip(tab.states, out)) < 1e-12
