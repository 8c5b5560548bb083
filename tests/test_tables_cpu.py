"""Host side of the sparse-table ABI on CPU: the generic tabulator (a ControlProblem seen only through its closures), the
BinaryBandit mirror (src/Problems/BinaryBandit.jl) with the reference's own assertions (test/runtests.jl:44-48), and the
oracle's table restatement pinned through a tabulated gridworld against the grid restatement (itself pinned to data/*.jld2)."""
import math

import numpy as np

import oracle
from conftest import Golden


def test_reference_bandit_assertions(piql):
    bb = piql.binary_bandit_problem(rng=np.random.default_rng(0))
    i0 = bb.initial_state()
    i1 = bb.propagator(i0, 0)
    assert i0.rounds - i1.rounds == 1
    assert piql.average_reward(bb, i0, 0) == 1 / 2                       # average reward is half payoff
    assert piql.average_reward(bb, i1, 0) == i1.belief[0][0] / sum(i1.belief[0])
    assert len(piql.reachable(i0)) == 4
    its = piql.bandit_state_iterator(rounds=3, ncoins=2)
    assert its[-1] == i0.__class__(i0.belief, 3) and all(s.rounds >= 0 for s in its)


def test_tabulator_recovers_the_bandit_tree(piql):
    bb = piql.binary_bandit_problem(rounds=5, ncoins=2)
    tab = piql.tabulate_problem(bb)
    assert len(tab.states) == len(piql.bandit_state_iterator(rounds=5, ncoins=2))
    na = len(tab.actions)
    for si, s in enumerate(tab.states):
        for a in range(na):
            lo, hi = tab.row_ptr[si * na + a], tab.row_ptr[si * na + a + 1]
            if tab.terminal[si]:
                assert lo == hi
                continue
            assert abs(tab.prob[lo:hi].sum() - 1.0) < 1e-12 and hi - lo == 2
            h, t = s.belief[a]
            got = {tab.states[j]: (p, r) for j, p, r in zip(tab.succ[lo:hi], tab.prob[lo:hi], tab.reward[lo:hi])}
            (ph, rh), (pt, rt) = got[piql.update(s, a, True)], got[piql.update(s, a, False)]
            assert abs(ph - h / (h + t)) < 1e-15 and abs(pt - t / (h + t)) < 1e-15 and (rh, rt) == (1.0, 0.0)
    assert np.allclose(tab.log_prior, math.log(0.5))


def test_table_oracle_solves_the_bandit_by_backward_induction(piql):
    """gamma = 1, finite horizon: nu is the soft Bellman value computed by plain recursion over the belief tree."""
    bb = piql.binary_bandit_problem(rounds=6, ncoins=2, payoff=0.7)
    tab = piql.tabulate_problem(bb)
    nu, sweeps, res = oracle.table_solve(tab, "nu")
    assert sweeps <= 8 and res == 0.0        # a DAG of depth 6 closes in 6 sweeps (+ the one that sees no change)
    memo = {}

    def soft(s):
        if s.rounds < 1:
            return 0.0
        if s not in memo:
            x = []
            for a in range(2):
                h, t = s.belief[a]
                ph = h / (h + t)
                x.append(math.log(0.5) + ph * (0.7 + soft(piql.update(s, a, True))) + (1 - ph) * soft(piql.update(s, a, False)))
            m = max(x)
            memo[s] = m + math.log(sum(math.exp(v - m) for v in x))
        return memo[s]
    assert max(abs(nu[i] - soft(s)) for i, s in enumerate(tab.states)) < 1e-12


def test_table_oracle_equals_grid_oracle_on_a_tabulated_gridworld(piql):
    g = Golden("r0_01g0_99")
    gw = piql.Gridworld(g.walls != 0, g.goal)
    ctrl = piql.modify(piql.gridworld_ctrl(gw, **{k: g.params[k] for k in ("randomness", "step_cost", "reward", "gamma")}), grid=None)
    tab = piql.tabulate_problem(ctrl, roots=[g.goal] + piql.state_iterator(gw))
    assert len(tab.states) == 251
    org = g.oracle()
    pick = lambda A: np.array([A[s] for s in tab.states])
    nu, _, _ = oracle.table_solve(tab, "nu")
    assert np.abs(nu - pick(g.nu)).max() < 1e-9                       # the reference's stored nu
    V = pick(np.nan_to_num(g.V(7, 0)))
    lz, _, _ = oracle.table_solve(tab, "logz", V=V, beta=1.0)
    assert np.abs(V + lz - pick(g.V(7, 1))).max() < 1e-9              # the reference's stored first beta = 1 iterate
    ref, _ = org.solve("logz", V=np.nan_to_num(g.V(7, 0)), beta=1.0)
    assert np.abs(lz - pick(ref)).max() < 1e-11
    gp, _, _ = oracle.table_solve(tab, "gap", V=pick(np.nan_to_num(g.V(7, 2))), nu=pick(np.nan_to_num(g.nu)))
    assert np.abs(gp - pick(g.G(7, 2))).max() < 1e-9                  # the reference's stored gap


def test_table_oracle_with_a_q_table_equals_the_grid_oracle(golden01):
    """A tabulated Q (any Q callable of PIQL.jl:72-89) through the table restatement against the grid restatement, which
    is pinned to data/*.jld2: log Z and gap solves with a random Q agree on the tabulated reference gridworld."""
    import piql_jl_b200 as piql
    g = golden01
    gw = piql.Gridworld(g.walls != 0, g.goal)
    ctrl = piql.modify(piql.gridworld_ctrl(gw, **g.params), grid=None)
    tab = piql.tabulate_problem(ctrl, roots=[g.goal] + piql.state_iterator(gw))
    org = g.oracle()
    rng = np.random.default_rng(3)
    Qd = rng.normal(size=(g.H * g.W, 4)) * 0.2
    idx = np.array([s[0] + g.H * s[1] for s in tab.states])
    Qt = Qd[idx]
    V = np.nan_to_num(g.V(5, 2))
    Vt = V.ravel(order="F")[idx]
    ref, _ = org.solve("logz", V=V, Q=Qd, beta=0.6)
    out, sweeps, res = oracle.table_solve(tab, "logz", V=Vt, beta=0.6, Q=Qt)
    assert np.abs(out - ref.ravel(order="F")[idx]).max() < 1e-10
    nu = np.nan_to_num(g.nu)
    refg, _ = org.solve("gap", V=V, Q=Qd, nu=nu)
    outg, _, _ = oracle.table_solve(tab, "gap", V=Vt, nu=nu.ravel(order="F")[idx], Q=Qt)
    assert np.abs(outg - refg.ravel(order="F")[idx]).max() < 1e-9
