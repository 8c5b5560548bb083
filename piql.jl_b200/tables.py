"""Generic tabulation of a ControlProblem and the BinaryBandit problem (host side of the sparse-table ABI).

The reference's solvers see a problem only through closures (src/PIQL.jl:12-23); `tabulate_problem` turns any such
problem with finitely many reachable states into the flat arrays of piql_table_load, using nothing but those closures:
states are discovered from `initial_state()` the way the reference's `Growable` discovers them (by being evaluated at
successors, src/value_iteration.jl:12-18), and the transition law of (s, a) is read off `propagator_average` by
probing it with a recording function and with indicator functions.

    binary_bandit_problem, BanditState, update, reachable, state_iterator     src/Problems/BinaryBandit.jl:8-73
"""
from dataclasses import dataclass
import math

import numpy as np

from .gridworld import ControlProblem


@dataclass(frozen=True)
class ProblemTable:
    """CSR over (state, action) rows s*nactions + a; `states[i]` is the host object of state i."""
    states: list
    index: dict
    actions: list
    row_ptr: np.ndarray
    succ: np.ndarray
    prob: np.ndarray
    reward: np.ndarray
    log_prior: np.ndarray
    terminal: np.ndarray
    gamma: float

    def key(self):
        return (id(self.states), len(self.states), self.gamma)

    def values(self, f):
        """Dense vector of a callable value function over the tabulated states."""
        return np.array([float(f(s)) for s in self.states], dtype=np.float64)


def successors(ctrl, s, a):
    """[(s', p)] of (s, a) through propagator_average alone: one recording probe, then one indicator probe per s'."""
    seen = []

    def rec(s1):
        if s1 not in seen:
            seen.append(s1)
        return 0.0
    ctrl.propagator_average(s, a, rec)
    out = []
    for s1 in seen:
        p = float(ctrl.propagator_average(s, a, lambda x, s1=s1: 1.0 if x == s1 else 0.0))
        if p > 0.0:
            out.append((s1, p))
    return out


def tabulate_problem(ctrl, roots=None, max_states=2_000_000):
    """Breadth-first closure of the reachable set from `roots` (default: ctrl.initial_state()), then the CSR arrays."""
    actions = list(ctrl.action_space)
    roots = [ctrl.initial_state()] if roots is None else list(roots)
    states, index = [], {}
    for r in roots:
        if r not in index:
            index[r] = len(states)
            states.append(r)
    rows = []  # per state: list over actions of [(succ index, p, reward)]
    i = 0
    while i < len(states):
        s = states[i]
        term = bool(ctrl.terminal_condition(s))
        per_a = []
        for a in actions:
            ent = []
            if not term:
                for s1, p in successors(ctrl, s, a):
                    if s1 not in index:
                        if len(states) >= max_states:
                            raise MemoryError("more than %d reachable states" % max_states)
                        index[s1] = len(states)
                        states.append(s1)
                    ent.append((index[s1], p, float(ctrl.reward_function(s, a, s1))))
            per_a.append(ent)
        rows.append((term, per_a))
        i += 1
    na = len(actions)
    ptr = np.zeros(len(states) * na + 1, dtype=np.int64)
    succ, prob, rew = [], [], []
    lprior = np.zeros(len(states) * na, dtype=np.float64)
    terminal = np.zeros(len(states), dtype=np.uint8)
    for si, (term, per_a) in enumerate(rows):
        terminal[si] = term
        for ai, ent in enumerate(per_a):
            for (j, p, r) in ent:
                succ.append(j); prob.append(p); rew.append(r)
            ptr[si * na + ai + 1] = len(succ)
            pr = float(ctrl.action_prior(states[si], actions[ai]))
            lprior[si * na + ai] = math.log(pr) if pr > 0 else -math.inf
    return ProblemTable(states, index, actions, ptr, np.asarray(succ, dtype=np.int32), np.asarray(prob, dtype=np.float64),
                        np.asarray(rew, dtype=np.float64), lprior, terminal, float(ctrl.gamma))


# ---- BinaryBandit (src/Problems/BinaryBandit.jl) -----------------------------------------------------------------
@dataclass(frozen=True)
class BanditState:
    """belief: ((heads, tails), ...) per coin; rounds left   BinaryBandit.jl:8-11"""
    belief: tuple
    rounds: int


def update(bs, a, o):
    """Observe outcome o of coin a (0-based)   BinaryBandit.jl:18-23"""
    b = list(bs.belief)
    h, t = b[a]
    b[a] = (h + (1 if o else 0), t + (0 if o else 1))
    return BanditState(tuple(b), bs.rounds - 1)


def reachable(bs):
    """BinaryBandit.jl:25-27"""
    return [update(bs, a, o) for a in range(len(bs.belief)) for o in (True, False)]


def binary_bandit_problem(rounds=10, ncoins=2, prior=(1, 1), payoff=1.0, gamma=1.0, γ=None, rng=None):
    """binary_bandit_problem(; rounds, ncoins, prior, payoff, γ)   BinaryBandit.jl:30-58.

    logpropagatorexp is the log-mean-exp the recursions need, log sum_o p(o) exp f(update(s,a,o)); the reference's line
    (:44-47) multiplies instead of adding logs and is not used by anything the reference tests."""
    gamma = gamma if γ is None else γ
    rng = np.random.default_rng() if rng is None else rng

    def p_heads(s, a):
        h, t = s.belief[a]
        return h / (h + t)

    def propagator(s, a):
        return update(s, a, bool(rng.random() < p_heads(s, a)))

    def propagator_average(s, a, f):
        ph = p_heads(s, a)
        return f(update(s, a, False)) * (1 - ph) + f(update(s, a, True)) * ph

    def logpropagatorexp(s, a, f):
        ph = p_heads(s, a)
        x = [math.log(1 - ph) + f(update(s, a, False)) if ph < 1 else -math.inf,
             math.log(ph) + f(update(s, a, True)) if ph > 0 else -math.inf]
        m = max(x)
        return m if math.isinf(m) else m + math.log(sum(math.exp(v - m) for v in x))

    start = BanditState(tuple(tuple(prior) for _ in range(ncoins)), rounds)
    return ControlProblem(
        action_space=range(ncoins),
        action_prior=lambda s, a: 1.0 / ncoins,
        propagator=propagator,
        reward_function=lambda s0, a, s1: (s1.belief[a][0] - s0.belief[a][0]) * payoff,
        propagator_average=propagator_average,
        logpropagatorexp=logpropagatorexp,
        terminal_condition=lambda s: s.rounds < 1,
        initial_state=lambda: start,
        gamma=float(gamma),
    )


def bandit_state_iterator(rounds=10, ncoins=2, prior=(1, 1)):
    """All states reachable within `rounds` observations, deepest first (what BinaryBandit.jl:61-73 means to return; the
    reference discards its `vcat` and so returns the initial state only)."""
    cur = [BanditState(tuple(tuple(prior) for _ in range(ncoins)), rounds)]
    out, seen = list(cur), set(cur)
    for _ in range(rounds):
        nxt = []
        for s in cur:
            for r in reachable(s):
                if r not in seen:
                    seen.add(r)
                    nxt.append(r)
        out.extend(nxt)
        cur = nxt
    return list(reversed(out))
