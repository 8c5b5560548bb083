"""Host-side mirror of PIQL's solver entry points; the work happens in the CUDA library.

Same names, arguments and return types as the reference (γ/β/ν spelled out in ASCII):
    Growable                              src/value_iteration.jl:3-18
    generate_nu (generate_ν)              src/value_iteration.jl:77-87
    generate_logz                         src/value_iteration.jl:107-117
    logz_updateV                          src/value_iteration.jl:119-124
    generate_optimality_gap               src/value_iteration.jl:144-154
    kl_cost, choose_action, path_transitions, fitness_reward_estimate
                                          src/path_functions.jl:12-41 (repaired: the reference
                                          file does not load as written)
    average_reward, controlV, Qnormalization, controlQ, action_probability, fitness
                                          src/PIQL.jl:55-89 (host algebra on closures, for
                                          callers and tests; the kernels restate it per cell)
    mc_logz                               new: Monte-Carlo estimate of generate_logz's fixed point
    learning_traj, run_experiment         scripts/experiments.jl:10-39 (+ batched forms: all beta chains / all gap
                                          solves in one library call, scripts/distributed_experiments.jl:26,28)
Every solver returns a Growable, as in the reference; `.array` is the dense (H, W) view
(NaN at walls) the library produced.
"""
import math

import numpy as np

from . import gridworld as gwmod
from .engine import Engine

_engines = {}


def engine_for(ctrl, device=0):
    """The per-device Engine with ctrl's tabulated problem loaded (re-used while it is unchanged)."""
    tab = gwmod.tabulate(ctrl)
    eng = _engines.get(device)
    if eng is None:
        eng = _engines[device] = Engine(device)
    key = tab.key()
    if eng.grid_key != key:
        eng.load_grid(tab.walls, tab.goal, tab.randomness, tab.step_cost, tab.reward, tab.gamma)
        eng.grid_key = key
    return eng


class Growable:
    """Callable value function backed by a dictionary that grows on demand (value_iteration.jl:3-18).

    Results coming back from the GPU carry the dense array as well, so V(s) is an array lookup
    and `.dict` is only materialised when somebody asks for it.
    """

    def __init__(self, s=None, default=lambda x: 0.0, array=None, walls=None, key=tuple):
        self.default = default
        self.array = array
        self._walls = walls
        self._key = key   # grid states are index pairs (any sequence); other problems use their state objects as they are
        self._dict = None if array is not None else ({} if s is None else {key(s): 0.0})

    @classmethod
    def from_dense(cls, array, walls, default=lambda x: 0.0):
        return cls(default=default, array=np.asfortranarray(array, dtype=np.float64), walls=np.asarray(walls) != 0)

    @classmethod
    def from_table(cls, tab, values, default=lambda x: 0.0):
        """Value function over the states of a tabulated (non-grid) problem."""
        g = cls(default=default, key=lambda s: s)
        g._dict = {st: float(v) for st, v in zip(tab.states, values)}
        return g

    @property
    def dict(self):
        if self._dict is None:
            idx = np.argwhere(~self._walls)
            self._dict = {(int(i), int(j)): float(self.array[i, j]) for i, j in idx}
        return self._dict

    def __call__(self, s):
        s = self._key(s)
        if self._dict is None and self.array is not None:
            if not self._walls[s]:
                return float(self.array[s])
            self.dict  # a wall cell was asked for: fall through to the growing dictionary
        if s not in self._dict:
            self._dict[s] = self.default(s)
        return self._dict[s]

    def dense(self, walls):
        """(H, W) array with NaN at walls, for handing the function back to the library."""
        if self.array is not None:
            return self.array
        out = np.full(walls.shape, np.nan, order="F")
        for (i, j), v in self._dict.items():
            out[i, j] = v
        return out


def _dense(V, eng):
    if isinstance(V, Growable):
        return np.nan_to_num(V.dense(eng.walls))
    return np.nan_to_num(np.asarray(V, dtype=np.float64))


# ---- KL-control algebra on closures (src/PIQL.jl:55-89) ---------------------------------------------
def average_reward(ctrl, s, a):
    return ctrl.propagator_average(s, a, lambda s1: ctrl.reward_function(s, a, s1))


def controlV(ctrl, V, s, a):
    return ctrl.gamma * ctrl.propagator_average(s, a, V) - V(s) + average_reward(ctrl, s, a)


def Qnormalization(ctrl, Q, s):
    x = [math.log(ctrl.action_prior(s, a)) + Q(s, a) for a in ctrl.action_space]
    m = max(x)
    return m + math.log(sum(math.exp(v - m) for v in x))


def controlQ(ctrl, Q, s, a):
    return Q(s, a) - Qnormalization(ctrl, Q, s)


def action_probability(ctrl, Q, s, a):
    return ctrl.action_prior(s, a) * math.exp(controlQ(ctrl, Q, s, a))


def fitness(ctrl, V, Q, s, a):
    return controlV(ctrl, V, s, a) - controlQ(ctrl, Q, s, a)


def kl_cost(ctrl, s, Q):
    """sum_a klterm(pi_Q(a|s), pi_0(a|s)), klterm(x, y) = xlogx(x) - xlogy(x, y)   path_functions.jl:17-20"""
    out = 0.0
    for a in ctrl.action_space:
        x, y = action_probability(ctrl, Q, s, a), ctrl.action_prior(s, a)
        out += (x * math.log(x) if x > 0 else 0.0) - (x * math.log(y) if x > 0 else 0.0)
    return out


def choose_action(ctrl, s, Q, rng):
    """sample(action_space, Weights(w)): t = rand()*sum(w), walk while cw < t   path_functions.jl:12-15"""
    w = [action_probability(ctrl, Q, s, a) for a in ctrl.action_space]
    t = rng.random() * sum(w)
    i, cw = 0, w[0]
    while cw < t and i < len(w) - 1:
        i += 1
        cw += w[i]
    return list(ctrl.action_space)[i]


def path_transitions(ctrl, s0, Q, rng, nsteps=20):
    """Generator of (s, a, s') along one path under pi_Q   path_functions.jl:23-32"""
    s = s0
    for _ in range(nsteps):
        a = choose_action(ctrl, s, Q, rng)
        ss = ctrl.propagator(s, a)
        yield (s, a, ss)
        s = ss


def fitness_reward_estimate(ctrl, s0, V, Q, rng, nsteps=20):
    """Running V(s0) + sum_t fitness(s_t, a_t) along one path   path_functions.jl:34-41"""
    total, out = V(s0), []
    for (s, a, _ss) in path_transitions(ctrl, s0, Q, rng, nsteps=nsteps):
        total += fitness(ctrl, V, Q, s, a)
        out.append(total)
    return out


def direct_reward_estimate(ctrl, s0, V, Q, rng, nsteps=20):
    """Running V(s0) + sum_t [reward_function(s_t, a_t, s'_t) - kl_cost(s_t)] along one path   path_functions.jl:43-50
    (the reference writes ctrl.reward, a field that does not exist; reward_function is meant)"""
    total, out = V(s0), []
    for (s, a, ss) in path_transitions(ctrl, s0, Q, rng, nsteps=nsteps):
        total += ctrl.reward_function(s, a, ss) - kl_cost(ctrl, s, Q)
        out.append(total)
    return out


def path_estimates(ctrl, V, s0=None, npaths=1 << 16, nsteps=20, seed=0, device=0):
    """Both path estimates for npaths paths at once on the GPU (piql_path_estimates), Q = the V-determined Q-function
    (s, a) -> controlV(ctrl, V, s, a) as everywhere in the reference's scripts.  s0: one start state, a list of them
    (path p starts at s0[p % len]) or None for every free cell.  Returns (fitness[npaths, nsteps], direct[npaths, nsteps])."""
    eng = engine_for(ctrl, device)
    starts = None
    if s0 is not None:
        lst = [s0] if isinstance(s0[0], (int, np.integer)) else list(s0)
        starts = [int(i) + eng.H * int(j) for (i, j) in lst]
    fit, dr = eng.path_estimates(_dense(V, eng), npaths, nsteps=nsteps, seed=seed, starts=starts)
    return fit, dr


# ---- problems that are not the implicit gridworld: tabulate through the closures, solve on the table ABI ---------
_tables = {}


def table_for(ctrl, device=0):
    """(engine, ProblemTable) with ctrl's tabulated problem loaded; the table is built once per ControlProblem object."""
    from . import tables
    ent = _tables.get(id(ctrl))
    if ent is None or ent[0] is not ctrl:
        ent = _tables[id(ctrl)] = (ctrl, tables.tabulate_problem(ctrl))
    tab = ent[1]
    eng = _engines.get(device)
    if eng is None:
        eng = _engines[device] = Engine(device)
    if getattr(eng, "table_key", None) != tab.key():
        eng.load_table(tab)
    return eng, tab


def _table_values(tab, V):
    return tab.values(V) if callable(V) else np.asarray(V, dtype=np.float64)


def _table_q(tab, Q):
    """Q over the tabulated (state, action) pairs: a callable Q(s, a) as in the reference (PIQL.jl:72-89), or an array."""
    if Q is None:
        return None
    if callable(Q):
        return np.array([[float(Q(s, a)) for a in tab.actions] for s in tab.states], dtype=np.float64)
    return np.asarray(Q, dtype=np.float64).reshape(len(tab.states), len(tab.actions))


# ---- solvers on the GPU -----------------------------------------------------------------------------
def generate_nu(ctrl, tol=1e-12, max_sweeps=10 ** 7, precision="f64", device=0, return_stats=False):
    if ctrl.grid is None:
        eng, tab = table_for(ctrl, device)
        out, st = eng.table_solve("nu", tol=tol, max_sweeps=max_sweeps)
        g = Growable.from_table(tab, out)
        return (g, st) if return_stats else g
    eng = engine_for(ctrl, device)
    out, st = eng.solve_nu(1.0, tol=tol, max_sweeps=max_sweeps, precision=precision)
    g = Growable.from_dense(out, eng.walls)
    return (g, st) if return_stats else g


generate_ν = generate_nu


def generate_logz(ctrl, V, beta=1.0, β=None, Q=None, tol=1e-12, max_sweeps=10 ** 7, precision="f64", device=0,
                  return_stats=False):
    beta = beta if β is None else β
    if ctrl.grid is None:
        eng, tab = table_for(ctrl, device)
        eng.table_set_q(_table_q(tab, Q))
        try:
            out, st = eng.table_solve("logz", V=_table_values(tab, V), beta=beta, tol=tol, max_sweeps=max_sweeps)
        finally:
            eng.table_set_q(None)
        g = Growable.from_table(tab, out)
        return (g, st) if return_stats else g
    eng = engine_for(ctrl, device)
    out, st = eng.solve_logz(_dense(V, eng), beta=beta, Q=Q, tol=tol, max_sweeps=max_sweeps, precision=precision)
    g = Growable.from_dense(out, eng.walls)
    return (g, st) if return_stats else g


def logz_updateV(ctrl, V, beta=1.0, β=None, tol=1e-12, max_sweeps=10 ** 7, precision="f64", device=0,
                 return_stats=False):
    beta = beta if β is None else β
    if ctrl.grid is None:
        eng, tab = table_for(ctrl, device)
        out, _logz, st = eng.table_update_v(_table_values(tab, V), beta=beta, tol=tol, max_sweeps=max_sweeps)
        g = Growable.from_table(tab, out, default=getattr(V, "default", lambda x: 0.0))
        return (g, st) if return_stats else g
    eng = engine_for(ctrl, device)
    out, st = eng.update_v(_dense(V, eng), beta=beta, tol=tol, max_sweeps=max_sweeps, precision=precision)
    g = Growable.from_dense(out, eng.walls, default=getattr(V, "default", lambda x: 0.0))
    return (g, st) if return_stats else g


def generate_z(ctrl, V, beta=1.0, β=None, Q=None, tol=1e-9, max_sweeps=10 ** 7, precision="f64", device=0,
               return_stats=False):
    """generate_z (src/value_iteration.jl:42-51): the linear-space twin of generate_logz, default x -> 1.0, stop rule
    sum |log new - log old| < 1e-9.  Gridworld problems only (the reference never runs it on anything else)."""
    beta = beta if β is None else β
    if ctrl.grid is None:
        raise NotImplementedError("generate_z is provided for gridworld problems; use generate_logz for tabulated ones")
    eng = engine_for(ctrl, device)
    out, st = eng.solve_z(_dense(V, eng), beta=beta, Q=Q, tol=tol, max_sweeps=max_sweeps, precision=precision)
    g = Growable.from_dense(out, eng.walls, default=lambda x: 1.0)
    return (g, st) if return_stats else g


def z_updateV(ctrl, V, beta=1.0, β=None, tol=1e-9, max_sweeps=10 ** 7, precision="f64", device=0, return_stats=False):
    """z_updateV (src/value_iteration.jl:53-59): V + log(Z)/beta."""
    beta = beta if β is None else β
    if ctrl.grid is None:
        raise NotImplementedError("z_updateV is provided for gridworld problems; use logz_updateV for tabulated ones")
    eng = engine_for(ctrl, device)
    out, st = eng.z_update_v(_dense(V, eng), beta=beta, tol=tol, max_sweeps=max_sweeps, precision=precision)
    g = Growable.from_dense(out, eng.walls, default=getattr(V, "default", lambda x: 0.0))
    return (g, st) if return_stats else g


def generate_optimality_gap(ctrl, V, nu=None, ν=None, Q=None, tol=1e-12, max_sweeps=10 ** 7, precision="f64", device=0,
                            return_stats=False):
    nu = nu if ν is None else ν
    if nu is None:
        nu = generate_nu(ctrl, device=device)
    if ctrl.grid is None:
        eng, tab = table_for(ctrl, device)
        eng.table_set_q(_table_q(tab, Q))
        try:
            out, st = eng.table_solve("gap", V=_table_values(tab, V), nu=_table_values(tab, nu), tol=tol, max_sweeps=max_sweeps)
        finally:
            eng.table_set_q(None)
        g = Growable.from_table(tab, out)
        return (g, st) if return_stats else g
    eng = engine_for(ctrl, device)
    out, st = eng.solve_gap(_dense(V, eng), _dense(nu, eng), Q=Q, tol=tol, max_sweeps=max_sweeps, precision=precision)
    g = Growable.from_dense(out, eng.walls)
    return (g, st) if return_stats else g


class McResult:
    """log Z estimates per (start cell, first action) and per start cell, with path counts."""

    def __init__(self, logz_sa, count_sa, logz_s, stats, walls):
        H, W = walls.shape
        self.logz_sa = logz_sa.reshape(W, H, 4).transpose(1, 0, 2)   # [i, j, a]
        self.count_sa = count_sa.reshape(W, H, 4).transpose(1, 0, 2)
        self.logz = Growable.from_dense(logz_s, walls)
        self.transitions = int(stats.transitions)
        self.paths = int(stats.paths)
        self.device_seconds = float(stats.device_seconds)


class TableMcResult:
    """mc_logz on a tabulated problem: estimates keyed by the problem's state objects."""

    def __init__(self, tab, sa, cnt, ls, stats):
        self.table, self.logz_sa, self.count_sa, self.stats = tab, sa, cnt, stats
        self.transitions = int(stats.transitions)
        self.logz = Growable.from_table(tab, ls)

    def logz_q(self, s, a):
        return float(self.logz_sa[self.table.index[s], list(self.table.actions).index(a)])


def mc_logz(ctrl, V, beta=1.0, β=None, Q=None, npaths=1 << 20, seed=0, estimator="A", horizon=1 << 20, starts=None,
            device=0):
    """Monte-Carlo estimate of generate_logz(ctrl, V; β) from rollouts under pi_Q.

    starts: list of (i, j) start states (path p starts at starts[p % len]); None = every free cell.
    """
    beta = beta if β is None else β
    if ctrl.grid is None:
        if estimator != "A":
            raise NotImplementedError("tabulated problems: the sampled-stop estimator only")
        eng, tab = table_for(ctrl, device)
        idx = None if starts is None else [tab.index[s] for s in starts]
        eng.table_set_q(_table_q(tab, Q))
        try:
            sa, cnt, ls, st = eng.table_mc_logz(_table_values(tab, V), beta=beta, npaths=npaths, seed=seed, horizon=horizon, starts=idx)
        finally:
            eng.table_set_q(None)
        return TableMcResult(tab, sa, cnt, ls, st)
    eng = engine_for(ctrl, device)
    cells = None if starts is None else [int(s[0]) + eng.H * int(s[1]) for s in starts]
    sa, cnt, s, st = eng.mc_logz(_dense(V, eng), beta=beta, Q=Q, npaths=npaths, seed=seed, estimator=estimator,
                                 horizon=horizon, starts=cells)
    return McResult(sa, cnt, s, st, eng.walls)


def mc_logz_updateV(ctrl, V, beta=1.0, β=None, npaths=1 << 22, niter=1, seed=0, estimator="A", horizon=1 << 20, device=0,
                    return_stats=False):
    """logz_updateV with the Monte-Carlo estimate of log Z in place of generate_logz's fixed point, niter times on the
    device: V <- V + log Z^/beta from rollouts under the soft policy of the current V (paths start at every free cell)."""
    beta = beta if β is None else β
    eng = engine_for(ctrl, device)
    out, st = eng.mc_update_v(_dense(V, eng), beta=beta, npaths=npaths, niter=niter, seed=seed, estimator=estimator,
                              horizon=horizon)
    g = Growable.from_dense(out, eng.walls, default=getattr(V, "default", lambda x: 0.0))
    return (g, st) if return_stats else g


# ---- experiment protocol (scripts/experiments.jl:10-39) -------------------------------------------------
def learning_traj(ctrl, v0, beta=1.0, n=20, **kw):
    out, vnew = [v0], v0
    for _ in range(n):
        vnew = logz_updateV(ctrl, vnew, beta=beta, **kw)
        out.append(vnew)
    return out


def learning_traj_batch(ctrl, v0, betas, n=5, tol=1e-12, max_sweeps=10 ** 7, precision="f64", device=0):
    """learning_traj for every beta in one library call (one CTA per chain): {beta: [V_0 .. V_n]}.
    Replaces the pmap over beta of scripts/distributed_experiments.jl:26."""
    eng = engine_for(ctrl, device)
    traj, _st = eng.learning_traj(_dense(v0, eng), [float(b) for b in betas], n=n, tol=tol, max_sweeps=max_sweeps,
                                  precision=precision)
    default = getattr(v0, "default", lambda x: 0.0)
    return {float(b): [Growable.from_dense(traj[i, k], eng.walls, default=default) for k in range(n + 1)]
            for i, b in enumerate(betas)}


def generate_optimality_gap_batch(ctrl, Vs, nu, tol=1e-12, max_sweeps=10 ** 7, precision="f64", device=0):
    """generate_optimality_gap for a list of value functions in one library call (one CTA per value function).
    Replaces the pmap over iterates of scripts/distributed_experiments.jl:28."""
    eng = engine_for(ctrl, device)
    gaps, _st = eng.solve_gap_batch(np.stack([_dense(v, eng) for v in Vs]), _dense(nu, eng), tol=tol,
                                    max_sweeps=max_sweeps, precision=precision)
    return [Growable.from_dense(gaps[i], eng.walls) for i in range(len(Vs))]


def run_experiment(gw, randomness=0.01, gamma=0.99, starting_temp=0.1, betarange=None, n=5, batched=True, **kw):
    """run_experiment (scripts/experiments.jl:20-39) == run_experiment_distributed (scripts/distributed_experiments.jl
    :15-33): nu, V_lo, 8 beta x n logz_updateV iterates and the optimality gap of every iterate.  batched=True sends
    the beta chains and the gap solves to the GPU as two batched calls instead of one call per solve."""
    betarange = np.exp(np.linspace(-6, 0, 8)) if betarange is None else betarange
    ctrl = gwmod.gridworld_ctrl(gw, randomness=randomness, gamma=gamma)
    weaker = gwmod.modify(ctrl, reward_function=lambda s, a, ss: starting_temp * ctrl.reward_function(s, a, ss))
    nu = generate_nu(ctrl, **kw)
    furthest = min(nu.dict, key=nu.dict.get)
    V_lo = generate_nu(weaker, **kw)
    if batched:
        traj = learning_traj_batch(ctrl, V_lo, betarange, n=n, **kw)
        flat = [v for b in traj for v in traj[b]]
        gaps = generate_optimality_gap_batch(ctrl, flat, nu, **kw)
        gap = {b: gaps[i * (n + 1):(i + 1) * (n + 1)] for i, b in enumerate(traj)}
    else:
        traj = {float(b): learning_traj(ctrl, V_lo, beta=float(b), n=n, **kw) for b in betarange}
        gap = {b: [generate_optimality_gap(ctrl, v, nu=nu, **kw) for v in vec] for b, vec in traj.items()}
    return dict(nu=nu, traj=traj, gap=gap, randomness=randomness, gamma=gamma, ctrl=ctrl, s0=furthest)


run_experiment_distributed = run_experiment
