"""ctypes front end of the CPU oracle (oracle/piql_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (piql.jl_b200/) never does.

Arrays follow the Julia layout of the reference: the wall array is (H, W) including the wall
border and is handed to C column-major, so cell (i, j) (0-based) has linear index i + H*j.
On the Python side every per-cell array is a numpy array of shape (H, W) in Fortran order.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpiql_oracle.so")

KIND = {"nu": 0, "logz": 1, "gap": 2, "z": 3}
MODE = {"jacobi": 0, "gauss_seidel": 1, "faithful": 2}


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "piql_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


class Stats(C.Structure):
    _fields_ = [("sweeps", C.c_int64), ("diff", C.c_double), ("seconds", C.c_double),
                ("backups", C.c_int64), ("nstates", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_kl_cost.restype = C.c_double
        _lib.orc_wall_unreachable.restype = C.c_int64
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _f(a, dtype=np.float64):
    return None if a is None else np.asfortranarray(a, dtype=dtype)


def _q(a):
    """Q table: shape (H*W, 4), action fastest (Julia Q[a, i, j])."""
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 4)


class Grid:
    """The tabulated gridworld_ctrl problem (src/Problems/GridWorld.jl:90-121)."""

    def __init__(self, walls, goal, randomness=0.1, reward_scale=0.2, step_cost=None, reward=None,
                 gamma=0.99):
        self.walls = np.asfortranarray(walls, dtype=np.uint8)
        self.H, self.W = self.walls.shape
        self.goal = (int(goal[0]), int(goal[1]))  # 0-based (i, j)
        self.goal_idx = self.goal[0] + self.H * self.goal[1]
        self.randomness = float(randomness)
        self.step_cost = 0.1 * reward_scale if step_cost is None else float(step_cost)
        self.reward = 1 * reward_scale if reward is None else float(reward)
        self.gamma = float(gamma)

    def _args(self, reward_scale=1.0):
        return (C.c_int64(self.H), C.c_int64(self.W), _p(self.walls, C.c_uint8), C.c_int64(self.goal_idx),
                C.c_double(self.randomness), C.c_double(self.step_cost), C.c_double(self.reward),
                C.c_double(self.gamma), C.c_double(reward_scale))

    def free_cells(self):
        return np.flatnonzero(self.walls.ravel(order="F") == 0).astype(np.int32)

    # ---- solvers -----------------------------------------------------------------------
    def solve(self, kind, V=None, Q=None, nu=None, beta=1.0, tol=1e-12, max_sweeps=10 ** 7, mode="jacobi",
              reward_scale=1.0, start=None, seed=0, nthreads=1):
        out = np.empty((self.H, self.W), dtype=np.float64, order="F")
        st = Stats()
        V, Q, nu = _f(V), _q(Q), _f(nu)
        if start is None:
            start = int(self.free_cells()[0])
        lib().orc_solve(C.c_int(KIND[kind]), C.c_int(MODE[mode]), *self._args(reward_scale),
                        _p(V, C.c_double), _p(Q, C.c_double), _p(nu, C.c_double), C.c_double(beta),
                        C.c_double(tol), C.c_int64(max_sweeps), C.c_int64(start), C.c_uint64(seed),
                        C.c_int(nthreads), _p(out, C.c_double), C.byref(st))
        return out, st

    def sweep(self, kind, X, V=None, Q=None, nu=None, beta=1.0, reward_scale=1.0):
        out = np.empty((self.H, self.W), dtype=np.float64, order="F")
        diff = C.c_double()
        V, Q, nu, X = _f(V), _q(Q), _f(nu), _f(X)
        lib().orc_sweep(C.c_int(KIND[kind]), *self._args(reward_scale), _p(V, C.c_double), _p(Q, C.c_double),
                        _p(nu, C.c_double), C.c_double(beta), _p(X, C.c_double), _p(out, C.c_double),
                        C.byref(diff))
        return out, diff.value

    def nu_jacobi_dense(self, sweeps, reward_scale=1.0, nthreads=1):
        """Exactly `sweeps` dense OpenMP Jacobi sweeps of the nu backup from zero (the optimised CPU variant)."""
        out = np.empty((self.H, self.W), dtype=np.float64, order="F")
        f = lib().orc_nu_jacobi_dense
        f.restype = C.c_double
        diff = f(*self._args(reward_scale), C.c_int64(sweeps), C.c_int(nthreads), _p(out, C.c_double))
        return out, diff

    def generate_nu(self, reward_scale=1.0, **kw):
        return self.solve("nu", reward_scale=reward_scale, **kw)[0]

    def generate_logz(self, V, beta=1.0, **kw):
        return self.solve("logz", V=V, beta=beta, **kw)[0]

    def logz_updateV(self, V, beta=1.0, **kw):
        logz = self.generate_logz(V, beta=beta, **kw)
        out = np.empty_like(logz)
        Vf = _f(V)
        lib().orc_update_v(C.c_int64(self.H * self.W), _p(self.walls, C.c_uint8), _p(Vf, C.c_double),
                           _p(logz, C.c_double), C.c_double(beta), _p(out, C.c_double))
        return out

    def generate_optimality_gap(self, V, nu, **kw):
        return self.solve("gap", V=V, nu=nu, **kw)[0]

    def kl_cost(self, V, s, Q=None, reward_scale=1.0):
        V, Q = _f(V), _q(Q)
        return lib().orc_kl_cost(*self._args(reward_scale), _p(V, C.c_double), _p(Q, C.c_double), C.c_int64(s))

    # ---- sampling ----------------------------------------------------------------------
    def policy_table(self, V, beta=1.0, Q=None, reward_scale=1.0, nthreads=1):
        V, Q = _f(V), _q(Q)
        tab = np.zeros((self.H * self.W, 8), dtype=np.float32)
        lib().orc_set_table_threads(C.c_int(nthreads))
        lib().orc_policy_table(*self._args(reward_scale), _p(V, C.c_double), _p(Q, C.c_double),
                               C.c_double(beta), _p(tab, C.c_float))
        return tab

    def rollout(self, table, Vgoal, starts, npaths, path0=0, seed=0, estimator="A", horizon=1 << 20,
                injected=None, trace_steps=0, nthreads=1):
        """table: the (H*W, 8) float32 policy records, or the (H*W, 4) uint32 16-byte records of large grids."""
        starts = np.ascontiguousarray(starts, dtype=np.int32)
        rec_bytes = 16 if table.dtype == np.uint32 else 32
        table = np.ascontiguousarray(table).view(np.float32) if rec_bytes == 16 else np.ascontiguousarray(table, dtype=np.float32)
        logw = np.empty(npaths, dtype=np.float32)
        meta = np.empty(npaths, dtype=np.uint32)
        inj_steps = 0
        if injected is not None:
            injected = np.ascontiguousarray(injected, dtype=np.float32)
            assert injected.shape[0] == npaths and injected.shape[2] == 4
            inj_steps = injected.shape[1]
        trace = np.full((npaths, trace_steps, 3), -1, dtype=np.int32) if trace_steps else None
        lib().orc_rollout(C.c_int64(self.H), C.c_int64(self.goal_idx), _p(table, C.c_float),
                          C.c_double(self.gamma), C.c_double(self.randomness), C.c_double(Vgoal),
                          C.c_int({"A": 0, "B": 1}[estimator]), C.c_int64(horizon), C.c_int64(path0),
                          C.c_int64(npaths), _p(starts, C.c_int32), C.c_int64(len(starts)), C.c_uint64(seed),
                          _p(injected, C.c_float), C.c_int64(inj_steps), _p(logw, C.c_float),
                          _p(meta, C.c_uint32), _p(trace, C.c_int32), C.c_int64(trace_steps), C.c_int(nthreads),
                          C.c_int(rec_bytes))
        return (logw, meta, trace) if trace_steps else (logw, meta)

    def kl_table(self, V, Q=None):
        """kl_cost per cell (path_functions.jl:17-20), rounded once to float like the sampler's records."""
        out = np.zeros(self.H * self.W, dtype=np.float32)
        for s in self.free_cells():
            out[s] = self.kl_cost(V, int(s), Q=Q)
        return out

    def path_estimates(self, table, kl, V, starts, npaths, nsteps=20, path0=0, seed=0, injected=None, trace=False):
        """fitness_reward_estimate / direct_reward_estimate (path_functions.jl:34-50) on the fp32 records:
        running totals fit[npaths, nsteps], dir[npaths, nsteps] (+ the (s, a, s') triples)."""
        starts = np.ascontiguousarray(starts, dtype=np.int32)
        table = np.ascontiguousarray(table, dtype=np.float32)
        kl = np.ascontiguousarray(kl, dtype=np.float32)
        V = _f(np.nan_to_num(np.asarray(V, dtype=np.float64)))
        fit = np.empty((npaths, nsteps), dtype=np.float32)
        dr = np.empty((npaths, nsteps), dtype=np.float32)
        inj_steps = 0
        if injected is not None:
            injected = np.ascontiguousarray(injected, dtype=np.float32)
            assert injected.shape[0] == npaths and injected.shape[1] >= nsteps and injected.shape[2] == 4
            inj_steps = injected.shape[1]
        tr = np.empty((npaths, nsteps, 3), dtype=np.int32) if trace else None
        lib().orc_path_estimates(C.c_int64(self.H), C.c_int64(self.goal_idx), _p(table, C.c_float), _p(kl, C.c_float),
                                 _p(V, C.c_double), C.c_double(self.randomness), C.c_double(self.step_cost),
                                 C.c_double(self.reward), C.c_int64(nsteps), C.c_int64(path0), C.c_int64(npaths),
                                 _p(starts, C.c_int32), C.c_int64(len(starts)), C.c_uint64(seed), _p(injected, C.c_float),
                                 C.c_int64(inj_steps), _p(fit, C.c_float), _p(dr, C.c_float), _p(tr, C.c_int32))
        return (fit, dr, tr) if trace else (fit, dr)

    def path_estimates_f64(self, V, triples, Q=None):
        """The same totals in Float64 through the reference's closures, along given (s, a, s') triples of ONE path."""
        V, Q = _f(np.nan_to_num(np.asarray(V, dtype=np.float64))), _q(Q)
        triples = np.ascontiguousarray(triples, dtype=np.int32)
        n = triples.shape[0]
        fit, dr = np.empty(n), np.empty(n)
        lib().orc_path_estimates_f64(*self._args(1.0), _p(V, C.c_double), _p(Q, C.c_double), _p(triples, C.c_int32),
                                     C.c_int64(n), _p(fit, C.c_double), _p(dr, C.c_double))
        return fit, dr

    def rollout_f64(self, V, start, injected, beta=1.0, Q=None, stop_rule=True, reward_scale=1.0):
        V, Q = _f(V), _q(Q)
        injected = np.ascontiguousarray(injected, dtype=np.float64)
        n = injected.shape[0]
        trace = np.full((n, 3), -1, dtype=np.int32)
        logw, ln = C.c_double(), C.c_int64()
        lib().orc_rollout_f64(*self._args(reward_scale), _p(V, C.c_double), _p(Q, C.c_double), C.c_double(beta),
                              C.c_int64(start), _p(injected, C.c_double), C.c_int64(n), C.c_int(int(stop_rule)),
                              _p(trace, C.c_int32), C.byref(logw), C.byref(ln))
        return trace, logw.value, ln.value


def reduce_paths(logw, meta, nstarts, path0=0):
    """(lse_max, lse_sum, count)[nstarts, 4] and the number of sampled transitions."""
    logw = np.ascontiguousarray(logw, dtype=np.float32)
    meta = np.ascontiguousarray(meta, dtype=np.uint32)
    mx = np.empty((nstarts, 4)); sm = np.empty((nstarts, 4)); cnt = np.empty((nstarts, 4), dtype=np.uint64)
    tr = C.c_uint64()
    lib().orc_reduce(C.c_int64(path0), C.c_int64(len(logw)), C.c_int64(nstarts), _p(logw, C.c_float),
                     _p(meta, C.c_uint32), _p(mx, C.c_double), _p(sm, C.c_double), _p(cnt, C.c_uint64),
                     C.byref(tr))
    return mx, sm, cnt, tr.value


def logz_from_stats(mx, sm, cnt):
    """log Z_hat(s0,a0) = max + log(sum) - log(count);  log Z_hat(s0) pools the four buckets."""
    with np.errstate(divide="ignore", invalid="ignore"):
        lsa = mx + np.log(sm) - np.log(cnt.astype(np.float64))
        m = mx.max(axis=1, keepdims=True)
        tot = (sm * np.exp(mx - m)).sum(axis=1)
        ls = m[:, 0] + np.log(tot) - np.log(cnt.sum(axis=1).astype(np.float64))
    return lsa, ls


def philox4x32_10(ctr, key):
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32); key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(ctr, C.c_uint32), _p(key, C.c_uint32), _p(out, C.c_uint32))
    return out


def wall_unreachable(walls, goal_idx):
    w = np.asfortranarray(walls, dtype=np.uint8).copy(order="F")
    n = lib().orc_wall_unreachable(C.c_int64(w.shape[0]), C.c_int64(w.shape[1]), _p(w, C.c_uint8),
                                   C.c_int64(goal_idx))
    return w, int(n)


def num_threads():
    """Host threads to use: the cores this process may run on.  (Not omp_get_max_threads: launchers such as torchrun export
    OMP_NUM_THREADS=1 to every rank, and the OpenMP loops here take their thread count as an explicit argument anyway.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or int(lib().orc_num_threads()))


def table_solve(tab, kind, V=None, nu=None, beta=1.0, reward_scale=1.0, tol=1e-12, max_sweeps=10 ** 7, x0=None, Q=None):
    """Soft VI on a tabulated ControlProblem (piql_jl_b200.tables.ProblemTable): the restatement in piql_oracle.c
    (orc_table_solve) of PIQL.jl:55-130 / value_iteration.jl:62-169 over CSR rows.  Returns (x, sweeps, residual)."""
    ns = len(tab.terminal)
    out = np.empty(ns, dtype=np.float64)
    sweeps, res = C.c_int64(), C.c_double()
    V = None if V is None else np.ascontiguousarray(V, dtype=np.float64)
    nu = None if nu is None else np.ascontiguousarray(nu, dtype=np.float64)
    x0 = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
    Q = None if Q is None else np.ascontiguousarray(Q, dtype=np.float64).reshape(ns, len(tab.actions))
    a = [np.ascontiguousarray(tab.row_ptr, dtype=np.int64), np.ascontiguousarray(tab.succ, dtype=np.int32),
         np.ascontiguousarray(tab.prob, dtype=np.float64), np.ascontiguousarray(tab.reward, dtype=np.float64),
         np.ascontiguousarray(tab.log_prior, dtype=np.float64), np.ascontiguousarray(tab.terminal, dtype=np.uint8)]
    rc = lib().orc_table_solve(C.c_int(KIND[kind]), C.c_int64(ns), C.c_int(len(tab.actions)), _p(a[0], C.c_int64),
                               _p(a[1], C.c_int32), _p(a[2], C.c_double), _p(a[3], C.c_double), _p(a[4], C.c_double),
                               _p(a[5], C.c_uint8), C.c_double(tab.gamma), C.c_double(reward_scale), _p(V, C.c_double),
                               _p(nu, C.c_double), C.c_double(beta), _p(x0, C.c_double), C.c_double(tol),
                               C.c_int64(max_sweeps), _p(out, C.c_double), C.byref(sweeps), C.byref(res), _p(Q, C.c_double))
    assert rc == 0
    return out, sweeps.value, res.value


def table_rollout(tab, cdf, bf, cum, vterm, npaths, starts=None, path0=0, seed=0, horizon=1 << 20):
    """Paths on a tabulated problem with the sampler's float tables (orc_table_rollout): (logw float32, meta uint32)."""
    ns = len(tab.terminal)
    logw, meta = np.empty(npaths, dtype=np.float32), np.empty(npaths, dtype=np.uint32)
    st = None if starts is None else np.ascontiguousarray(starts, dtype=np.int32)
    a = [np.ascontiguousarray(tab.row_ptr, dtype=np.int64), np.ascontiguousarray(tab.succ, dtype=np.int32),
         np.ascontiguousarray(tab.terminal, dtype=np.uint8), np.ascontiguousarray(cdf, dtype=np.float32),
         np.ascontiguousarray(bf, dtype=np.float32), np.ascontiguousarray(cum, dtype=np.float32),
         np.ascontiguousarray(vterm, dtype=np.float32)]
    lib().orc_table_rollout(C.c_int64(ns), C.c_int(len(tab.actions)), _p(a[0], C.c_int64), _p(a[1], C.c_int32), _p(a[2], C.c_uint8),
                            _p(a[3], C.c_float), _p(a[4], C.c_float), _p(a[5], C.c_float), _p(a[6], C.c_float),
                            C.c_double(tab.gamma), _p(st, C.c_int32), C.c_int64(ns if st is None else len(st)),
                            C.c_int64(path0), C.c_int64(npaths), C.c_uint64(seed), C.c_int64(horizon),
                            _p(logw, C.c_float), _p(meta, C.c_uint32))
    return logw, meta
