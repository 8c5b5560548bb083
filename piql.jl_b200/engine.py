"""numpy-level wrapper of one piql_ctx (one GPU).  Thin: argument marshalling and error mapping.

Per-cell arrays are (H, W) Fortran-ordered float64 (the Julia layout the C ABI expects);
per-(cell, action) arrays are (H*W, 4) C-ordered, i.e. action fastest.
"""
import ctypes as C

import numpy as np

from . import _capi as K


def _pd(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _cell(a, H, W, name):
    if a is None:
        return None
    a = np.asfortranarray(a, dtype=np.float64)
    if a.shape != (H, W):
        raise ValueError("%s must have shape (%d, %d), got %s" % (name, H, W, a.shape))
    return a


def _sa(a, H, W, name):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 4)
    if a.shape[0] != H * W:
        raise ValueError("%s must have H*W*4 entries" % name)
    return a


class Engine:
    def __init__(self, device=0):
        self._lib = K.load()
        self._ctx = C.c_void_p()
        rc = self._lib.piql_create(C.byref(self._ctx), int(device))
        if rc != K.OK:
            self._ctx = None
            raise K.PiqlError(rc, {K.ERR_NO_DEVICE: "no CUDA device visible (there is no CPU fallback)",
                                   K.ERR_UNSUPPORTED: "device is not sm_100 (B200)",
                                   K.ERR_ARG: "bad device id"}.get(rc, "piql_create failed"))
        self.device = int(device)
        self.H = self.W = 0
        self.walls = None
        self.grid_key = None

    # ---- plumbing ----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.piql_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, allow=()):
        if rc != K.OK and rc not in allow:
            raise K.PiqlError(rc, self._lib.piql_last_error(self._ctx).decode())
        return rc

    @property
    def stream(self):
        p = C.c_void_p()
        self._ck(self._lib.piql_stream(self._ctx, C.byref(p)))
        return p.value or 0

    @property
    def launches(self):
        n = C.c_int64()
        self._ck(self._lib.piql_launch_count(self._ctx, C.byref(n)))
        return n.value

    # ---- problem -----------------------------------------------------------------------------
    def load_grid(self, walls, goal, randomness, step_cost, reward, gamma):
        w = np.asfortranarray(walls, dtype=np.uint8)
        H, W = w.shape
        p = K.GridParams(float(randomness), float(step_cost), float(reward), float(gamma))
        self._ck(self._lib.piql_grid_load(self._ctx, w.ctypes.data_as(C.POINTER(C.c_uint8)), H, W, int(goal[0]),
                                          int(goal[1]), C.byref(p)))
        self.H, self.W, self.walls = H, W, w
        self.goal = (int(goal[0]), int(goal[1]))
        n = C.c_int64()
        self._ck(self._lib.piql_grid_free_count(self._ctx, C.byref(n)))
        self.nfree = n.value

    @staticmethod
    def _opts(tol, max_sweeps, precision, check_every):
        prec = K.PRECISIONS[precision]
        return K.SolveOpts(float(tol), int(max_sweeps), prec, int(check_every))

    def _out(self):
        return np.empty((self.H, self.W), dtype=np.float64, order="F")

    # ---- soft value iteration ----------------------------------------------------------------
    def solve_nu(self, reward_scale=1.0, tol=1e-12, max_sweeps=10 ** 7, precision="f64", check_every=0,
                 allow_unconverged=False):
        out, st = self._out(), K.SolveStats()
        o = self._opts(tol, max_sweeps, precision, check_every)
        self._ck(self._lib.piql_solve_nu(self._ctx, float(reward_scale), C.byref(o), _pd(out), C.byref(st)),
                 allow=(K.ERR_NOT_CONVERGED,) if allow_unconverged else ())
        return out, st

    def solve_logz(self, V, beta=1.0, Q=None, tol=1e-12, max_sweeps=10 ** 7, precision="f64", check_every=0,
                   allow_unconverged=False):
        V, Q = _cell(V, self.H, self.W, "V"), _sa(Q, self.H, self.W, "Q")
        out, st = self._out(), K.SolveStats()
        o = self._opts(tol, max_sweeps, precision, check_every)
        self._ck(self._lib.piql_solve_logz(self._ctx, _pd(V), _pd(Q), float(beta), C.byref(o), _pd(out), C.byref(st)),
                 allow=(K.ERR_NOT_CONVERGED,) if allow_unconverged else ())
        return out, st

    def update_v(self, V, beta=1.0, tol=1e-12, max_sweeps=10 ** 7, precision="f64", check_every=0,
                 allow_unconverged=False, want_logz=False):
        V = _cell(V, self.H, self.W, "V")
        out, st = self._out(), K.SolveStats()
        logz = self._out() if want_logz else None
        o = self._opts(tol, max_sweeps, precision, check_every)
        self._ck(self._lib.piql_update_v(self._ctx, _pd(V), float(beta), C.byref(o), _pd(out), _pd(logz), C.byref(st)),
                 allow=(K.ERR_NOT_CONVERGED,) if allow_unconverged else ())
        return (out, logz, st) if want_logz else (out, st)

    def solve_z(self, V, beta=1.0, Q=None, tol=1e-9, max_sweeps=10 ** 7, precision="f64", check_every=0,
                allow_unconverged=False):
        """generate_z (value_iteration.jl:42-51): the linear-space Z solve; the reference's stop rule is 1e-9."""
        V, Q = _cell(V, self.H, self.W, "V"), _sa(Q, self.H, self.W, "Q")
        out, st = self._out(), K.SolveStats()
        o = self._opts(tol, max_sweeps, precision, check_every)
        self._ck(self._lib.piql_solve_z(self._ctx, _pd(V), _pd(Q), float(beta), C.byref(o), _pd(out), C.byref(st)),
                 allow=(K.ERR_NOT_CONVERGED,) if allow_unconverged else ())
        return out, st

    def z_update_v(self, V, beta=1.0, tol=1e-9, max_sweeps=10 ** 7, precision="f64", check_every=0,
                   allow_unconverged=False):
        """z_updateV (value_iteration.jl:53-59): V + log(Z)/beta."""
        V = _cell(V, self.H, self.W, "V")
        out, st = self._out(), K.SolveStats()
        o = self._opts(tol, max_sweeps, precision, check_every)
        self._ck(self._lib.piql_z_update_v(self._ctx, _pd(V), float(beta), C.byref(o), _pd(out), None, C.byref(st)),
                 allow=(K.ERR_NOT_CONVERGED,) if allow_unconverged else ())
        return out, st

    def solve_gap(self, V, nu, Q=None, tol=1e-12, max_sweeps=10 ** 7, precision="f64", check_every=0,
                  allow_unconverged=False):
        V, nu, Q = _cell(V, self.H, self.W, "V"), _cell(nu, self.H, self.W, "nu"), _sa(Q, self.H, self.W, "Q")
        out, st = self._out(), K.SolveStats()
        o = self._opts(tol, max_sweeps, precision, check_every)
        self._ck(self._lib.piql_solve_gap(self._ctx, _pd(V), _pd(Q), _pd(nu), C.byref(o), _pd(out), C.byref(st)),
                 allow=(K.ERR_NOT_CONVERGED,) if allow_unconverged else ())
        return out, st

    def learning_traj(self, V0, betas, n=5, tol=1e-12, max_sweeps=10 ** 7, precision="f64", allow_unconverged=False):
        """All beta chains of learning_traj in one call: returns traj[nbeta, n+1, H, W] (F-order cells) and stats."""
        V0 = _cell(np.nan_to_num(np.asarray(V0, dtype=np.float64)), self.H, self.W, "V0")
        betas = np.ascontiguousarray(betas, dtype=np.float64)
        out = np.empty((len(betas), n + 1, self.W, self.H), dtype=np.float64)   # cell index i + H*j: (W, H) C-order
        st = (K.SolveStats * (len(betas) * n))()
        o = self._opts(tol, max_sweeps, precision, 0)
        self._ck(self._lib.piql_learning_traj(self._ctx, _pd(V0), _pd(betas), len(betas), int(n), C.byref(o), _pd(out), st),
                 allow=(K.ERR_NOT_CONVERGED,) if allow_unconverged else ())
        return out.transpose(0, 1, 3, 2), list(st)

    def solve_gap_batch(self, Vs, nu, tol=1e-12, max_sweeps=10 ** 7, precision="f64", allow_unconverged=False):
        """generate_optimality_gap for a stack of value functions Vs[m, H, W]; returns gaps[m, H, W] and stats."""
        Vs = np.nan_to_num(np.asarray(Vs, dtype=np.float64))
        m = Vs.shape[0]
        vin = np.ascontiguousarray(Vs.transpose(0, 2, 1))                       # (m, W, H) C-order = column-major cells
        nu = _cell(np.nan_to_num(np.asarray(nu, dtype=np.float64)), self.H, self.W, "nu")
        out = np.empty((m, self.W, self.H), dtype=np.float64)
        st = (K.SolveStats * m)()
        o = self._opts(tol, max_sweeps, precision, 0)
        self._ck(self._lib.piql_solve_gap_batch(self._ctx, _pd(vin), m, _pd(nu), C.byref(o), _pd(out), st),
                 allow=(K.ERR_NOT_CONVERGED,) if allow_unconverged else ())
        return out.transpose(0, 2, 1), list(st)

    # ---- tabulated (non-grid) problems ---------------------------------------------------------------
    def load_table(self, tab):
        """tab: tables.ProblemTable (CSR over (state, action))."""
        a = dict(row_ptr=np.ascontiguousarray(tab.row_ptr, dtype=np.int64), succ=np.ascontiguousarray(tab.succ, dtype=np.int32),
                 prob=np.ascontiguousarray(tab.prob, dtype=np.float64), reward=np.ascontiguousarray(tab.reward, dtype=np.float64),
                 log_prior=np.ascontiguousarray(tab.log_prior, dtype=np.float64),
                 terminal=np.ascontiguousarray(tab.terminal, dtype=np.uint8))
        p = K.TableProblem(len(tab.terminal), len(tab.actions), a["row_ptr"].ctypes.data_as(C.POINTER(C.c_int64)),
                           a["succ"].ctypes.data_as(C.POINTER(C.c_int32)), _pd(a["prob"]), _pd(a["reward"]), _pd(a["log_prior"]),
                           a["terminal"].ctypes.data_as(C.POINTER(C.c_uint8)), float(tab.gamma))
        self._ck(self._lib.piql_table_load(self._ctx, C.byref(p)))
        self.table_n = len(tab.terminal)
        self.table_na = len(tab.actions)
        self.table_key = tab.key()

    def table_set_q(self, Q=None):
        """Tabulated Q[(state, action)] for the table solves / sampler that follow; None: the V-determined Q."""
        q = None if Q is None else np.ascontiguousarray(Q, dtype=np.float64).reshape(self.table_n, self.table_na)
        self._ck(self._lib.piql_table_set_q(self._ctx, _pd(q)))

    def table_solve(self, kind, V=None, nu=None, beta=1.0, reward_scale=1.0, tol=1e-12, max_sweeps=10 ** 7, check_every=0,
                    allow_unconverged=False):
        kind = {"nu": K.KIND_NU, "logz": K.KIND_LOGZ, "gap": K.KIND_GAP}[kind]
        V = None if V is None else np.ascontiguousarray(V, dtype=np.float64)
        nu = None if nu is None else np.ascontiguousarray(nu, dtype=np.float64)
        out, st = np.empty(self.table_n, dtype=np.float64), K.SolveStats()
        o = self._opts(tol, max_sweeps, "f64", check_every)
        self._ck(self._lib.piql_table_solve(self._ctx, kind, _pd(V), _pd(nu), float(beta), float(reward_scale), C.byref(o),
                                            _pd(out), C.byref(st)),
                 allow=(K.ERR_NOT_CONVERGED,) if allow_unconverged else ())
        return out, st

    def table_update_v(self, V, beta=1.0, tol=1e-12, max_sweeps=10 ** 7, check_every=0, allow_unconverged=False):
        V = np.ascontiguousarray(V, dtype=np.float64)
        out, logz, st = np.empty(self.table_n, dtype=np.float64), np.empty(self.table_n, dtype=np.float64), K.SolveStats()
        o = self._opts(tol, max_sweeps, "f64", check_every)
        self._ck(self._lib.piql_table_update_v(self._ctx, _pd(V), float(beta), C.byref(o), _pd(out), _pd(logz), C.byref(st)),
                 allow=(K.ERR_NOT_CONVERGED,) if allow_unconverged else ())
        return out, logz, st

    def table_mc_logz(self, V, beta=1.0, npaths=1 << 20, nnz=None, want_tables=False, **kw):
        """Monte-Carlo log Z on the loaded tabulated problem: (logz_sa (ns, na), count (ns, na), logz_s (ns,), stats[, tables])."""
        V = np.ascontiguousarray(V, dtype=np.float64)
        o, _keep = self._mc_opts(npaths, **kw)
        ns = self.table_n
        na = self.table_na
        sa, cnt, s = np.empty((ns, na)), np.empty((ns, na), dtype=np.uint64), np.empty(ns)
        st = K.McStats()
        fp = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))
        cdf = bf = cum = None
        if want_tables:
            cdf, bf, cum = np.empty((ns, na), np.float32), np.empty((ns, na), np.float32), np.empty(int(nnz), np.float32)
        self._ck(self._lib.piql_table_mc_logz(self._ctx, _pd(V), float(beta), C.byref(o), _pd(sa),
                                              cnt.ctypes.data_as(C.POINTER(C.c_uint64)), _pd(s), C.byref(st), fp(cdf), fp(bf), fp(cum)))
        return (sa, cnt, s, st, (cdf, bf, cum)) if want_tables else (sa, cnt, s, st)

    def sweep_once(self, kind, x, V=None, Q=None, nu=None, beta=1.0, reward_scale=1.0, precision="f64"):
        kind = {"nu": K.KIND_NU, "logz": K.KIND_LOGZ, "gap": K.KIND_GAP, "z": K.KIND_Z}[kind]
        prec = K.PRECISIONS[precision]
        x = _cell(np.nan_to_num(np.asarray(x, dtype=np.float64)), self.H, self.W, "x")
        V, nu, Q = _cell(V, self.H, self.W, "V"), _cell(nu, self.H, self.W, "nu"), _sa(Q, self.H, self.W, "Q")
        out, res = self._out(), C.c_double()
        self._ck(self._lib.piql_sweep_once(self._ctx, kind, _pd(V), _pd(Q), _pd(nu), float(beta), float(reward_scale),
                                           prec, _pd(x), _pd(out), C.byref(res)))
        return out, res.value

    # device-resident sweeps (bench)
    def vi_prepare(self, kind, V=None, Q=None, nu=None, beta=1.0, reward_scale=1.0, precision="f32"):
        kind = {"nu": K.KIND_NU, "logz": K.KIND_LOGZ, "gap": K.KIND_GAP}[kind]
        prec = K.PRECISIONS[precision]
        V, nu, Q = _cell(V, self.H, self.W, "V"), _cell(nu, self.H, self.W, "nu"), _sa(Q, self.H, self.W, "Q")
        self._ck(self._lib.piql_vi_prepare_dev(self._ctx, kind, _pd(V), _pd(Q), _pd(nu), float(beta),
                                               float(reward_scale), prec))

    def vi_tile_plan(self, kind="nu", precision="f32"):
        out = (C.c_int32 * 8)()
        self._ck(self._lib.piql_vi_tile_plan(self._ctx, {"nu": K.KIND_NU, "logz": K.KIND_LOGZ, "gap": K.KIND_GAP}[kind],
                                             K.PRECISIONS[precision], out))
        d = dict(zip(("TJ", "warps", "tiles_i", "tiles_j", "ctas_per_sm", "smem", "sweeps_per_round", "persistent"), list(out)))
        flags = d["persistent"]
        d["persistent"], d["planes_in_smem"], d["column_resident"] = flags & 1, (flags >> 1) & 7, (flags >> 4) & 1
        return d

    def vi_set_iterate(self, x):
        x = _cell(np.nan_to_num(np.asarray(x, dtype=np.float64)), self.H, self.W, "x")
        self._ck(self._lib.piql_vi_set_iterate(self._ctx, _pd(x)))

    # multi-GPU soft VI (column slabs, halo exchange over NVLink peer memory)
    def vi_shard_begin(self, rank, world):
        h = K.ViShardHandles()
        self._ck(self._lib.piql_vi_shard_begin(self._ctx, int(rank), int(world), C.byref(h)))
        return bytes(h), (int(h.col0), int(h.col1))

    def vi_shard_connect(self, blobs):
        arr = (K.ViShardHandles * len(blobs))()
        for r, b in enumerate(blobs):
            C.memmove(C.addressof(arr[r]), b, C.sizeof(K.ViShardHandles))
        self._ck(self._lib.piql_vi_shard_connect(self._ctx, arr))

    def vi_shard_end(self):
        self._ck(self._lib.piql_vi_shard_end(self._ctx))

    def vi_shard_wait_stats(self):
        """(microseconds edge CTAs waited for halos, number of waits) since the last call."""
        us, n = C.c_double(), C.c_int64()
        self._ck(self._lib.piql_vi_shard_wait_stats(self._ctx, C.byref(us), C.byref(n)))
        return us.value, n.value

    def vi_sweeps(self, sweeps):
        self._ck(self._lib.piql_vi_sweeps_dev(self._ctx, int(sweeps)))

    def vi_fetch(self, want_x=True):
        out, res = (self._out() if want_x else None), C.c_double()
        self._ck(self._lib.piql_vi_fetch(self._ctx, _pd(out), C.byref(res)))
        return out, res.value

    # ---- Monte-Carlo -------------------------------------------------------------------------
    def policy_table(self, V, beta=1.0, Q=None, download=True):
        V, Q = _cell(V, self.H, self.W, "V"), _sa(Q, self.H, self.W, "Q")
        tab = np.empty((self.H * self.W, 8), dtype=np.float32) if download else None
        self._ck(self._lib.piql_policy_table(self._ctx, _pd(V), _pd(Q), float(beta),
                                             None if tab is None else tab.ctypes.data_as(C.POINTER(C.c_float))))
        return tab

    def policy_table16(self, download=True):
        """The 16-byte policy records the production rollout reads on large grids: (table (H*W, 4) uint32 or None, in_use)."""
        use = C.c_int32()
        self._ck(self._lib.piql_policy_table16(self._ctx, None, C.byref(use)))
        if not use.value or not download:
            return None, bool(use.value)
        tab = np.empty((self.H * self.W, 4), dtype=np.uint32)
        self._ck(self._lib.piql_policy_table16(self._ctx, tab.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(use)))
        return tab, True

    def policy_table_async(self, V, beta=1.0, Q=None):
        """piql_policy_table_async: upload + build on the library's upload stream, overlapping work already enqueued.  V (and
        Q) must not be modified until a later blocking call returns; they are kept alive here until the next call."""
        V, Q = _cell(V, self.H, self.W, "V"), _sa(Q, self.H, self.W, "Q")
        self._table_keep = (V, Q)
        self._ck(self._lib.piql_policy_table_async(self._ctx, _pd(V), _pd(Q), float(beta)))

    def policy_table_set(self, table, v_goal):
        table = np.ascontiguousarray(table, dtype=np.float32)
        assert table.size == self.H * self.W * 8
        self._ck(self._lib.piql_policy_table_set(self._ctx, table.ctypes.data_as(C.POINTER(C.c_float)), float(v_goal)))

    def _mc_opts(self, npaths, path0=0, seed=0, estimator="A", horizon=1 << 20, starts=None, injected=None):
        keep = []
        o = K.McOpts()
        o.npaths, o.path0, o.seed = int(npaths), int(path0), int(seed)
        o.estimator = {"A": K.EST_SAMPLED_STOP, "B": K.EST_DISCOUNTED}[estimator]
        o.horizon = int(horizon)
        if starts is None:
            o.start_mode, o.starts, o.nstarts = K.START_ALL_FREE, None, 0
        else:
            s = np.ascontiguousarray(starts, dtype=np.int32)
            keep.append(s)
            o.start_mode, o.starts, o.nstarts = K.START_LIST, s.ctypes.data_as(C.POINTER(C.c_int32)), len(s)
        if injected is not None:
            u = np.ascontiguousarray(injected, dtype=np.float32)
            assert u.ndim == 3 and u.shape[0] == npaths and u.shape[2] == 4
            keep.append(u)
            o.injected, o.injected_steps = u.ctypes.data_as(C.POINTER(C.c_float)), u.shape[1]
        else:
            o.injected, o.injected_steps = None, 0
        return o, keep

    def mc_logz(self, V, beta=1.0, Q=None, npaths=1 << 20, want=("sa", "count", "s"), **kw):
        V, Q = _cell(V, self.H, self.W, "V"), _sa(Q, self.H, self.W, "Q")
        o, _keep = self._mc_opts(npaths, **kw)
        n = self.H * self.W
        sa = np.empty((n, 4), dtype=np.float64) if "sa" in want else None
        cnt = np.empty((n, 4), dtype=np.uint64) if "count" in want else None
        s = np.empty((self.H, self.W), dtype=np.float64, order="F") if "s" in want else None
        st = K.McStats()
        self._ck(self._lib.piql_mc_logz(self._ctx, _pd(V), _pd(Q), float(beta), C.byref(o), _pd(sa),
                                        None if cnt is None else cnt.ctypes.data_as(C.POINTER(C.c_uint64)), _pd(s),
                                        C.byref(st)))
        return sa, cnt, s, st

    def mc_update_v(self, V, beta=1.0, npaths=1 << 20, niter=1, **kw):
        """niter Monte-Carlo logz_updateV iterations on the device; returns (V_new (H, W) with NaN at walls, stats)."""
        V = _cell(np.nan_to_num(np.asarray(V, dtype=np.float64)), self.H, self.W, "V")
        o, _keep = self._mc_opts(npaths, **kw)
        out, st = self._out(), K.McStats()
        self._ck(self._lib.piql_mc_update_v(self._ctx, _pd(V), float(beta), C.byref(o), int(niter), _pd(out), C.byref(st)))
        return out, st

    def trace_paths(self, npaths, max_steps, **kw):
        o, _keep = self._mc_opts(npaths, **kw)
        tr = np.empty((npaths, max_steps, 3), dtype=np.int32)
        logw = np.empty(npaths, dtype=np.float32)
        meta = np.empty(npaths, dtype=np.uint32)
        self._ck(self._lib.piql_trace_paths(self._ctx, C.byref(o), int(max_steps), tr.ctypes.data_as(C.POINTER(C.c_int32)),
                                            logw.ctypes.data_as(C.POINTER(C.c_float)),
                                            meta.ctypes.data_as(C.POINTER(C.c_uint32))))
        return tr, logw, meta

    def path_estimates(self, V, npaths, nsteps=20, Q=None, want_trace=False, want_tables=False, **kw):
        """fitness_reward_estimate / direct_reward_estimate (path_functions.jl:34-50) for npaths paths of nsteps
        transitions: running totals fit[npaths, nsteps], dir[npaths, nsteps] (float32)."""
        V, Q = _cell(V, self.H, self.W, "V"), _sa(Q, self.H, self.W, "Q")
        o, _keep = self._mc_opts(npaths, **kw)
        fit = np.empty((npaths, nsteps), dtype=np.float32)
        dr = np.empty((npaths, nsteps), dtype=np.float32)
        tr = np.empty((npaths, nsteps, 3), dtype=np.int32) if want_trace else None
        tab = np.empty((self.H * self.W, 8), dtype=np.float32) if want_tables else None
        kl = np.empty(self.H * self.W, dtype=np.float32) if want_tables else None
        pf = lambda a, t=C.c_float: None if a is None else a.ctypes.data_as(C.POINTER(t))
        self._ck(self._lib.piql_path_estimates(self._ctx, _pd(V), _pd(Q), C.byref(o), int(nsteps), pf(fit), pf(dr),
                                               pf(tr, C.c_int32), pf(tab), pf(kl)))
        out = [fit, dr]
        if want_trace:
            out.append(tr)
        if want_tables:
            out += [tab, kl]
        return tuple(out)

    # device-resident pieces (bench, multi-GPU)
    def mc_reset(self, starts=None):
        o, _keep = self._mc_opts(0, starts=starts)
        self._ck(self._lib.piql_mc_reset_dev(self._ctx, C.byref(o)))

    def mc_rollout(self, npaths, **kw):
        o, _keep = self._mc_opts(npaths, **kw)
        self._ck(self._lib.piql_mc_rollout_dev(self._ctx, C.byref(o)))

    def mc_stats_ptrs(self):
        m, s, c, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
        self._ck(self._lib.piql_mc_stats_dev(self._ctx, C.byref(m), C.byref(s), C.byref(c), C.byref(n)))
        return m.value, s.value, c.value, n.value

    def mc_rescale(self, global_max_ptr):
        self._ck(self._lib.piql_mc_rescale_dev(self._ctx, C.c_void_p(global_max_ptr)))

    # the exchange step over NVLink peer memory (one process per GPU)
    def peer_export(self):
        h = K.PeerHandles()
        self._ck(self._lib.piql_peer_export(self._ctx, C.byref(h)))
        return bytes(h)

    def peer_import(self, rank, world, blobs):
        arr = (K.PeerHandles * world)()
        for r, b in enumerate(blobs):
            C.memmove(C.addressof(arr[r]), b, C.sizeof(K.PeerHandles))
        self._ck(self._lib.piql_peer_import(self._ctx, int(rank), int(world), arr))

    def mc_allreduce_peer(self):
        self._ck(self._lib.piql_mc_allreduce_peer(self._ctx))

    def peer_close(self):
        self._ck(self._lib.piql_peer_close(self._ctx))

    def mc_finalize(self, want=("sa", "count", "s"), out=None):
        """out: optional preallocated (sa (n,4) f64, count (n,4) u64, s (H,W) f64 F-order), e.g. pinned."""
        n = self.H * self.W
        if out is not None:
            sa, cnt, s = out
        else:
            sa = np.empty((n, 4), dtype=np.float64) if "sa" in want else None
            cnt = np.empty((n, 4), dtype=np.uint64) if "count" in want else None
            s = np.empty((self.H, self.W), dtype=np.float64, order="F") if "s" in want else None
        st = K.McStats()
        self._ck(self._lib.piql_mc_finalize(self._ctx, _pd(sa),
                                            None if cnt is None else cnt.ctypes.data_as(C.POINTER(C.c_uint64)), _pd(s),
                                            C.byref(st)))
        return sa, cnt, s, st

    def mc_finalize_async(self, out):
        """Snapshot the accumulators and finalise + download them on the library's copy stream while this context's stream
        goes on.  out = (sa (n,4), count (n,4), s (H,W) F-order), each None or preallocated (pinned for a truly asynchronous
        copy): float64 / uint64 arrays, or float32 / uint32 for half the PCIe bytes.  Call mc_finalize_wait() before reading."""
        sa, cnt, s = out
        kinds = {a.dtype.itemsize for a in (sa, cnt, s) if a is not None}
        if len(kinds) > 1 or (kinds and next(iter(kinds)) not in (4, 8)):
            raise ValueError("outputs must be all float64/uint64 or all float32/uint32")
        f32 = bool(kinds) and next(iter(kinds)) == 4          # all None: only the statistics come back
        ptr = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
        self._async_keep = out
        self._ck(self._lib.piql_mc_finalize_async(self._ctx, ptr(sa), ptr(cnt), ptr(s), 1 if f32 else 0))

    def mc_finalize_wait(self):
        st = K.McStats()
        self._ck(self._lib.piql_mc_finalize_wait(self._ctx, C.byref(st)))
        out, self._async_keep = getattr(self, "_async_keep", None), None
        return out, st

    def mc_timing(self):
        ro, re, n = C.c_double(), C.c_double(), C.c_int64()
        self._ck(self._lib.piql_mc_timing(self._ctx, C.byref(ro), C.byref(re), C.byref(n)))
        return ro.value, re.value, n.value

    def mc_transitions_total(self):
        n = C.c_uint64()
        self._ck(self._lib.piql_mc_transitions_total(self._ctx, C.byref(n)))
        return n.value

    def mc_records(self, n):
        logw = np.empty(n, dtype=np.float32)
        meta = np.empty(n, dtype=np.uint32)
        self._ck(self._lib.piql_mc_records(self._ctx, logw.ctypes.data_as(C.POINTER(C.c_float)),
                                           meta.ctypes.data_as(C.POINTER(C.c_uint32)), int(n)))
        return logw, meta
