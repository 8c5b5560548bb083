#!/usr/bin/env python3
"""Benchmark of the PIQL hot path on B200: sampled transitions/s (+ soft-VI sweeps/s).

    python bench.py --gpus N --steps K --warmup W              # this build (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU restatement (oracle)

A step is one iteration of the Monte-Carlo hot path on one batch of synthetic trajectories:
reset the per-(cell, action) accumulators, roll out the rank's block of path ids under the soft
policy of V, reduce the per-path records with log-sum-exp, and (N > 1) all-reduce the accumulators.
Workloads (BASELINE.json configs 4 and 5): synthetic gridworld of 1024^2 (N = 1) or 4096^2 (N > 1)
interior cells, wall density 0.2, randomness 0.01, gamma 0.99, beta 1; V is 1000 Jacobi sweeps of
the soft Bellman backup with rewards x 0.1 (the experiment protocol's starting V_lo); paths start
at the free cells in cyclic order; estimator A (sampled stop).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (interior size, numpy Philox seed, paths per step [whole job])
    "grid1024": dict(size=1024, seed=1234, paths=1 << 27),
    "grid4096": dict(size=4096, seed=5678, paths=1 << 27),
    "grid256": dict(size=256, seed=99, paths=1 << 22),   # quick functional check
}
RANDOMNESS, GAMMA, BETA, DENSITY = 0.01, 0.99, 1.0, 0.2
REWARD_SCALE = 0.2
VLO_SCALE, VLO_SWEEPS = 0.1, 1000
MC_SEED = 20261019


def make_world(name):
    import piql_jl_b200 as piql
    w = WORKLOADS[name]
    gw = piql.make_gridworld((w["size"], w["size"]), density=DENSITY, rng=np.random.Generator(np.random.Philox(w["seed"])))
    walls = np.asfortranarray(gw.walls, dtype=np.uint8)
    params = dict(randomness=RANDOMNESS, step_cost=0.1 * REWARD_SCALE, reward=1 * REWARD_SCALE, gamma=GAMMA)
    return walls, tuple(int(v) for v in gw.goal), params


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("PIQL_BENCH_SMI_MS", "100")],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx = float(f[1])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def pinned(shape, dtype, order="C"):
    import torch
    n = int(np.prod(shape))
    t = torch.empty(n * np.dtype(dtype).itemsize, dtype=torch.uint8, pin_memory=True)
    a = t.numpy().view(dtype)
    return a.reshape(shape, order=order), t   # keep t alive as long as the view is used


# ------------------------------------------------------------------------------------------------
def sample_starts(free, path0, sample):
    """Start list for a bounded sample of consecutive path ids [path0, path0 + sample): the step's paths start at
    free[p % nfree]; a sample shorter than the free-cell cycle touches only `sample` consecutive slots of it, so the CPU
    arm reduces into those (a rotation keeps slot = p % nstarts) instead of initialising all 4 x nfree buckets per step."""
    nfree = len(free)
    if sample >= nfree:
        return free
    p = path0 + np.arange(sample, dtype=np.int64)
    starts = np.empty(sample, dtype=np.int32)
    starts[p % sample] = free[p % nfree]
    return starts


def run_reference(args, rank, world):
    """The reference's algorithm on the host cores: the C restatement (oracle/), all threads."""
    if rank != 0:
        return
    import oracle
    name = args.workload
    walls, goal, params = make_world(name)
    org = oracle.Grid(walls, goal, **params)
    threads = oracle.num_threads()
    # V_lo as in the GPU arm (VLO_SWEEPS dense Jacobi sweeps); at 4096^2 the host needs ~0.4 s per sweep on 8 cores, so the
    # bounded run builds it from 64 sweeps there: the policy away from the goal is the same near-uniform one either way and
    # the path-length distribution (set by gamma) and the cost per transition do not depend on how converged V is
    ref_sweeps = min(args.setup_sweeps, 64) if name == "grid4096" else args.setup_sweeps
    V, _ = org.nu_jacobi_dense(ref_sweeps, reward_scale=VLO_SCALE, nthreads=threads)
    V = np.nan_to_num(V)
    table = org.policy_table(V, beta=BETA, nthreads=threads)
    free = org.free_cells()
    sample = args.ref_paths
    vgoal = V[goal]
    done = 0

    def step():
        nonlocal done
        starts = sample_starts(free, done, sample)
        lw, meta = org.rollout(table, vgoal, starts, sample, path0=done, seed=MC_SEED, nthreads=threads)
        mx, sm, cnt, tr = oracle.reduce_paths(lw, meta, len(starts), path0=done)
        done += sample
        return tr
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    transitions = sum(step() for _ in range(args.steps))
    dt = time.perf_counter() - t0
    value = transitions / dt
    cfg = workload_config(name, args, world)
    cfg["paths_per_step"] = int(sample)
    cfg["sharding"] = "one host process, %d threads" % threads
    cfg["v_lo_sweeps"] = int(ref_sweeps)
    cfg["sample"] = ("each step rolls out %d consecutive path ids of the GPU arm's %d-path step and reduces them into the "
                     "(start cell, action) buckets they touch" % (sample, WORKLOADS[name]["paths"]))
    line = {
        "impl": "reference", "metric": "sampled transitions/sec", "value": value, "unit": "transitions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "transitions/s", "cores": threads, "kind": "port",
                         "sample": "%d paths per step (of the %d-path step per GPU), oracle/piql_oracle.c with OpenMP over paths, V_lo from "
                                   "%d Jacobi sweeps; Julia is not installed, so the reference itself cannot run"
                                   % (sample, WORKLOADS[name]["paths"], ref_sweeps)},
        "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "transitions_per_step": transitions / args.steps,
    }
    print(json.dumps(line), flush=True)


def run_golden_experiment(device, with_cpu=True, name="r0_01g0_99"):
    """run_experiment of the reference (scripts/experiments.jl:20-39) on its own 20x20 gridworld, Float64, host buffers in
    and out: nu, V_lo, 8 beta x 5 logz_updateV, 48 optimality gaps; checked against the stored data/*.jld2 values."""
    import piql_jl_b200 as piql
    d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    walls = np.asfortranarray(d["walls"])
    goal = tuple(int(v) - 1 for v in d["goal"])
    params = dict(randomness=float(d["randomness"]), step_cost=float(d["step_cost"]), reward=float(d["reward"]), gamma=float(d["gamma"]))
    betas = [float(b) for b in d["betas"]]
    states = d["states"] - 1
    eng = piql.Engine(device)
    eng.load_grid(walls, goal, **params)

    def once():
        nu, s1 = eng.solve_nu()
        vlo, s2 = eng.solve_nu(reward_scale=0.1)
        traj, st = eng.learning_traj(np.nan_to_num(vlo), betas, n=5)
        gaps, gst = eng.solve_gap_batch(np.nan_to_num(traj.reshape(48, *walls.shape)), np.nan_to_num(nu))
        return nu, traj, gaps, s1.sweeps + s2.sweeps + sum(x.sweeps for x in st) + sum(x.sweeps for x in gst)
    once()
    t0 = time.perf_counter()
    nu, traj, gaps, sweeps = once()
    sec = time.perf_counter() - t0
    i, j = states[:, 0], states[:, 1]
    err = max(float(np.abs(nu[i, j] - d["nu"]).max()),
              float(np.abs(traj[7, 5][i, j] - d["traj"][7, 5]).max()), float(np.abs(gaps.reshape(8, 6, *walls.shape)[7, 1][i, j] - d["gap"][7, 1]).max()))
    out = {"config": "golden gridworld %s (BASELINE configs 1-2), f64, tol 1e-12" % name, "solves": "2 nu + 40 logz_updateV + 48 optimality gap",
           "seconds": sec, "sweeps": int(sweeps), "state_backups_per_s": sweeps * eng.nfree / sec,
           "max_abs_err_vs_reference_data": err, "call": "piql_solve_nu x2 + piql_learning_traj + piql_solve_gap_batch"}
    eng.close()
    if with_cpu:
        import oracle
        org = oracle.Grid(walls, goal, **params)
        V = np.nan_to_num(traj[7, 0])
        t0 = time.perf_counter()
        _x, a = org.solve("logz", V=V, beta=1.0, mode="faithful")
        t1 = time.perf_counter()
        _y, b = org.solve("gap", V=np.nan_to_num(traj[7, 1]), nu=np.nan_to_num(nu), mode="faithful")
        t2 = time.perf_counter()
        out["cpu_baseline"] = {"kind": "port", "cores": 1, "unit": "s",
                               "sample": "1 of the 40 logz solves (beta=1, k=0) and 1 of the 48 gap solves (beta=1, k=1), oracle 'faithful' mode "
                                         "(hash-map Growable, shuffled Gauss-Seidel, nested closure recomputation), one thread = one Julia worker",
                               "logz_solve_seconds": t1 - t0, "logz_sweeps": int(a.sweeps), "gap_solve_seconds": t2 - t1, "gap_sweeps": int(b.sweeps),
                               "state_backups_per_s": (a.sweeps + b.sweeps) * len(states) / (t2 - t0)}
    return out


def run_extra(name, fn):
    """The sections after the headline (sweep rates, CPU baseline, golden experiment) must not be able to lose the line."""
    try:
        return fn()
    except Exception as e:  # reported in the line instead
        return {"error": "%s: %s" % (type(e).__name__, e)}


def vi_traffic(name, key):
    """ncu DRAM bytes (read + write) of one launch of the tiled sweep kernel, from profiles/vi_traffic.json
    ({"grid1024:f32fast": {"bytes": ..., "sweeps_per_launch": 4}, ...}); None when that case has not been captured."""
    path = os.path.join(ROOT, "profiles", "vi_traffic.json")
    try:
        return json.load(open(path)).get("%s:%s" % (name, key))
    except Exception:
        return None


def table_section(device, with_cpu=True):
    """The tabulated-problem path on the reference's BinaryBandit (30 rounds, 2 coins: 46,376 belief states, 163,680 transitions):
    Float64 solves of nu and of one logz_updateV, then Monte-Carlo path sampling on the tables.  Functional numbers: the table
    kernels are one thread per state in Float64 (table.cu), not tuned like the gridworld path."""
    import piql_jl_b200 as piql
    bb = piql.binary_bandit_problem(rounds=30, ncoins=2, payoff=0.8, gamma=0.98)
    t0 = time.perf_counter()
    tab = piql.tabulate_problem(bb)
    t_tab = time.perf_counter() - t0
    eng = piql.Engine(device)
    try:
        eng.load_table(tab)
        eng.table_solve("nu", max_sweeps=8, allow_unconverged=True)          # warm-up
        t0 = time.perf_counter()
        nu, st = eng.table_solve("nu", tol=1e-9)
        t_nu = time.perf_counter() - t0
        t0 = time.perf_counter()
        v1, logz, st1 = eng.table_update_v(0.5 * nu, beta=1.0, tol=1e-9)   # (46,376 states: the rounding floor of sum |delta| is ~1e-11)
        t_up = time.perf_counter() - t0
        npaths = 1 << 22
        s0 = tab.index[bb.initial_state()]
        eng.table_mc_logz(0.5 * nu, beta=1.0, npaths=1 << 16, seed=3, starts=[s0])       # warm-up
        t0 = time.perf_counter()
        sa, cnt, ls, mst = eng.table_mc_logz(0.5 * nu, beta=1.0, npaths=npaths, seed=4, starts=[s0])
        t_mc = time.perf_counter() - t0
        out = {"problem": "binary_bandit_problem(rounds=30, ncoins=2, payoff=0.8, gamma=0.98) tabulated through its closures",
               "states": len(tab.states), "transitions_nnz": int(len(tab.succ)), "tabulation_seconds_host": t_tab,
               "tol_sum_abs_delta": 1e-9, "solve_nu": {"sweeps": int(st.sweeps), "converged": bool(st.converged), "seconds": t_nu,
                            "state_backups_per_s": st.sweeps * len(tab.states) / t_nu},
               "logz_update_v": {"sweeps": int(st1.sweeps), "converged": bool(st1.converged), "seconds": t_up},
               "mc_logz": {"paths": npaths, "starts": "the initial belief state", "transitions": float(mst.transitions), "seconds": t_mc,
                           "transitions_per_s": float(mst.transitions) / t_mc,
                           "logz_s0_mc": float(ls[s0]), "logz_s0_exact": float(logz[s0])},
               "call": "piql_table_load + piql_table_solve + piql_table_update_v + piql_table_mc_logz (host buffers in and out)"}
        if with_cpu:
            import oracle
            t0 = time.perf_counter()
            ref, sw, _ = oracle.table_solve(tab, "nu", tol=1e-9)
            t_ref = time.perf_counter() - t0
            out["cpu_baseline"] = {"kind": "port", "cores": 1, "solve_nu_seconds": t_ref, "sweeps": int(sw),
                                   "max_abs_diff_vs_gpu": float(np.abs(nu - ref).max())}
        return out
    finally:
        eng.close()


def vi_section(args, eng, stream, V, peak, vst, name=None):
    """Soft-VI sweep rates on the loaded grid (rank 0): nu in three precisions, log Z and gap in fast fp32, the ground-truth solve."""
    import torch
    vi = {}
    cases = [("nu", "f32fast", 8, {}), ("nu", "f32", 8, {}), ("nu", "f64", 16, {}),
             ("logz", "f32fast", 12, dict(V=V, beta=BETA)), ("gap", "f32fast", 16, dict(V=V, nu=V))]
    for kind, prec, nbytes, kw in cases:
        eng.vi_prepare(kind, reward_scale=1.0, precision=prec, **kw)
        with torch.cuda.stream(stream):
            eng.vi_sweeps(20)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            eng.vi_sweeps(args.vi_sweeps)
            b.record(stream)
            torch.cuda.synchronize()
        sec = a.elapsed_time(b) * 1e-3
        sps = args.vi_sweeps / sec
        gbs = nbytes * eng.nfree * sps / 1e9
        key = prec if kind == "nu" else "%s_%s" % (kind, prec)
        if key == "f32fast":
            vi["tile_plan"] = eng.vi_tile_plan("nu", "f32fast")   # the shape the library timed fastest on this device
        plan = eng.vi_tile_plan(kind, prec)
        kh = max(int(plan.get("sweeps_per_round", 4)), 1)
        if plan.get("column_resident"):
            kh = args.vi_sweeps          # the column-resident schedule runs the whole batch of sweeps in one launch
        tr = vi_traffic(name, key)
        vi[key] = {"sweeps_per_s": sps, "state_backups_per_s": sps * eng.nfree, "us_per_sweep": 1e6 / sps,
                   "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                "bytes_per_state": nbytes, "sweeps_per_launch": kh,
                                "algorithmic_bytes_per_launch": nbytes * eng.nfree * kh,
                                "avg_launch_us": kh * 1e6 / sps,
                                # ncu DRAM bytes of the captured launch, scaled to a round of kh sweeps like the algorithmic figure
                                "traffic": None if tr is None else tr["bytes"] * kh / max(tr.get("sweeps_in_launch", kh), 1),
                                "traffic_note": None if tr is None else "%s; kernel %s, %d sweeps in the captured launch" % (
                                    tr.get("note"), tr.get("kernel"), tr.get("sweeps_in_launch", kh))},
                   "tile_plan": plan}
    # the ground-truth solve itself (config 4: "soft VI ground truth on 1 B200"), host buffers in and out
    solves = {}
    for prec, tol in (("f64", 1e-9), ("f32fast", 0.05)):
        t0 = time.perf_counter()
        _nu, sst = eng.solve_nu(tol=tol, max_sweeps=20000, precision=prec, allow_unconverged=True)
        solves[prec] = {"tol_sum_abs_delta": tol, "sweeps": int(sst.sweeps), "converged": bool(sst.converged),
                        "seconds": time.perf_counter() - t0, "device_seconds": float(sst.device_seconds)}
    vi["solve_nu"] = solves
    vi["kernel"] = ("per (kind, precision) the schedule the library timed fastest on this device (tile_plan): tile_sweep_kernel<T,KIND,FAST> "
                    "= 4 or 8 Jacobi sweeps per launch on 128 x (TJ + 2 kh) shared-memory tiles; col_resident_kernel<T,KIND,FAST> "
                    "(tile_plan.column_resident = 1) = one CTA per SM keeps a slab of whole columns in shared memory for the whole batch of "
                    "sweeps and exchanges one column per sweep with each neighbour through L2 mailboxes, every cell computed once per sweep; "
                    "timings include the residual reduction of the last sweep; f32fast = MUFU ex2/lg2; keys without a kind are nu")
    vi["setup_vlo"] = {"sweeps": int(vst.sweeps), "device_seconds": float(vst.device_seconds)}
    return vi


class OracleSide:
    """The CPU oracle on this rank's problem, reading the policy table the GPU built (test infrastructure: the checker and the
    cpu_baseline leg, never the thing measured)."""

    def __init__(self, eng, walls, goal, params, V, threads=None, table=None):
        import oracle
        self.oracle = oracle
        self.org = oracle.Grid(walls, goal, **params)
        self.threads = threads or oracle.num_threads()
        self.table = eng.policy_table(V, beta=BETA) if table is None else table   # the fp32 table the GPU read
        self.record_bytes = 32
        if table is None:
            t16, in_use = eng.policy_table16()
            if in_use:                       # 4096^2: the production kernel reads the 16-byte form of the same records
                self.table, self.record_bytes = t16, 16
        self.free = self.org.free_cells()
        self.vgoal = float(V[goal])

    def records(self, path0, n):
        return self.org.rollout(self.table, self.vgoal, self.free, n, path0=path0, seed=MC_SEED, nthreads=self.threads)

    def slot_stats(self, slots, id0, id1):
        """(max, sum, count)[len(slots), 4] and log Z^(s0, a0) over the path ids of [id0, id1) that start in the given slots."""
        nfree = len(self.free)
        lw, me, owner = [], [], []
        for k, sl in enumerate(slots):
            first = id0 + ((int(sl) - id0) % nfree)
            for pid in range(first, id1, nfree):
                a, b = self.org.rollout(self.table, self.vgoal, self.free[sl:sl + 1], 1, path0=pid, seed=MC_SEED)
                lw.append(a[0]); me.append(b[0]); owner.append(k)
        lw, me, owner = np.array(lw, dtype=np.float32), np.array(me, dtype=np.uint32), np.array(owner)
        # reduce_paths buckets by (path0 + q) % nstarts: feed it one slot's records at a time
        mx = np.full((len(slots), 4), -np.inf); sm = np.zeros((len(slots), 4)); cn = np.zeros((len(slots), 4), dtype=np.uint64)
        for k in range(len(slots)):
            sel = owner == k
            if sel.any():
                m1, s1, c1, _tr = self.oracle.reduce_paths(lw[sel], me[sel], 1)
                mx[k], sm[k], cn[k] = m1[0], s1[0], c1[0]
        lsa, _ls = self.oracle.logz_from_stats(mx, sm, cn)
        return cn, lsa, len(lw)


def parity_section(args, eng, osd, last_path0, last_count, step_id0, step_id1, world, sample):
    """Driver-run parity of the production launch: the first `sample` per-path records of the LAST TIMED rollout launch
    against the oracle, bit for bit, and the accumulators the step left on this rank (after the exchange at N > 1: the
    reduction over all ranks' shards of the step's path ids [step_id0, step_id1)) for a sample of start slots."""
    from piql_jl_b200 import distributed as pd
    import torch
    n = int(min(sample, last_count))
    logw, meta = eng.mc_records(n)
    lw, me = osd.records(last_path0, n)
    rec_ok = bool(np.array_equal(meta, me) and np.array_equal(logw.view(np.uint32), lw.view(np.uint32)))
    nfree = len(osd.free)
    slots = np.unique(np.random.default_rng(7).integers(0, nfree, size=args.parity_slots))
    m, sm, c = pd.stats_tensors(eng)
    idx = torch.as_tensor((slots[:, None] * 4 + np.arange(4)[None, :]).ravel(), device=m.device)
    gm = m[idx].cpu().numpy().astype(np.float64).reshape(-1, 4)
    gs = sm[idx].cpu().numpy().astype(np.float64).reshape(-1, 4)
    gc = c[idx].cpu().numpy().astype(np.int64).reshape(-1, 4)
    cn, lsa, npaths = osd.slot_stats(slots, step_id0, step_id1)
    cnt_ok = bool(np.array_equal(gc, cn.astype(np.int64)))
    ok = cn > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        glsa = gm + np.log(gs) - np.log(gc.astype(np.float64))
    err = float(np.max(np.abs(glsa[ok] - lsa[ok]))) if ok.any() else 0.0
    return {"records_bit_exact": rec_ok, "paths": n, "policy_record_bytes": osd.record_bytes,
            "records_of": "first %d records of the last timed launch of the production rollout_kernel<0,false,false,STAGED=false> (path ids from %d) "
                          "against oracle.rollout on the GPU's table: log-weight bits, first action, length" % (n, last_path0),
            "accumulator_counts_exact": cnt_ok, "accumulator_logz_max_abs_err": err, "accumulator_logz_ok": bool(err < 1e-5),
            "accumulator_slots": int(len(slots)), "accumulator_paths": int(npaths),
            "accumulators_of": "the step's (max, sum, count) accumulators on this rank%s for %d random start slots against "
                               "oracle.reduce_paths over every path id of the step that starts there"
                               % (" after the exchange (all ranks' shards)" if world > 1 else "", len(slots))}


def cpu_section(args, osd, vi):
    """The oracle on the box's host cores beside the GPU numbers (rank 0, N = 1): a bounded sample of the same step."""
    oracle, org, threads = osd.oracle, osd.org, osd.threads
    free = osd.free
    sample = args.ref_paths
    starts = sample_starts(free, 0, sample)
    org.rollout(osd.table, osd.vgoal, free, 1 << 14, seed=MC_SEED, nthreads=threads)
    t0 = time.perf_counter()
    lw, meta = org.rollout(osd.table, osd.vgoal, starts, sample, path0=0, seed=MC_SEED, nthreads=threads)
    mxo, smo, cno, tr = oracle.reduce_paths(lw, meta, len(starts), path0=0)
    dt = time.perf_counter() - t0
    cpu = {"value": tr / dt, "unit": "transitions/s", "cores": threads, "kind": "port",
           "sample": "first %d paths of the step (%.2e transitions), oracle/piql_oracle.c with OpenMP over paths"
                     % (sample, tr)}
    if vi is not None:
        t0 = time.perf_counter()
        org.nu_jacobi_dense(3, nthreads=threads)
        vi["cpu_baseline"] = {"sweeps_per_s": 3 / (time.perf_counter() - t0), "cores": threads, "kind": "port",
                              "sample": "3 dense OpenMP Jacobi sweeps (f64)"}
    return cpu


def grid4096_section(args, eng, stream):
    """The config-5 workload on ONE GPU (N = 1 runs only): the same-config anchor of the 1 -> 8 scaling curve, whose N > 1
    lines run 4096^2 while the N = 1 headline is config 4 (1024^2)."""
    import torch
    name = "grid4096"
    walls, goal, params = make_world(name)
    eng.load_grid(walls, goal, **params)
    V, _ = eng.solve_nu(reward_scale=VLO_SCALE, tol=0.0, max_sweeps=args.setup_sweeps)
    eng.policy_table(np.nan_to_num(V), beta=BETA, download=False)
    P = WORKLOADS[name]["paths"]
    steps, k = max(2, min(args.steps, 4)), [0]

    def step():
        eng.mc_reset()
        eng.mc_rollout(P, path0=k[0] * P, seed=MC_SEED)
        k[0] += 1
    with torch.cuda.stream(stream):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        eng.mc_timing()
        tr0 = eng.mc_transitions_total()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            step()
        b.record(stream)
        torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    tr = eng.mc_transitions_total() - tr0
    roll_s, red_s, nl = eng.mc_timing()
    peak, _src = measured_peak()
    rec_bytes = 16 if eng.policy_table16(download=False)[1] else 32
    alg = 32.0 * tr + 8.0 * P * steps
    return {"value": tr / (ms * 1e-3), "unit": "transitions/s", "steps": steps, "warmup": 3, "ms_per_step": ms / steps,
            "config": workload_config(name, args, 1),
            "roofline": rollout_roofline(name, alg / roll_s / 1e9 if roll_s > 0 else 0.0, peak, nl, roll_s, red_s, tr, rec_bytes=rec_bytes),
            "note": "same workload as the N > 1 lines (BASELINE config 5) on one GPU, no exchange: divide an N-GPU `value` by "
                    "N x this for the same-config weak-scaling efficiency"}


ROLL_LIMITER = {
    "grid1024": ("issue/ALU", "the 34 MB policy table stays in L2/L1 (DRAM traffic is ~0.3 % of the algorithmic bytes): the kernel is bound "
                              "by instruction issue and the integer pipe (Philox4x32-10; ncu: ALU pipe 60 %, issue slots 62 % busy), not by HBM; "
                              "`achieved` is the algorithmic byte rate, comparable across grids, not DRAM traffic"),
    "grid4096": ("issue/ALU + load latency", "the kernel reads the 16-byte form of the record (269 MB table): DRAM traffic per launch fell from "
                                             "237 GB (32-byte records) to `traffic` while the time barely moved -- every thread has one dependent load "
                                             "in flight (long_scoreboard 6.0 cycles per issue, was 11.2) and the ALU pipe (56 %) / issue slots (58 %) "
                                             "are as busy as on the L2-resident grid; `achieved` is the algorithmic byte rate (32 B per transition, "
                                             "SURVEY 8d), `dram_gbs` the measured DRAM rate"),
}


def rollout_roofline(name, achieved, peak, nl, roll_s, red_s, tr_local, peak_src=None, paths_per_launch=None, rec_bytes=32):
    traffic = prof = None
    tpath = os.path.join(ROOT, "profiles", "rollout_traffic.json")
    # the ncu capture is one launch of the workload's own size: it says nothing about other launch sizes
    if os.path.exists(tpath) and paths_per_launch in (None, WORKLOADS.get(name, {}).get("paths")):
        try:
            t = json.load(open(tpath))
            traffic, prof = t.get(name), t.get("_pipes", {}).get(name)
        except Exception:
            traffic = None
    bound, why = ROLL_LIMITER.get(name, ("hbm", ""))
    launch_ms = 1e3 * roll_s / max(nl, 1)
    out = {"bound": bound, "kernel": "rollout_kernel<0,false,false,STAGED=false,REC16=%s>" % ("true" if rec_bytes == 16 else "false"), "achieved": achieved, "peak": peak, "unit": "GB/s",
           "frac": achieved / peak, "traffic": traffic, "bytes_per_transition": 32, "bytes_per_path": 8, "launches": nl,
           "record_bytes_read_by_kernel": rec_bytes,
           "avg_launch_ms": launch_ms, "reduce_share": red_s / max(roll_s + red_s, 1e-12),
           "transitions_per_s_kernel": tr_local / roll_s if roll_s > 0 else 0.0, "limiter": why}
    if traffic and launch_ms > 0:
        # ncu DRAM bytes of one full-size launch over the live launch time of this run
        out["dram_gbs"] = traffic / (launch_ms * 1e-3) / 1e9
        out["dram_frac_of_peak"] = out["dram_gbs"] / peak
    if prof:
        out["ncu_pipes"] = prof
    if peak_src:
        out["peak_source"] = peak_src
    return out


def workload_config(name, args, world):
    w = WORKLOADS[name]
    return {"workload": "synthetic %dx%d gridworld (BASELINE config %s), density %.1f, randomness %.2f, gamma %.2f, beta %.0f, "
                        "V = %d Jacobi sweeps at reward x %.1f, estimator A, starts cycle over free cells"
                        % (w["size"], w["size"], "4" if name == "grid1024" else "5" if name == "grid4096" else "-", DENSITY,
                           RANDOMNESS, GAMMA, BETA, VLO_SWEEPS, VLO_SCALE),
            "paths_per_step": int(args.paths or w["paths"]) * world,
            "sharding": "path ids split in %d contiguous blocks of %d (weak scaling)" % (world, int(args.paths or w["paths"])),
            "l2": "each step writes 8 B x paths of records (> L2) between reads of the policy table; the table is "
                  "%.0f MB (%d-byte records)" % ((16 if (w["size"] + 2) ** 2 >= 1 << 22 else 32) * (w["size"] + 2) ** 2 / 1e6,
                                              16 if (w["size"] + 2) ** 2 >= 1 << 22 else 32)}


def run_ours(args, rank, world, local_rank):
    import torch
    import piql_jl_b200 as piql
    from piql_jl_b200 import distributed as pd
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    name = args.workload
    w = WORKLOADS[name]
    P = int(args.paths or w["paths"]) * world     # weak scaling: every rank rolls out the same number of paths per step
    walls, goal, params = make_world(name)
    eng = piql.Engine(local_rank)
    eng.load_grid(walls, goal, **params)
    H, W = eng.H, eng.W
    n = H * W
    stream = torch.cuda.ExternalStream(eng.stream, device="cuda:%d" % local_rank)

    # V_lo on the GPU: exactly VLO_SWEEPS Jacobi sweeps in float64 (setup, untimed)
    V, vst = eng.solve_nu(reward_scale=VLO_SCALE, tol=0.0, max_sweeps=args.setup_sweeps)
    V = np.nan_to_num(V)
    eng.policy_table(V, beta=BETA, download=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_no = [0]
    xev = []   # CUDA events around the exchange step (N > 1)

    def step():
        p0, cnt = pd.shard_range(P, rank, world)
        eng.mc_reset()
        eng.mc_rollout(cnt, path0=step_no[0] * P + p0, seed=MC_SEED)
        if world > 1:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            pd.allreduce_engine_stats(eng, mode=args.exchange, rank=rank, world=world)
            b.record(stream)
            xev.append((a, b))
        step_no[0] += 1

    with torch.cuda.stream(stream):
        sampler = ClockSampler(local_rank)
        if rank == 0 and os.environ.get("PIQL_BENCH_NO_SMI") is None:
            sampler.start()        # nvidia-smi needs ~0.1 s to come up: it also sees the warm-up steps
        for _ in range(args.warmup):
            step()
        barrier()
        eng.mc_timing()            # drop warm-up launches from the kernel timers
        del xev[:]
        tr0 = eng.mc_transitions_total()
        l0 = eng.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = e0.elapsed_time(e1)
        launches = eng.launches - l0
        roll_s, red_s, nl = eng.mc_timing()
        xch_ms = sum(a.elapsed_time(b) for a, b in xev) / max(len(xev), 1)
        tr_local = eng.mc_transitions_total() - tr0
        stat = torch.tensor([ms, float(tr_local)], dtype=torch.float64, device="cuda:%d" % local_rank)
        if world > 1:
            mx = stat.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stat.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, transitions = float(mx[0]), float(sm[1])
        else:
            transitions = float(tr_local)
    value = transitions / (ms * 1e-3)

    # roofline of the dominant kernel (rollout) on this rank: 32 B policy record per sampled transition
    # + 8 B record per finished path, over the CUDA-event time of the rollout launches
    peak, peak_src = measured_peak()
    p0, cnt = pd.shard_range(P, rank, world)
    # algorithmic bytes: SURVEY 8(d)'s per-unit figure, 32 B per sampled transition (one policy record {cdf[3], mask, beta fit[4]})
    # + 8 B per finished path -- whatever the kernel actually moves (on grids beyond L2 it reads a 16-byte form of the record)
    rec_bytes = 16 if eng.policy_table16(download=False)[1] else 32
    alg_bytes = 32.0 * tr_local + 8.0 * cnt * args.steps
    achieved = alg_bytes / roll_s / 1e9 if roll_s > 0 else 0.0
    roofline = rollout_roofline(name, achieved, peak, nl, roll_s, red_s, tr_local, peak_src, paths_per_launch=cnt, rec_bytes=rec_bytes)

    # ---- parity of the launch just timed, against the oracle (every rank checks its own shard) -------------------
    parity = None
    osd = None
    if not args.no_parity:
        import oracle
        def _parity():
            nonlocal osd
            osd = OracleSide(eng, walls, goal, params, V, threads=max(1, oracle.num_threads() // world))
            last = step_no[0] - 1
            chunk = 1 << 28                                  # the library's records per launch (MAX_RECORDS_PER_LAUNCH)
            off = ((cnt - 1) // chunk) * chunk if cnt > 0 else 0
            return parity_section(args, eng, osd, last * P + p0 + off, cnt - off, last * P, (last + 1) * P, world,
                                  args.parity_paths if world == 1 else max(args.parity_paths // world, 1 << 16))
        parity = run_extra("parity", _parity)
        if world > 1:
            flags = [parity.get("records_bit_exact", False), parity.get("accumulator_counts_exact", False), parity.get("accumulator_logz_ok", False)]
            t = torch.tensor([1.0 if f else 0.0 for f in flags], dtype=torch.float64, device="cuda:%d" % local_rank)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            parity["all_ranks"] = {"records_bit_exact": bool(t[0] > 0.5), "accumulator_counts_exact": bool(t[1] > 0.5),
                                   "accumulator_logz_ok": bool(t[2] > 0.5), "ranks": world}
    exchange = None
    if world > 1:
        # the exchange alone: ranks aligned by a barrier first, so no rank waits for a slower rollout
        iso = []
        with torch.cuda.stream(stream):
            for _ in range(5):
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                pd.allreduce_engine_stats(eng, mode=args.exchange, rank=rank, world=world)
                b.record(stream)
                torch.cuda.synchronize()
                iso.append(a.elapsed_time(b))
        iso_t = torch.tensor([min(iso[1:])], dtype=torch.float64, device="cuda:%d" % local_rank)
        dist.all_reduce(iso_t, op=dist.ReduceOp.MAX)
        iso_ms = float(iso_t[0])
        nacc = 4 * eng.nfree
        # each rank pulls (W-1)/W of 12 B x nacc from its peers and pushes as much; one direction of its links carries the
        # read responses of its pulls AND the peers' pushes (the other direction: its pushes and the peers' read responses)
        link_bytes = 2 * 12.0 * nacc * (world - 1) / world
        exchange = {"mode": args.exchange, "ms_per_step": xch_ms, "accumulator_bytes": 12 * nacc,
                    "nvlink_bytes_per_direction": link_bytes,
                    "isolated_ms": iso_ms, "achieved_gbs_per_direction": link_bytes / (iso_ms * 1e-3) / 1e9 if iso_ms > 0 else None,
                    "note": "ms_per_step is measured inside the step and includes waiting for the slowest rank's rollout; "
                            "isolated_ms is the exchange alone after a barrier (max over ranks of the best of 4)",
                    "peer_copy_reference_gbs": 770.0}

    # ---- end to end through the public call, host buffers in pinned memory --------------------
    e2e = None
    if not args.no_e2e:
        fdt, cdt, ob = (np.float32, np.uint32, 4) if args.e2e_f32 else (np.float64, np.uint64, 8)
        Vp, _k1 = pinned((H, W), np.float64, order="F")
        Vp[...] = V
        bufs, keep = [], []
        for _ in range(2):      # two sets of pinned result buffers alternate: iteration k is read while k + 1 is produced
            sa, k2 = pinned((n, 4), fdt)
            cn, k3 = pinned((n, 4), cdt)
            ls, k4 = pinned((H, W), fdt, order="F")
            bufs.append((sa, cn, ls)); keep += [k2, k3, k4]
        est = pd.ShardedEstimator(eng, rank, world, P, mode=args.exchange, root_only=True, out_dtype=fdt, buffers=bufs)
        reps = max(2, min(args.steps, 12))     # the last iteration's download is not hidden by a next rollout: enough iterations to amortise that tail
        tot = 0.0
        est.step(Vp, BETA, seed=MC_SEED + 1)   # warm-up
        est.finish()
        barrier()
        eng.mc_timing()
        t0 = time.perf_counter()
        for i in range(reps):
            prev = est.step(Vp, BETA, seed=MC_SEED + 2 + i)       # returns iteration i - 1, downloaded under iteration i's rollout
            if prev is not None:
                tot += float(prev[1].transitions)
        last_out, last_st = est.finish()
        tot += float(last_st.transitions)
        barrier()
        dt = time.perf_counter() - t0
        e_roll_s, e_red_s, e_nl = eng.mc_timing()
        # the pipelined calls against the blocking ones: the last iteration again through piql_policy_table / piql_mc_finalize
        same = None
        if not args.no_parity:
            ref = pd.mc_logz_sharded(eng, Vp, BETA, P, rank, world, seed=MC_SEED + 2 + reps - 1, mode=args.exchange, root_only=True)
            if rank == 0:
                rsa, rcn, rls, rst = ref
                ok = rcn > 0
                same = bool(np.array_equal(last_out[1].astype(np.uint64), rcn) and
                            np.allclose(last_out[0][ok], rsa[ok], rtol=0, atol=1e-6 if args.e2e_f32 else 0) and
                            np.allclose(last_out[2][~np.isnan(rls)], rls[~np.isnan(rls)], rtol=0, atol=1e-6 if args.e2e_f32 else 0))
        stat = torch.tensor([dt, tot], dtype=torch.float64, device="cuda:%d" % local_rank)
        if world > 1:
            mx = stat.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stat.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            dt, tot = float(mx[0]), float(sm[1])
        e2e = {"value": tot / dt, "unit": "transitions/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 9 * ob * n,
               "steps": reps, "ms_per_step": 1e3 * dt / reps,
               "rollout_ms_per_launch": 1e3 * e_roll_s / max(e_nl, 1), "reduce_ms_per_launch": 1e3 * e_red_s / max(e_nl, 1),
               "matches_blocking_calls": same, "output_dtype": "f32/u32" if args.e2e_f32 else "f64/u64",
               "call": "per iteration: piql_policy_table (pinned host V in, every rank) + piql_mc_reset_dev + piql_mc_rollout_dev "
                       "(+ exchange) + piql_mc_finalize_async; logz_sa / count_sa / logz_s land in pinned host buffers on rank 0 "
                       "(results are replicated) through the library's copy stream while the next iteration rolls out; the timed "
                       "region ends when the last iteration's results are on the host (piql_mc_finalize_wait)"}

    # ---- soft-VI sweeps on the same grid (rank 0) ----------------------------------------------
    # algorithmic bytes per state per sweep: nu 2 x sizeof(T), logz 3 x, gap 4 x
    vi = None
    if rank == 0 and not args.no_vi:
        vi = run_extra("vi", lambda: vi_section(args, eng, stream, V, peak, vst, name))

    # ---- soft-VI sweeps of ONE grid on all N GPUs (column slabs, halo exchange over NVLink peer memory) -------------
    vi_sharded = None
    if world > 1 and not args.no_vi:
        vi_sharded = {}
        if not args.no_parity:
            # north star: value-iteration results are bit-stable across GPU counts -- 37 sharded sweeps from a random iterate
            # against the same 37 sweeps on rank 0's GPU alone
            def _vi_bits():
                x0 = np.asfortranarray(np.random.default_rng(5).normal(size=(H, W)))
                one = None
                if rank == 0:
                    eng.vi_prepare("nu", precision="f32fast")
                    eng.vi_set_iterate(x0)
                    eng.vi_sweeps(37)
                    one, _r = eng.vi_fetch()
                x, sweeps, _res = pd.solve_sharded(eng, "nu", rank, world, precision="f32fast", tol=0.0, max_sweeps=37, check_every=10, x0=x0)
                same = True
                if rank == 0:
                    same = bool(np.array_equal(np.nan_to_num(x).view(np.uint64), np.nan_to_num(one).view(np.uint64)))
                t = torch.tensor([1.0 if same else 0.0], dtype=torch.float64, device="cuda:%d" % local_rank)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                return {"vi_sharded_bit_identical": bool(t[0] > 0.5), "sweeps": int(sweeps), "kind": "nu", "precision": "f32fast",
                        "what": "assembled slabs of %d GPUs vs the single-GPU iterate, every bit" % world}
            vb = run_extra("vi_bits", _vi_bits)
            if parity is not None:
                parity["vi_sharded"] = vb
        for prec, nbytes in (("f32fast", 8), ("f64", 16)):
            # the ranks agree on the tile shape by timing a few candidates together (a library call cannot: every rank must
            # run the same shape); the winner stays pinned for this precision
            cands = None if prec == "f32fast" else [{}] + [dict(PIQL_TILE_KH=kh, PIQL_TILE_TJ=tj, PIQL_TILE_NW=16)
                                                           for kh in (4, 8) for tj in (27, 41, 54, 76)]
            pinned_env, tuning = pd.tune_sharded(eng, "nu", rank, world, precision=prec, candidates=cands)
            eng.vi_prepare("nu", precision=prec)
            blob, cols = eng.vi_shard_begin(rank, world)
            blobs = [None] * world
            dist.all_gather_object(blobs, blob)
            eng.vi_shard_connect(blobs)
            barrier()
            # two rounds first: a halo wait that timed out on any rank (bounded at 10 s per launch in the kernel) must not
            # be repeated for the hundred launches of the timed run
            if not pd.sharded_sweeps_ok(eng, sweeps=8):
                vi_sharded[prec] = {"error": "a halo wait timed out on some rank; timed run skipped"}
                barrier()
                eng.vi_shard_end()
                pd._pin_shape(None)
                continue
            with torch.cuda.stream(stream):
                eng.vi_sweeps(40)
                barrier()
                eng.vi_shard_wait_stats()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                eng.vi_sweeps(args.vi_sweeps)
                b.record(stream)
                barrier()
            wait_us, waits = eng.vi_shard_wait_stats()
            plan = eng.vi_tile_plan("nu", prec)
            t = torch.tensor([a.elapsed_time(b), -wait_us / max(waits, 1)], dtype=torch.float64, device="cuda:%d" % local_rank)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t[0]) * 1e-3
            sps = args.vi_sweeps / sec
            gbs = nbytes * eng.nfree * sps / 1e9
            one = (vi or {}).get(prec, {}).get("us_per_sweep") if isinstance(vi, dict) else None
            vi_sharded[prec] = {"sweeps_per_s": sps, "us_per_sweep": 1e6 / sps, "state_backups_per_s": sps * eng.nfree,
                                "one_gpu_us_per_sweep": one, "efficiency_vs_one_gpu": (one / (world * 1e6 / sps)) if one else None,
                                "tile_plan": plan, "pinned": pinned_env,
                                "min_over_ranks_mean_halo_wait_us": -float(t[1]),
                                "tuning_us_per_sweep": {json.dumps(dict(k)): round(v, 2) for k, v in sorted(tuning.items(), key=lambda kv: kv[1])[:6]},
                                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak * world, "unit": "GB/s",
                                             "frac": gbs / (peak * world), "bytes_per_state": nbytes}}
            barrier()
            eng.vi_shard_end()
            pd._pin_shape(None)
        vi_sharded["note"] = ("nu sweeps of the one %s grid split into %d column slabs; 4 sweeps per round, edge columns pushed to "
                              "the neighbours over NVLink after every round; values bit-identical to the 1-GPU solve" % (name, world))

    # ---- CPU baseline beside it (rank 0, N = 1) ------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        def _cpu():
            return cpu_section(args, osd or OracleSide(eng, walls, goal, params, V), vi)
        cpu = run_extra("cpu_baseline", _cpu)

    # ---- the reference's own experiment (BASELINE configs 1-3: golden 20x20 grid) through the batched calls ------
    experiment = None
    if rank == 0 and world == 1 and not args.no_exp:
        experiment = run_extra("experiment", lambda: run_golden_experiment(local_rank, with_cpu=not args.no_cpu))

    # ---- the sparse-table path (any ControlProblem through its closures): the reference's binary bandit ----
    table = None
    if rank == 0 and world == 1 and not args.no_exp:
        table = run_extra("table", lambda: table_section(local_rank, with_cpu=not args.no_cpu))

    # ---- the config-5 workload on this one GPU: same-config anchor of the scaling curve (last: it reloads the grid) ----
    g4096 = None
    if rank == 0 and world == 1 and name == "grid1024" and not args.no_4096:
        g4096 = run_extra("rollout_grid4096", lambda: grid4096_section(args, eng, stream))

    if rank == 0:
        line = {"metric": "sampled transitions/sec", "value": value, "unit": "transitions/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(name, args, world), "clocks": clocks, "exchange": exchange, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "parity": parity, "cpu_baseline": cpu, "vi": vi, "vi_sharded": vi_sharded, "experiment": experiment,
                "rollout_grid4096": g4096, "table": table,
                "transitions_per_step": transitions / args.steps}
        print(json.dumps(line), flush=True)
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + list(WORKLOADS))
    ap.add_argument("--paths", type=int, default=0, help="paths per step over all ranks (default: workload's)")
    ap.add_argument("--ref-paths", type=int, default=1 << 20, help="CPU sample: paths per reference step")
    ap.add_argument("--vi-sweeps", type=int, default=400)
    ap.add_argument("--setup-sweeps", type=int, default=VLO_SWEEPS, help="Jacobi sweeps that build V_lo (profiling runs use fewer)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: fused log-sum-exp all-reduce kernel over NVLink peer memory, or three NCCL all-reduces")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-f32", action="store_true", help="end-to-end leg downloads float32 / uint32 results (36 B per cell instead of 72)")
    ap.add_argument("--no-vi", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed launch's records and accumulators")
    ap.add_argument("--no-4096", action="store_true", help="N = 1: skip the config-5 (4096^2) single-GPU rate")
    ap.add_argument("--parity-paths", type=int, default=1 << 20, help="records of the last timed launch compared with the oracle")
    ap.add_argument("--parity-slots", type=int, default=64, help="start slots whose accumulators are compared with the oracle")
    ap.add_argument("--no-exp", action="store_true", help="skip the golden-grid experiment (2 nu + 40 logz_updateV + 48 gap solves)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.workload == "auto":
        args.workload = "grid1024" if args.gpus <= 1 else "grid4096"
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
