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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
def run_reference(args, rank, world):
    """The reference's algorithm on the host cores: the C restatement (oracle/), all threads."""
    if rank != 0:
        return
    import oracle
    name = args.workload
    walls, goal, params = make_world(name)
    org = oracle.Grid(walls, goal, **params)
    threads = oracle.num_threads()
    # V_lo as in the GPU arm, but the CPU cannot afford 1000 dense sweeps of the 4096^2 grid inside a bounded run: 64 there
    # (the rollout cost per transition does not depend on how converged V is; path lengths are set by gamma)
    ref_sweeps = min(args.setup_sweeps, 64) if name == "grid4096" else args.setup_sweeps
    V, _ = org.nu_jacobi_dense(ref_sweeps, reward_scale=VLO_SCALE, nthreads=threads)
    V = np.nan_to_num(V)
    table = org.policy_table(V, beta=BETA)
    free = org.free_cells()
    sample = args.ref_paths
    vgoal = V[goal]
    done = 0

    def step():
        nonlocal done
        lw, meta = org.rollout(table, vgoal, free, sample, path0=done, seed=MC_SEED, nthreads=threads)
        mx, sm, cnt, tr = oracle.reduce_paths(lw, meta, len(free), path0=done)
        done += sample
        return tr
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    transitions = sum(step() for _ in range(args.steps))
    dt = time.perf_counter() - t0
    value = transitions / dt
    line = {
        "impl": "reference", "metric": "sampled transitions/sec", "value": value, "unit": "transitions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, args, world),
        "cpu_baseline": {"value": value, "unit": "transitions/s", "cores": threads, "kind": "port",
                         "sample": "%d paths per step (of the %d-path step per GPU), oracle/piql_oracle.c with OpenMP over paths, V_lo from "
                                   "%d Jacobi sweeps; Julia is not installed, so the reference itself cannot run"
                                   % (sample, WORKLOADS[name]["paths"], ref_sweeps)},
        "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
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


def vi_section(args, eng, stream, V, peak, vst):
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
        vi[key] = {"sweeps_per_s": sps, "state_backups_per_s": sps * eng.nfree, "us_per_sweep": 1e6 / sps,
                   "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                "bytes_per_state": nbytes, "traffic": None}}
    # the ground-truth solve itself (config 4: "soft VI ground truth on 1 B200"), host buffers in and out
    solves = {}
    for prec, tol in (("f64", 1e-9), ("f32fast", 0.05)):
        t0 = time.perf_counter()
        _nu, sst = eng.solve_nu(tol=tol, max_sweeps=20000, precision=prec, allow_unconverged=True)
        solves[prec] = {"tol_sum_abs_delta": tol, "sweeps": int(sst.sweeps), "converged": bool(sst.converged),
                        "seconds": time.perf_counter() - t0, "device_seconds": float(sst.device_seconds)}
    vi["solve_nu"] = solves
    vi["kernel"] = ("tile_sweep_kernel<T,KIND,FAST>: 4 Jacobi sweeps per launch on 128 x (TJ+8) shared-memory tiles "
                    "(includes the residual reduction on the last launch); f32fast = MUFU ex2/lg2; keys without a kind are nu")
    vi["setup_vlo"] = {"sweeps": int(vst.sweeps), "device_seconds": float(vst.device_seconds)}
    return vi


def cpu_section(args, eng, walls, goal, params, V, vi):
    """The oracle on the box's host cores beside the GPU numbers (rank 0, N = 1): a bounded sample of the same step."""
    import oracle
    org = oracle.Grid(walls, goal, **params)
    threads = oracle.num_threads()
    table = eng.policy_table(V, beta=BETA)        # same fp32 table the GPU read
    free = org.free_cells()
    sample = args.ref_paths
    org.rollout(table, V[goal], free, 1 << 14, seed=MC_SEED, nthreads=threads)
    t0 = time.perf_counter()
    lw, meta = org.rollout(table, V[goal], free, sample, path0=0, seed=MC_SEED, nthreads=threads)
    mxo, smo, cno, tr = oracle.reduce_paths(lw, meta, len(free), path0=0)
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


def workload_config(name, args, world):
    w = WORKLOADS[name]
    return {"workload": "synthetic %dx%d gridworld (BASELINE config %s), density %.1f, randomness %.2f, gamma %.2f, beta %.0f, "
                        "V = %d Jacobi sweeps at reward x %.1f, estimator A, starts cycle over free cells"
                        % (w["size"], w["size"], "4" if name == "grid1024" else "5" if name == "grid4096" else "-", DENSITY,
                           RANDOMNESS, GAMMA, BETA, VLO_SWEEPS, VLO_SCALE),
            "paths_per_step": int(args.paths or w["paths"]) * world,
            "sharding": "path ids split in %d contiguous blocks of %d (weak scaling)" % (world, int(args.paths or w["paths"])),
            "l2": "each step writes 8 B x paths of records (> L2) between reads of the policy table; the table is "
                  "%.0f MB" % (32 * (w["size"] + 2) ** 2 / 1e6)}


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
        if rank == 0:
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
    alg_bytes = 32.0 * tr_local + 8.0 * cnt * args.steps
    achieved = alg_bytes / roll_s / 1e9 if roll_s > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "rollout_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(name)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "rollout_kernel<0,false,false>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_transition": 32, "bytes_per_path": 8, "launches": nl,
                "avg_launch_ms": 1e3 * roll_s / max(nl, 1), "reduce_share": red_s / max(roll_s + red_s, 1e-12),
                "transitions_per_s_kernel": tr_local / roll_s if roll_s > 0 else 0.0}
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
        Vp, _k1 = pinned((H, W), np.float64, order="F")
        Vp[...] = V
        sa, _k2 = pinned((n, 4), np.float64)
        cn, _k3 = pinned((n, 4), np.uint64)
        ls, _k4 = pinned((H, W), np.float64, order="F")
        reps = max(1, min(args.steps, 3))
        tot = 0.0
        pd.mc_logz_sharded(eng, Vp, BETA, P, rank, world, seed=MC_SEED + 1, out=(sa, cn, ls), mode=args.exchange, root_only=True)   # warm-up
        barrier()
        t0 = time.perf_counter()
        for i in range(reps):
            _sa, _cn, _ls, st = pd.mc_logz_sharded(eng, Vp, BETA, P, rank, world, seed=MC_SEED + 2 + i, out=(sa, cn, ls), mode=args.exchange, root_only=True)
            tot += float(st.transitions)
        barrier()
        dt = time.perf_counter() - t0
        stat = torch.tensor([dt, tot], dtype=torch.float64, device="cuda:%d" % local_rank)
        if world > 1:
            mx = stat.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stat.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            dt, tot = float(mx[0]), float(sm[1])
        e2e = {"value": tot / dt, "unit": "transitions/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 72 * n,
               "steps": reps, "ms_per_step": 1e3 * dt / reps,
               "call": "piql_policy_table + piql_mc_reset_dev + piql_mc_rollout_dev (+ all-reduce) + piql_mc_finalize "
                       "with pinned host V in on every rank, logz_sa / count_sa / logz_s out on rank 0 (results are replicated)"}

    # ---- soft-VI sweeps on the same grid (rank 0) ----------------------------------------------
    # algorithmic bytes per state per sweep: nu 2 x sizeof(T), logz 3 x, gap 4 x
    vi = None
    if rank == 0 and not args.no_vi:
        vi = run_extra("vi", lambda: vi_section(args, eng, stream, V, peak, vst))

    # ---- soft-VI sweeps of ONE grid on all N GPUs (column slabs, halo exchange over NVLink peer memory) -------------
    vi_sharded = None
    if world > 1 and not args.no_vi:
        vi_sharded = {}
        for prec, nbytes in (("f32fast", 8), ("f64", 16)):
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
                continue
            with torch.cuda.stream(stream):
                eng.vi_sweeps(40)
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                eng.vi_sweeps(args.vi_sweeps)
                b.record(stream)
                barrier()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda:%d" % local_rank)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t[0]) * 1e-3
            sps = args.vi_sweeps / sec
            gbs = nbytes * eng.nfree * sps / 1e9
            vi_sharded[prec] = {"sweeps_per_s": sps, "us_per_sweep": 1e6 / sps, "state_backups_per_s": sps * eng.nfree,
                                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak * world, "unit": "GB/s",
                                             "frac": gbs / (peak * world), "bytes_per_state": nbytes}}
            barrier()
            eng.vi_shard_end()
        vi_sharded["note"] = ("nu sweeps of the one %s grid split into %d column slabs; 4 sweeps per round, edge columns pushed to "
                              "the neighbours over NVLink after every round; values bit-identical to the 1-GPU solve" % (name, world))

    # ---- CPU baseline beside it (rank 0, N = 1) ------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = run_extra("cpu_baseline", lambda: cpu_section(args, eng, walls, goal, params, V, vi))

    # ---- the reference's own experiment (BASELINE configs 1-3: golden 20x20 grid) through the batched calls ------
    experiment = None
    if rank == 0 and world == 1 and not args.no_exp:
        experiment = run_extra("experiment", lambda: run_golden_experiment(local_rank, with_cpu=not args.no_cpu))

    if rank == 0:
        line = {"metric": "sampled transitions/sec", "value": value, "unit": "transitions/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(name, args, world), "clocks": clocks, "exchange": exchange, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu, "vi": vi, "vi_sharded": vi_sharded, "experiment": experiment,
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
    ap.add_argument("--no-vi", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
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
