#!/usr/bin/env python3
"""Why are rollouts a few percent faster inside the end-to-end pipeline than in the device-timed loop?  Times the rollout
kernel (library events) under variations: table rebuilt before every launch or not, path0 fixed or advancing, seed fixed or varying."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench, piql_jl_b200 as piql
walls, goal, params = bench.make_world("grid1024")
eng = piql.Engine(0); eng.load_grid(walls, goal, **params)
V = np.nan_to_num(eng.solve_nu(reward_scale=0.1, tol=0.0, max_sweeps=1000)[0])
eng.policy_table(V, beta=1.0, download=False)
P = 1 << 27
def run(label, rebuild, adv_path, vary_seed, n=6):
    eng.mc_reset(); eng.mc_rollout(P, seed=1); eng.mc_finalize(want=()); eng.mc_timing()
    for k in range(n):
        if rebuild: eng.policy_table(V, beta=1.0, download=False)
        eng.mc_reset()
        eng.mc_rollout(P, path0=(k * P if adv_path else 0), seed=(bench.MC_SEED + k if vary_seed else bench.MC_SEED))
    _sa, _c, _s, st = eng.mc_finalize(want=())
    ro, re, nl = eng.mc_timing()
    print("%-44s rollout %.2f ms/launch, reduce %.2f ms" % (label, 1e3 * ro / nl, 1e3 * re / nl), flush=True)
run("fixed table, advancing path0, fixed seed", False, True, False)
run("fixed table, path0 = 0, fixed seed", False, False, False)
run("fixed table, path0 = 0, varying seed", False, False, True)
run("table rebuilt each step, advancing path0", True, True, False)
run("table rebuilt each step, path0 = 0, vary seed", True, False, True)
run("fixed table, advancing path0, fixed seed", False, True, False)

def run2(label, table_async, fin_async, n=6):
    eng.mc_reset(); eng.mc_rollout(P, seed=1); eng.mc_finalize(want=()); eng.mc_timing()
    nn = eng.H * eng.W
    bufs = (np.empty((nn, 4)), np.empty((nn, 4), dtype=np.uint64), np.empty((eng.H, eng.W), order="F"))
    pend = False
    for k in range(n):
        if table_async: eng.policy_table_async(V, beta=1.0)
        eng.mc_reset()
        eng.mc_rollout(P, path0=k * P, seed=bench.MC_SEED)
        if fin_async:
            if pend: eng.mc_finalize_wait()
            eng.mc_finalize_async(bufs); pend = True
    if pend: eng.mc_finalize_wait()
    eng.mc_finalize(want=())
    ro, re, nl = eng.mc_timing()
    print("%-44s rollout %.2f ms/launch, reduce %.2f ms" % (label, 1e3 * ro / nl, 1e3 * re / nl), flush=True)
run2("async table only", True, False)
run2("async finalize only", False, True)
run2("both", True, True)
run("fixed table again", False, True, False)

# exactly bench.py's end-to-end leg
import torch
from piql_jl_b200 import distributed as pd
H, W = eng.H, eng.W
nn = H * W
Vp, _k = bench.pinned((H, W), np.float64, order="F"); Vp[...] = V
bufs, keep = [], []
for _ in range(2):
    sa, k2 = bench.pinned((nn, 4), np.float64); cn, k3 = bench.pinned((nn, 4), np.uint64); ls, k4 = bench.pinned((H, W), np.float64, order="F")
    bufs.append((sa, cn, ls)); keep += [k2, k3, k4]
for variant in ("pinned V + pinned out", "pageable"):
    est = pd.ShardedEstimator(eng, 0, 1, P, root_only=True, buffers=bufs if variant.startswith("pinned") else None)
    est.step(Vp if variant.startswith("pinned") else V, 1.0, seed=1); est.finish(); eng.mc_timing()
    t0 = time.perf_counter()
    for i in range(6):
        est.step(Vp if variant.startswith("pinned") else V, 1.0, seed=bench.MC_SEED + 2 + i)
    est.finish(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ro, re, nl = eng.mc_timing()
    print("estimator (%s): rollout %.2f ms/launch, step %.2f ms" % (variant, 1e3 * ro / nl, 1e3 * dt / 6), flush=True)
run("fixed table afterwards", False, True, False)
