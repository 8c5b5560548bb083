#!/usr/bin/env python3
"""Does a device-to-device copy / memset on the stream before the rollout change the rollout kernel's time?"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench, piql_jl_b200 as piql
walls, goal, params = bench.make_world("grid1024")
eng = piql.Engine(0); eng.load_grid(walls, goal, **params)
V = np.nan_to_num(eng.solve_nu(reward_scale=0.1, tol=0.0, max_sweeps=1000)[0])
eng.policy_table(V, beta=1.0, download=False)
P = 1 << 27
stream = torch.cuda.ExternalStream(eng.stream, device="cuda:0")
a = torch.empty(40 << 20, dtype=torch.uint8, device="cuda:0"); b = torch.empty_like(a)
other = torch.cuda.Stream()
def run(label, pre, n=6):
    eng.mc_reset(); eng.mc_rollout(P, seed=1); eng.mc_finalize(want=()); eng.mc_timing()
    for k in range(n):
        if pre == "d2d":
            with torch.cuda.stream(stream): b.copy_(a)
        elif pre == "d2d_small":
            with torch.cuda.stream(stream): b[:8].copy_(a[:8])
        elif pre == "memset":
            with torch.cuda.stream(stream): a.zero_()
        elif pre == "other_stream_kernel":
            with torch.cuda.stream(other): b.add_(1)
        elif pre == "other_stream_d2h":
            with torch.cuda.stream(other): hp.copy_(b, non_blocking=True)
        eng.mc_reset()
        eng.mc_rollout(P, path0=0, seed=bench.MC_SEED + k)
    torch.cuda.synchronize()
    ro, re, nl = eng.mc_timing()
    print("%-40s rollout %.2f ms/launch" % (label, 1e3 * ro / nl), flush=True)
hp = torch.empty(40 << 20, dtype=torch.uint8, pin_memory=True)
run("nothing", None)
run("40 MB D2D copy on the stream before", "d2d")
run("8 B D2D copy on the stream before", "d2d_small")
run("40 MB memset on the stream before", "memset")
run("kernel on another stream (enqueued before)", "other_stream_kernel")
run("D2H 40 MB on another stream", "other_stream_d2h")
run("nothing", None)
