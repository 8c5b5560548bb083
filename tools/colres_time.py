#!/usr/bin/env python3
"""Per-sweep cost and per-call overhead of a soft-VI configuration: times vi_sweeps(N) for several N (CUDA events on the
engine's stream) and fits time = a + b N.   usage: colres_time.py grid1024 nu f32fast [N ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import piql_jl_b200 as piql  # noqa: E402


def main():
    name, kind, prec = sys.argv[1], sys.argv[2], sys.argv[3]
    ns = [int(v) for v in sys.argv[4:]] or [25, 50, 100, 200, 400, 800]
    walls, goal, params = bench.make_world(name)
    eng = piql.Engine(0)
    eng.load_grid(walls, goal, **params)
    kw = {}
    if kind != "nu":
        V = np.nan_to_num(eng.solve_nu(reward_scale=0.1, tol=0.0, max_sweeps=64, precision="f32fast")[0])
        kw = dict(V=V, beta=1.0) if kind == "logz" else dict(V=V, nu=V)
    stream = torch.cuda.ExternalStream(eng.stream, device="cuda:0")
    eng.vi_prepare(kind, precision=prec, **kw)
    eng.vi_sweeps(64)
    torch.cuda.synchronize()
    plan = eng.vi_tile_plan(kind, prec)
    ts = []
    for n in ns:
        best = 1e9
        for _ in range(3):
            with torch.cuda.stream(stream):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                eng.vi_sweeps(n)
                b.record(stream)
                torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) * 1e3)
        ts.append(best)
    A = np.vstack([np.ones(len(ns)), np.array(ns, dtype=float)]).T
    (a0, b0), *_ = np.linalg.lstsq(A, np.array(ts), rcond=None)
    print(name, kind, prec, "cr%d nw%d" % (plan.get("column_resident", 0), plan["warps"]),
          " ".join("%d:%.0fus" % (n, t) for n, t in zip(ns, ts)), "| per call %.1f us + %.3f us/sweep" % (a0, b0))


if __name__ == "__main__":
    main()
