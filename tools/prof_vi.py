#!/usr/bin/env python3
"""One tuned soft-VI configuration under the profiler: tunes the tile plan first (untimed, unprofiled), then brackets a few
rounds of sweeps with cudaProfilerStart / Stop so that `ncu --profile-from-start off` captures only them.
usage: prof_vi.py grid1024|grid4096 nu|logz|gap f32fast|f32|f64 [rounds]"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import piql_jl_b200 as piql  # noqa: E402


def main():
    name, kind, prec = sys.argv[1], sys.argv[2], sys.argv[3]
    rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    walls, goal, params = bench.make_world(name)
    eng = piql.Engine(0)
    eng.load_grid(walls, goal, **params)
    kw = {}
    if kind != "nu":
        V = np.nan_to_num(eng.solve_nu(reward_scale=0.1, tol=0.0, max_sweeps=64, precision="f32fast")[0])
        kw = dict(V=V, beta=1.0) if kind == "logz" else dict(V=V, nu=V)
    eng.vi_prepare(kind, precision=prec, **kw)
    eng.vi_sweeps(64)                      # first call: times the plan variants
    eng.vi_fetch(want_x=False)
    plan = eng.vi_tile_plan(kind, prec)
    rt = ctypes.CDLL("libcudart.so")
    rt.cudaProfilerStart()
    eng.vi_sweeps(rounds * plan["sweeps_per_round"])
    eng.vi_fetch(want_x=False)
    rt.cudaProfilerStop()
    print(name, kind, prec, plan)


if __name__ == "__main__":
    main()
