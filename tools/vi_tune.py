#!/usr/bin/env python3
"""Empirical sweep of the tiled soft-VI kernel's tile shape (PIQL_TILE_TJ x PIQL_TILE_NW) on a bench workload.
usage: vi_tune.py grid1024|grid4096 [precision] [kind]   -> prints us/sweep per shape, best first, and the default plan."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import piql_jl_b200 as piql  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "grid1024"
    prec = sys.argv[2] if len(sys.argv) > 2 else "f32fast"
    kind = sys.argv[3] if len(sys.argv) > 3 else "nu"
    walls, goal, params = bench.make_world(name)
    eng = piql.Engine(0)
    eng.load_grid(walls, goal, **params)
    kw = {}
    if kind != "nu":
        V = np.nan_to_num(eng.solve_nu(reward_scale=0.1, tol=0.0, max_sweeps=64, precision="f32fast")[0])
        kw = dict(V=V, beta=1.0) if kind == "logz" else dict(V=V, nu=V)
    nsw = 400

    def run():
        eng.vi_prepare(kind, precision=prec, **kw)
        eng.vi_sweeps(40)
        eng.vi_fetch(want_x=False)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            eng.vi_sweeps(nsw)
            eng.vi_fetch(want_x=False)
            best = min(best, (time.perf_counter() - t0) / nsw * 1e6)
        return best
    res = []
    tjs = list(range(16, 64, 1)) if name == "grid1024" else list(range(24, 121, 6))
    for nw in (8, 10, 12, 13, 14, 16):
        for tj in tjs:
            os.environ["PIQL_TILE_TJ"], os.environ["PIQL_TILE_NW"] = str(tj), str(nw)
            try:
                us = run()
            except Exception as e:  # shape not feasible
                continue
            p = eng.vi_tile_plan(kind, prec)
            res.append((us, nw, tj, p["tiles_i"] * p["tiles_j"], p["ctas_per_sm"], p["persistent"]))
    os.environ.pop("PIQL_TILE_TJ"); os.environ.pop("PIQL_TILE_NW")
    res.sort()
    for us, nw, tj, tiles, c, pers in res[:25]:
        print("%s %s %s nw=%d TJ=%d tiles=%d ctas/SM=%d persistent=%d  %.2f us/sweep" % (name, prec, kind, nw, tj, tiles, c, pers, us))
    print("worst %.2f" % res[-1][0])
    print("default plan", eng.vi_tile_plan(kind, prec), "%.2f us/sweep" % run())


if __name__ == "__main__":
    main()
