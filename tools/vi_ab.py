#!/usr/bin/env python3
"""A/B timing of compile-time variants of the soft-VI kernels.
    build (CPU box):  python tools/vi_ab.py build NAME -DPIQL_TILE_UNROLL=3 ...   -> tools/probe/variants/libpiql_NAME.so
    time  (GPU box):  python tools/vi_ab.py time [grid1024|grid4096 ...]          -> us/sweep of every variant + the product build
Each timing runs in its own process with PIQL_B200_LIB pointing at the variant.
Switches the sources honour: PIQL_TILE_UNROLL (columns a warp interleaves, default 1), PIQL_TILE_PACKED (FADD2/FFMA2 nu cell, 1),
PIQL_LOGZ_ZDOM (linear-domain fast log Z, 1), PIQL_TILE_PLANES (constant planes in shared memory, 1).  The depth of the temporal
blocking (4 or 8 sweeps per round) and planes-in-shared-memory are plan variants the library times itself; PIQL_TILE_KH and
PIQL_TILE_PLANES_MODE pin them."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "tools", "probe", "variants")

CHILD = r"""
import sys, time, json
import numpy as np
sys.path.insert(0, %r)
import bench, piql_jl_b200 as piql
name = sys.argv[1]
walls, goal, params = bench.make_world(name)
eng = piql.Engine(0); eng.load_grid(walls, goal, **params)
V = np.nan_to_num(eng.solve_nu(reward_scale=0.1, tol=0.0, max_sweeps=64, precision="f32fast")[0])
out = {}
for kind, prec, kw in (("nu", "f32fast", {}), ("nu", "f32", {}), ("nu", "f64", {}), ("logz", "f32fast", dict(V=V, beta=1.0)), ("gap", "f32fast", dict(V=V, nu=V))):
    eng.vi_prepare(kind, precision=prec, **kw)
    eng.vi_sweeps(40); eng.vi_fetch(want_x=False)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); eng.vi_sweeps(400); eng.vi_fetch(want_x=False)
        best = min(best, (time.perf_counter() - t0) / 400 * 1e6)
    p = eng.vi_tile_plan(kind, prec)
    out[kind + "_" + prec] = (round(best, 2), p["TJ"], p["warps"], p["ctas_per_sm"], p["persistent"], p["sweeps_per_round"], p["planes_in_smem"])
print(json.dumps(out))
""" % ROOT


def main():
    if sys.argv[1] == "build":
        sys.path.insert(0, os.path.join(ROOT, "piql.jl_b200"))
        import build as B
        print(B.build(out=os.path.join(VDIR, "libpiql_%s.so" % sys.argv[2]), defines=sys.argv[3:]))
        return
    grids = sys.argv[2:] or ["grid1024", "grid4096"]
    libs = [("product", None)] + [(os.path.basename(p)[8:-3], p) for p in sorted(glob.glob(os.path.join(VDIR, "libpiql_*.so")))]
    for g in grids:
        for name, path in libs:
            env = dict(os.environ)
            if path:
                env["PIQL_B200_LIB"] = path
            r = subprocess.run([sys.executable, "-c", CHILD, g], env=env, capture_output=True, text=True, timeout=300)
            print(g, name, r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else "FAILED " + r.stderr[-300:], flush=True)


if __name__ == "__main__":
    main()
