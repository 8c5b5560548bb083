#!/usr/bin/env python3
"""PIQL_ROLL_TILED experiment (policy records in 8 x 8 blocks of cells for the production rollout kernel, see csrc/internal.h).
    build (CPU box):  python tools/roll_tiled_check.py build      -> tools/probe/variants/libpiql_rolltiled.so
    check (GPU box):  python tools/roll_tiled_check.py [grid4096] -> per library: sha1 of the per-path records of 2^22 paths
                      (must be identical: same records, other addresses) and transitions/s of 2^27-path launches.
NOT YET RUN ON A GPU (written when the round's GPU minutes were spent); the default build does not contain it."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tools", "probe", "variants", "libpiql_rolltiled.so")

CHILD = r"""
import sys, time, json, hashlib
import numpy as np
sys.path.insert(0, %r)
import bench, piql_jl_b200 as piql
name = sys.argv[1]
walls, goal, params = bench.make_world(name)
eng = piql.Engine(0); eng.load_grid(walls, goal, **params)
V = np.nan_to_num(eng.solve_nu(reward_scale=0.1, tol=0.0, max_sweeps=64, precision="f32fast")[0])
eng.policy_table(V, beta=1.0, download=False)
n = 1 << 22
eng.mc_reset(); eng.mc_rollout(n, seed=5); sa, cnt, ls, st = eng.mc_finalize(want=("count",))
logw, meta = eng.mc_records(n)
h = hashlib.sha1(logw.tobytes() + meta.tobytes()).hexdigest()
rates = []
for rep in range(3):
    eng.mc_reset(); eng.mc_timing()
    eng.mc_rollout(1 << 27, seed=7 + rep, path0=rep << 27)
    _sa, _c, _ls, st2 = eng.mc_finalize(want=())
    ro, re, nl = eng.mc_timing()
    rates.append(st2.transitions / ro)
print(json.dumps({"records_sha1": h, "transitions": int(st.transitions), "count_sum": int(cnt.sum()), "rollout_transitions_per_s": max(rates)}))
""" % ROOT


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        sys.path.insert(0, os.path.join(ROOT, "piql.jl_b200"))
        import build as B
        print(B.build(out=LIB, defines=["-DPIQL_ROLL_TILED=1"]))
        return
    grid = sys.argv[1] if len(sys.argv) > 1 else "grid4096"
    out = {}
    for name, path in (("product", None), ("rolltiled", LIB)):
        env = dict(os.environ)
        if path:
            env["PIQL_B200_LIB"] = path
        r = subprocess.run([sys.executable, "-c", CHILD, grid], env=env, capture_output=True, text=True, timeout=600)
        if r.returncode != 0 or not r.stdout.strip():
            print(name, "FAILED", r.stderr[-400:])
            continue
        out[name] = json.loads(r.stdout.strip().splitlines()[-1])
        print(grid, name, out[name], flush=True)
    if len(out) == 2:
        same = out["product"]["records_sha1"] == out["rolltiled"]["records_sha1"]
        print("records identical:", same, " speed-up: %.3f" % (out["rolltiled"]["rollout_transitions_per_s"] / out["product"]["rollout_transitions_per_s"]))


if __name__ == "__main__":
    main()
