#!/usr/bin/env python3
"""Accuracy of the fp32 soft-VI modes against the Float64 fixed point on a bench workload (north star: 1e-5 relative).
usage: vi_acc.py [grid1024]   -> max relative error of nu and log Z for f32 and f32fast after enough sweeps to converge"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import piql_jl_b200 as piql  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "grid1024"
    walls, goal, params = bench.make_world(name)
    eng = piql.Engine(0)
    eng.load_grid(walls, goal, **params)
    free = walls == 0
    nsw = 3000
    ref, _ = eng.solve_nu(tol=0.0, max_sweeps=nsw, precision="f64")
    for prec in ("f32", "f32fast"):
        x, st = eng.solve_nu(tol=0.0, max_sweeps=nsw, precision=prec)
        rel = np.abs(x[free] - ref[free]) / np.maximum(1.0, np.abs(ref[free]))
        print("%s nu %s: max rel err %.3e  mean %.3e  (%.1f us/sweep incl. tuning)" % (name, prec, rel.max(), rel.mean(), 1e6 * st.device_seconds / nsw))
    V = np.nan_to_num(eng.solve_nu(reward_scale=0.1, tol=0.0, max_sweeps=1000, precision="f64")[0])
    for beta in (1.0, 0.2):
        lref, _ = eng.solve_logz(V, beta=beta, tol=0.0, max_sweeps=nsw, precision="f64")
        for prec in ("f32", "f32fast"):
            x, st = eng.solve_logz(V, beta=beta, tol=0.0, max_sweeps=nsw, precision=prec)
            rel = np.abs(x[free] - lref[free]) / np.maximum(1.0, np.abs(lref[free]))
            print("%s logz beta=%g %s: max rel err %.3e  mean %.3e" % (name, beta, prec, rel.max(), rel.mean()))


if __name__ == "__main__":
    main()
