#!/usr/bin/env python3
"""Rollout throughput on the reference's own 20x20 gridworld (BASELINE configs 1-3): paths from s0, estimator A."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import piql_jl_b200 as piql
from conftest import Golden
for name in ("r0_01g0_99", "r0_00g0_9999"):
    g = Golden(name)
    eng = piql.Engine(0)
    eng.load_grid(g.walls, g.goal, **g.params)
    V = np.nan_to_num(g.V(7, 0))
    eng.policy_table(V, beta=1.0, download=False)
    n = 1 << 24 if "9999" not in name else 1 << 21
    for starts in ([g.s0_idx], None):
        eng.mc_reset(starts=starts); eng.mc_rollout(n, seed=1, starts=starts); eng.mc_finalize()
        eng.mc_timing()
        eng.mc_reset(starts=starts); eng.mc_rollout(n, seed=2, starts=starts)
        sa, cnt, ls, st = eng.mc_finalize()
        ro, re, nl = eng.mc_timing()
        print("%s starts=%s paths=%d transitions=%d rollout %.2f ms -> %.3e transitions/s (reduce %.2f ms)%s"
              % (name, "s0" if starts else "all free", n, st.transitions, ro * 1e3, st.transitions / ro, re * 1e3,
                 " [no staging]" if os.environ.get("PIQL_ROLL_NO_STAGE") else ""))
    eng.close()
