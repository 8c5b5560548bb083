import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import piql_jl_b200 as piql
from conftest import random_gridworld, Golden
g = Golden("r0_01g0_99")
eng = piql.Engine(0)
eng.load_grid(g.walls, g.goal, **g.params)
eng.solve_nu()
walls, goal = random_gridworld((100, 100), 0.2, seed=5)
eng.load_grid(walls, goal, 0.02, 0.02, 0.2, 0.95)
for rep in range(3):
    res = []
    for k in (32, 320, 608, 640, 960):
        out, st = eng.solve_nu(tol=0.0, max_sweeps=k, check_every=32)
        res.append((k, st.residual, float(np.nanmax(np.abs(out)))))
    print(rep, res)
    out, st = eng.solve_nu(tol=1e-11, check_every=32); print('  solve', st.sweeps, st.residual)
