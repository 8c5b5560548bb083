import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, bench, piql_jl_b200 as piql
walls, goal, params = bench.make_world("grid1024")
eng = piql.Engine(0); eng.load_grid(walls, goal, **params)
for tj, nw in ((38, 16), (33, 12), (60, 16), (90,16)):
    os.environ["PIQL_TILE_TJ"], os.environ["PIQL_TILE_NW"] = str(tj), str(nw)
    eng.vi_prepare("nu", precision="f32fast"); eng.vi_sweeps(40); eng.vi_fetch(want_x=False)
    for n in (400, 4000):
        t0 = time.perf_counter(); eng.vi_sweeps(n); t1 = time.perf_counter(); eng.vi_fetch(want_x=False); t2 = time.perf_counter()
        print("TJ=%d nw=%d sweeps=%d enqueue %.2f us/launch, total %.2f us/launch (%.2f us/sweep)" % (tj, nw, n, (t1-t0)/(n/4)*1e6, (t2-t0)/(n/4)*1e6, (t2-t0)/n*1e6))
os.system("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader; nproc; cat /proc/cpuinfo | grep 'model name' | head -1")
