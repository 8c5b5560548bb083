#!/bin/bash
# bench lines of both grids + ncu capture of the fast nu tile kernel at 1026^2 (multi-launch shape pinned, persistent kernel off)
TAG=${1:-fin}
O=gpurun_out
mkdir -p $O
timeout 200 python bench.py > $O/bench1024_$TAG.json 2> $O/bench1024_$TAG.err
timeout 200 python bench.py --workload grid4096 --steps 4 --no-cpu --no-exp > $O/bench4096_$TAG.json 2> $O/bench4096_$TAG.err
V="--steps 1 --warmup 3 --no-e2e --no-cpu --no-exp --vi-sweeps 40 --setup-sweeps 8 --paths 1048576"
PIQL_TILE_NO_PERSIST=1 PIQL_TILE_TJ=22 PIQL_TILE_NW=12 timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_sweep_kernel -s 8 -c 1 -f -o $O/prof_tile1024_$TAG python bench.py --workload grid1024 $V > $O/ncu_tile1024_$TAG.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench*_$TAG.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %s frac %s clocks %s" % (d["value"], (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("clocks")))
        print("  vi", {k:(round(v.get("roofline",{}).get("frac"),3), round(v.get("us_per_sweep"),2)) for k,v in (d.get("vi") or {}).items() if isinstance(v,dict) and "roofline" in v}, (d.get("vi") or {}).get("tile_plan"), (d.get("vi") or {}).get("solve_nu"))
    except Exception as e:
        print(f, "unreadable", e)
PY
