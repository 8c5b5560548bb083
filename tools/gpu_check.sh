#!/bin/bash
# Short GPU-box pass for a tight budget: gpu tests with durations, then the default bench line, then the 4096^2 sweep rates.
# usage (under gpurun): bash tools/gpu_check.sh <tag>
TAG=${1:-chk}
O=gpurun_out
mkdir -p $O
(timeout 420 python -m pytest tests -m gpu -q -x --durations=25 2>&1 | tail -60) > $O/tests_$TAG.log 2>&1
tail -4 $O/tests_$TAG.log
timeout 200 python bench.py > $O/bench1024_$TAG.json 2> $O/bench1024_$TAG.err
timeout 200 python bench.py --workload grid4096 --steps 4 --no-cpu --no-exp --no-e2e > $O/bench4096_$TAG.json 2> $O/bench4096_$TAG.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench*_$TAG.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %s frac %s clocks %s" % (d["value"], (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("clocks")))
        print("  vi", {k:(round(v.get("roofline",{}).get("frac"),3), round(v.get("us_per_sweep"),2)) for k,v in (d.get("vi") or {}).items() if isinstance(v,dict) and "roofline" in v}, (d.get("vi") or {}).get("tile_plan"))
    except Exception as e:
        print(f, "unreadable", e)
PY
