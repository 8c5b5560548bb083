#!/bin/bash
# One GPU-box pass: gpu tests, smoke, both bench arms, ncu launch list, ncu full captures.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > $O/tests_$TAG.log 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > $O/smoke_$TAG.log 2>&1
timeout 600 python bench.py > $O/bench1024_$TAG.json 2> $O/bench1024_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/benchref_$TAG.json 2> $O/benchref_$TAG.err
timeout 600 python bench.py --workload grid4096 --steps 4 --no-cpu > $O/bench4096_$TAG.json 2> $O/bench4096_$TAG.err
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --vi-sweeps 40 --setup-sweeps 8 --paths 16777216"
timeout 300 $B > $O/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv $B > $O/ncu_launches_$TAG.log 2>&1
timeout 300 $B > $O/plain2_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o $O/prof_rollout_$TAG $B > $O/ncu_rollout_$TAG.log 2>&1
timeout 300 $B > $O/plain3_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_sweep_kernel -s 4 -c 1 -o $O/prof_tile_$TAG $B > $O/ncu_tile_$TAG.log 2>&1
tail -3 $O/tests_$TAG.log; cat $O/smoke_$TAG.log; python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench*_$TAG.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %s frac %s vi %s cpu %s" % (d["value"], (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"),
              {k:(round(v.get("roofline",{}).get("frac"),3), round(v.get("us_per_sweep"),2)) for k,v in (d.get("vi") or {}).items() if isinstance(v,dict) and "roofline" in v}, (d.get("cpu_baseline") or {}).get("value")))
    except Exception as e:
        print(f, "unreadable", e)
PY
