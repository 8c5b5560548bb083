#!/bin/bash
TAG=${1:-r2h}
O=gpurun_out
mkdir -p $O
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > $O/tests_$TAG.log 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > $O/smoke_$TAG.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench1024_$TAG.json 2> $O/bench1024_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/benchref_$TAG.json 2> $O/benchref_$TAG.err
tail -4 $O/tests_$TAG.log; cat $O/smoke_$TAG.log; tail -c 400 $O/bench1024_$TAG.err
python - <<PY
import json
d=json.loads(open("$O/bench1024_$TAG.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g (%.3f) ms %.2f / %.2f" % (d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
print("parity", {k:v for k,v in d["parity"].items() if not isinstance(v,str)})
print("vi", {k:(round(v["us_per_sweep"],2), round(v["roofline"]["frac"],3), v["tile_plan"]["sweeps_per_round"], v["tile_plan"]["planes_in_smem"], v["tile_plan"]["persistent"]) for k,v in d["vi"].items() if isinstance(v,dict) and "us_per_sweep" in v})
print("g4096", d["rollout_grid4096"]["value"] if d.get("rollout_grid4096") and "value" in d["rollout_grid4096"] else d.get("rollout_grid4096"))
print("exp", d["experiment"].get("seconds"), d["experiment"].get("max_abs_err_vs_reference_data"), "cpu", d["cpu_baseline"]["value"])
PY
