#!/bin/bash
TAG=${1:-r2j}
O=gpurun_out
mkdir -p $O
(timeout 1200 python -m pytest tests/test_gpu_production.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -12) > $O/tests_$TAG.log 2>&1
timeout 600 python bench.py --workload grid4096 --steps 6 --no-cpu --no-exp --no-vi > $O/bench4096_$TAG.json 2> $O/bench4096_$TAG.err
PIQL_ROLL_REC16=0 timeout 600 python bench.py --workload grid4096 --steps 6 --no-cpu --no-exp --no-vi --no-e2e > $O/bench4096_rec32_$TAG.json 2> $O/bench4096_rec32_$TAG.err
PIQL_ROLL_REC16=1 timeout 600 python bench.py --workload grid1024 --steps 6 --no-cpu --no-exp --no-vi --no-e2e --no-4096 > $O/bench1024_rec16_$TAG.json 2> $O/bench1024_rec16_$TAG.err
tail -5 $O/tests_$TAG.log
python - <<PY
import json
for f in ("bench4096_$TAG","bench4096_rec32_$TAG","bench1024_rec16_$TAG"):
    try:
        d=json.loads(open("$O/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], d["roofline"]["bytes_per_transition"], "parity", d["parity"]["records_bit_exact"], d["parity"].get("policy_record_bytes"), d["parity"]["accumulator_counts_exact"], d["parity"]["accumulator_logz_ok"])
    except Exception as e:
        print(f, "unreadable", e, open("$O/%s.err"%f).read()[-400:])
PY
