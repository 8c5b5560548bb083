#!/bin/bash
# rollout throughput on both synthetic grids (short bench runs, no VI / e2e / CPU legs)
for W in grid1024 grid4096; do
timeout 200 python bench.py --workload $W --no-vi --no-exp --no-cpu --no-e2e --steps 4 --setup-sweeps 100 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$W value %.4g  frac %.3f' % (d['value'], d['roofline']['frac']))"
done
