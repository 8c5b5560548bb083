#!/bin/bash
# quick soft-VI timing of both synthetic grids with the default tile plan (+ optional env overrides)
for W in grid1024 grid4096; do
timeout 120 python bench.py --workload $W --steps 1 --warmup 3 --paths 1048576 --no-e2e --no-cpu --setup-sweeps 8 --vi-sweeps 400 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    v=d['vi']
    print('$W', ' '.join('%s %.2fus frac %.3f' % (k, v[k]['us_per_sweep'], v[k]['roofline']['frac']) for k in ('f32fast','f32','f64')))
"
done
