#!/bin/bash
# sweep the tile shape of the soft-VI kernel (PIQL_TILE_TJ x PIQL_TILE_NW) on one workload
W=${1:-grid1024}
for nw in 8 16; do for tj in 18 26 34 42 50 58 66 74 82; do
  PIQL_TILE_TJ=$tj PIQL_TILE_NW=$nw timeout 120 python bench.py --workload $W --steps 1 --warmup 3 --paths 1048576 --no-e2e --no-cpu --setup-sweeps 8 --vi-sweeps 400 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); v=d['vi']
    print('$W nw=$nw tj=$tj', ' '.join('%s %.2fus frac %.3f' % (k, v[k]['us_per_sweep'], v[k]['roofline']['frac']) for k in ('f32fast','f32','f64')))
"
done; done
echo default:
timeout 120 python bench.py --workload $W --steps 1 --warmup 3 --paths 1048576 --no-e2e --no-cpu --setup-sweeps 8 --vi-sweeps 400 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); v=d['vi']
    print('$W default', ' '.join('%s %.2fus frac %.3f' % (k, v[k]['us_per_sweep'], v[k]['roofline']['frac']) for k in ('f32fast','f32','f64')))
"
