#!/bin/bash
# soft-VI sweep rates at 1026^2 with the schedule the tuner picks, twice (is the choice stable?), and with the column-resident schedule off
O=gpurun_out
exec > $O/colres_times.log 2>&1
run() {
  timeout 200 python bench.py --workload grid1024 --steps 1 --warmup 3 --paths 1048576 --no-e2e --no-cpu --no-exp --no-parity --no-4096 --setup-sweeps 8 --vi-sweeps 400 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    v=d['vi']
    print('$1', ' | '.join('%s %.2fus %.3f cr%d nw%d' % (k, v[k]['us_per_sweep'], v[k]['roofline']['frac'], v[k]['tile_plan'].get('column_resident',0), v[k]['tile_plan']['warps']) for k in ('f32fast','f32','f64','logz_f32fast','gap_f32fast')))
    print('   solve', json.dumps(v.get('solve_nu')))
"
}
run tuned
run tuned
PIQL_TILE_COLRES=0 run colres_off
