#!/bin/bash
O=gpurun_out
exec > $O/colres_diag.log 2>&1
./tools/probe/mail_latency
run() {
  timeout 200 python bench.py --workload grid1024 --steps 1 --warmup 3 --paths 1048576 --no-e2e --no-cpu --no-exp --no-parity --no-4096 --setup-sweeps 8 --vi-sweeps 400 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    v=d['vi']
    print('$1', ' | '.join('%s %.2fus %.3f cr%d nw%d' % (k, v[k]['us_per_sweep'], v[k]['roofline']['frac'], v[k]['tile_plan'].get('column_resident',0), v[k]['tile_plan']['warps']) for k in ('f32fast','f32','f64','logz_f32fast','gap_f32fast')))
"
}
for nw in 32 28 20 16; do
  PIQL_TILE_COLRES_DIAG=1 PIQL_TILE_COLRES=1 PIQL_TILE_NW=$nw run nomail_nw$nw
done
