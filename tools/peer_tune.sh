#!/bin/bash
# exchange-kernel variants on 2 GPUs: grid size and read-only / write-only link tests
for cfg in "4 0" "8 0" "16 0" "2 0" "8 1" "8 2"; do set -- $cfg
PIQL_PEER_GRID=$1 PIQL_PEER_DBG=$2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 1 --warmup 3 --no-vi --no-e2e --paths 4194304 --setup-sweeps 8 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('grid x$1 dbg $2: isolated %.3f ms -> %.0f GB/s per direction' % (d['exchange']['isolated_ms'], d['exchange']['achieved_gbs_per_direction']))
"
done
