#!/bin/bash
O=gpurun_out
B="python bench.py --steps 12 --warmup 5 --no-cpu --no-exp --no-vi --no-parity --no-4096"
for cfg in "100" "500" "none"; do
  if [ "$cfg" = "none" ]; then export PIQL_BENCH_NO_SMI=1; else export PIQL_BENCH_SMI_MS=$cfg; fi
  timeout 300 $B > $O/smi_$cfg.json 2> $O/smi_$cfg.err
  python -c "
import json
d=json.loads(open('$O/smi_$cfg.json').read().strip().splitlines()[-1])
print('$cfg', 'ms/step %.2f' % d['ms_per_step'], 'launch %.2f' % d['roofline']['avg_launch_ms'], 'e2e ms %.2f' % d['e2e']['ms_per_step'], d['clocks'])"
done
