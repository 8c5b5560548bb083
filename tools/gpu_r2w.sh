#!/bin/bash
# ncu evidence for the column-resident schedule (round 2, second half): launch list of a short bench run, full captures of the
# tuned soft-VI kernels at 1026^2 (100 sweeps in the captured launch)
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity --no-4096 --no-exp --vi-sweeps 40 --setup-sweeps 8 --paths 16777216"
timeout 300 $B > $O/plain_r2w.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_r2w.csv $B > $O/ncu_launches_r2w.log 2>&1
for C in "grid1024 nu f32fast" "grid1024 logz f32fast" "grid1024 gap f32fast" "grid1024 nu f32"; do
  set -- $C
  timeout 300 python tools/prof_vi.py $1 $2 $3 25 > $O/plainvi_$1_$2_$3_r2w.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"col_resident|tile_" -c 1 -o $O/prof_vi_$1_$2_$3_r2w python tools/prof_vi.py $1 $2 $3 25 > $O/ncu_vi_$1_$2_$3_r2w.log 2>&1
  tail -1 $O/plainvi_$1_$2_$3_r2w.log
done
ls -la $O/*r2w.ncu-rep
