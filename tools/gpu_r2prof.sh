#!/bin/bash
# ncu evidence, round 2 (one GPU): launch list of a short bench run, full captures of the rollout kernel at both grids (launch
# size of the bench) and of the tuned soft-VI kernels
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity --no-4096 --no-exp --vi-sweeps 40 --setup-sweeps 8 --paths 16777216"
timeout 300 $B > $O/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$TAG.csv $B > $O/ncu_launches_$TAG.log 2>&1
for W in grid1024 grid4096; do
  R="python bench.py --workload $W --steps 1 --warmup 3 --no-e2e --no-cpu --no-vi --no-exp --no-parity --no-4096 --setup-sweeps 8"
  timeout 300 $R > $O/plainroll_${W}_$TAG.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o $O/prof_rollout_${W}_$TAG $R > $O/ncu_rollout_${W}_$TAG.log 2>&1
done
for C in "grid1024 nu f32fast" "grid1024 logz f32fast" "grid1024 gap f32fast" "grid4096 nu f32fast" "grid4096 logz f32fast" "grid4096 nu f64"; do
  set -- $C
  timeout 300 python tools/prof_vi.py $1 $2 $3 > $O/plainvi_$1_$2_$3_$TAG.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tile_ -c 2 -o $O/prof_vi_$1_$2_$3_$TAG python tools/prof_vi.py $1 $2 $3 > $O/ncu_vi_$1_$2_$3_$TAG.log 2>&1
done
ls -la $O/*.ncu-rep | tail -12
