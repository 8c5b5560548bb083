#!/bin/bash
TAG=${1:-r2k}
O=gpurun_out
mkdir -p $O
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > $O/tests_$TAG.log 2>&1
R="python bench.py --workload grid4096 --steps 1 --warmup 3 --no-e2e --no-cpu --no-vi --no-exp --no-parity --no-4096 --setup-sweeps 8"
timeout 300 $R > $O/plainroll16_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o $O/prof_rollout_grid4096_rec16_$TAG $R > $O/ncu_rollout16_$TAG.log 2>&1
tail -4 $O/tests_$TAG.log; ls -la $O/prof_rollout_grid4096_rec16_$TAG.ncu-rep
