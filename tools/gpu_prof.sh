#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run, full captures of the tiled nu kernel (fp32 fast) on both grids
# with the shapes the tuner chose (pinned by env so that the skip count is known).  usage (under gpurun): bash tools/gpu_prof.sh <tag>
TAG=${1:-prof}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-exp --no-vi --setup-sweeps 8 --paths 16777216"
PIQL_TILE_NO_TUNE=1 timeout 200 $B > $O/plain_$TAG.log 2>&1 && \
PIQL_TILE_NO_TUNE=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$TAG.csv $B > $O/ncu_launches_$TAG.log 2>&1
V="--steps 1 --warmup 3 --no-e2e --no-cpu --no-exp --vi-sweeps 40 --setup-sweeps 8 --paths 1048576"
PIQL_TILE_TJ=54 PIQL_TILE_NW=12 timeout 200 python bench.py --workload grid4096 $V > $O/plain4096_$TAG.log 2>&1 && \
PIQL_TILE_TJ=54 PIQL_TILE_NW=12 timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_sweep_kernel -s 8 -c 1 -f -o $O/prof_tile4096_$TAG python bench.py --workload grid4096 $V > $O/ncu_tile4096_$TAG.log 2>&1
PIQL_TILE_TJ=22 PIQL_TILE_NW=12 timeout 200 python bench.py --workload grid1024 $V > $O/plain1024_$TAG.log 2>&1 && \
PIQL_TILE_TJ=22 PIQL_TILE_NW=12 timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_sweep_kernel -s 8 -c 1 -f -o $O/prof_tile1024_$TAG python bench.py --workload grid1024 $V > $O/ncu_tile1024_$TAG.log 2>&1
ls -la $O | tail -12
