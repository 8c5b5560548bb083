#!/bin/bash
# round 2, first pass: new parity tests + bench with parity + ROLL_TILED check
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/gpus_$TAG.log 2>&1
(timeout 1200 python -m pytest tests -m gpu -q -x --durations=15 2>&1 | tail -40) > $O/tests_$TAG.log 2>&1
timeout 600 python bench.py > $O/bench1024_$TAG.json 2> $O/bench1024_$TAG.err
(timeout 600 python tools/roll_tiled_check.py grid4096 2>&1 | tail -8) > $O/rolltiled_$TAG.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/benchref_$TAG.json 2> $O/benchref_$TAG.err
tail -5 $O/tests_$TAG.log; cat $O/rolltiled_$TAG.log; tail -c 600 $O/bench1024_$TAG.err
