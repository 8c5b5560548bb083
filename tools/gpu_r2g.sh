#!/bin/bash
TAG=${1:-r2g}
O=gpurun_out
mkdir -p $O
(timeout 1200 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_production.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -25) > $O/tests_$TAG.log 2>&1
(timeout 900 python tools/vi_ab.py time grid1024 grid4096 2>&1 | tail -10) > $O/viab_$TAG.log 2>&1
(PIQL_TILE_KH=4 timeout 600 python tools/vi_tune.py grid1024 f32fast nu 2>&1 | tail -30) > $O/vitune_nu_$TAG.log 2>&1
tail -6 $O/tests_$TAG.log; cat $O/viab_$TAG.log; head -8 $O/vitune_nu_$TAG.log; tail -2 $O/vitune_nu_$TAG.log
