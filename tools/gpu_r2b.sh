#!/bin/bash
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > $O/tests_$TAG.log 2>&1
(timeout 600 python tools/vi_tune.py grid1024 f32fast nu 2>&1 | tail -40) > $O/vitune_nu_$TAG.log 2>&1
(PIQL_B200_LIB=tools/probe/variants/libpiql_kh8.so timeout 300 python -m pytest tests/test_gpu_tiled.py -m gpu -q -x 2>&1 | tail -5) > $O/tests_kh8_$TAG.log 2>&1
(timeout 900 python tools/vi_ab.py time grid1024 grid4096 2>&1 | tail -10) > $O/viab_$TAG.log 2>&1
tail -3 $O/tests_$TAG.log; tail -3 $O/tests_kh8_$TAG.log; cat $O/viab_$TAG.log; head -12 $O/vitune_nu_$TAG.log; tail -3 $O/vitune_nu_$TAG.log
