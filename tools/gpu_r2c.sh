#!/bin/bash
TAG=${1:-r2c}
O=gpurun_out
mkdir -p $O
(timeout 1200 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_production.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py tests/test_gpu_tables.py -m gpu -q -x 2>&1 | tail -25) > $O/tests_$TAG.log 2>&1
(timeout 900 python tools/vi_ab.py time grid1024 grid4096 2>&1 | tail -10) > $O/viab_$TAG.log 2>&1
(timeout 300 python tools/vi_acc.py 2>&1 | tail -12) > $O/viacc_$TAG.log 2>&1
tail -6 $O/tests_$TAG.log; cat $O/viab_$TAG.log; cat $O/viacc_$TAG.log
