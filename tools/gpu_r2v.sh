#!/bin/bash
O=gpurun_out
exec > $O/f64_variants.log 2>&1
for lib in product f64_384x2 f64_640x1 f64_256x3; do
  if [ $lib = product ]; then unset PIQL_B200_LIB; else export PIQL_B200_LIB=$PWD/tools/probe/variants/libpiql_$lib.so; fi
  echo "== $lib"
  python tools/colres_time.py grid4096 nu f64 40 120
  python tools/colres_time.py grid1024 nu f64 100 400
  python tools/colres_time.py grid4096 logz f64 40 120
done
