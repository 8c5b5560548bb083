#!/bin/bash
O=gpurun_out
export PIQL_TILE_COLRES=1 PIQL_TILE_NW=32
timeout 300 python tools/prof_vi.py grid1024 nu f32fast 50 > $O/plain_colres_nu.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:col_resident -c 1 -o $O/prof_colres_nu python tools/prof_vi.py grid1024 nu f32fast 50 > $O/ncu_colres_nu.log 2>&1
tail -3 $O/plain_colres_nu.log $O/ncu_colres_nu.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:col_resident -c 1 -o $O/prof_colres_logz python tools/prof_vi.py grid1024 logz f32fast 50 > $O/ncu_colres_logz.log 2>&1
