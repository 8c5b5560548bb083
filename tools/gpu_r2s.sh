#!/bin/bash
O=gpurun_out
export PIQL_TILE_COLRES=1 PIQL_TILE_NW=28 PIQL_TILE_COLRES_TRACE=70
PIQL_TILE_COLRES_TRACE_SW=300 timeout 300 python tools/prof_vi.py grid1024 nu f32fast 100 > $O/trace_nu.log 2>&1
PIQL_TILE_COLRES_TRACE_SW=300 PIQL_TILE_COLRES_DIAG=1 timeout 300 python tools/prof_vi.py grid1024 nu f32fast 100 > $O/trace_nu_nomail.log 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/smi.log
