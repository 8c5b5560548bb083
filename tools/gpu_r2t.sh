#!/bin/bash
O=gpurun_out
exec > $O/colres_fit.log 2>&1
export PIQL_TILE_COLRES=1 PIQL_TILE_NW=28
python tools/colres_time.py grid1024 nu f32fast 1 2 4 8 16 32 64 128
python tools/colres_time.py grid1024 logz f32fast 1 2 4 8 16 32 64 128
unset PIQL_TILE_NW
PIQL_TILE_COLRES=0 python tools/colres_time.py grid1024 nu f32fast 1 2 4 8 16 32 64 128
PIQL_TILE_COLRES=0 python tools/colres_time.py grid1024 logz f32fast 1 2 4 8 16 32 64 128
