#!/bin/bash
# 3-D kernel sweep: Fisher-KPP 7-point, 512^3, periodic (z-march ring/chunk/tile/L2 prefetch distance vs the brick kernel)
run() { echo "== $*"; env "$@" python tools/rhs_bench.py fisher3d 512 2>&1 | grep -E "us/RHS|rror" | tail -2; }
run MOL_X=0
run MOL_TILE_ZMARCH=0
run MOL_TILE_L2AHEAD=0
run MOL_TILE_L2AHEAD=3
run MOL_TILE_L2AHEAD=12
run MOL_TILE_L2AHEAD=12 MOL_TILE_TZ=64
run MOL_TILE_L2AHEAD=6 MOL_TILE_TZ=64
run MOL_TILE_L2AHEAD=6 MOL_TILE_TZ=16
run MOL_TILE_L2AHEAD=6 MOL_TILE_RING=4
run MOL_TILE_L2AHEAD=6 MOL_TILE_RING=4 MOL_TILE_MINCTAS=5
run MOL_TILE_L2AHEAD=6 MOL_TILE_TX=128 MOL_TILE_TY=8
run MOL_TILE_L2AHEAD=6 MOL_TILE_TX=128 MOL_TILE_TY=8 MOL_TILE_TZ=64
run MOL_TILE_L2AHEAD=6 MOL_TILE_TX=128 MOL_TILE_TY=8 MOL_TILE_RING=4
