#!/bin/bash
# round 2, call 8 (1 GPU): persistent solver, config-3 tile variants, bench with the solve-shaped e2e
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2h_pytest.log
timeout 300 python tools/config1_bench.py > $O/r2h_config1.json 2> $O/r2h_config1.err
timeout 200 python tools/rhs_bench.py burgers2d_nu 4096 > $O/r2h_burgers2d_nu_4096.log 2>&1
MOL_TILE_TY=32 MOL_TILE_STAGES=2 MOL_TILE_MINCTAS=2 timeout 200 python tools/rhs_bench.py burgers2d_nu 4096 > $O/r2h_burgers2d_nu_4096_ty32.log 2>&1
MOL_TILE_TY=32 MOL_TILE_STAGES=2 MOL_TILE_MINCTAS=2 MOL_TILE_THREADS=512 timeout 200 python tools/rhs_bench.py burgers2d_nu 4096 > $O/r2h_burgers2d_nu_4096_ty32_t512.log 2>&1
MOL_TILE_TX=128 MOL_TILE_TY=16 MOL_TILE_STAGES=2 MOL_TILE_MINCTAS=2 timeout 200 python tools/rhs_bench.py burgers2d_nu 4096 > $O/r2h_burgers2d_nu_4096_tx128.log 2>&1
timeout 200 python tools/rhs_bench.py weno1d_nu 4194304 > $O/r2h_weno1d_nu_4194304.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2h_bench.json 2> $O/r2h_bench.err
tail -5 $O/r2h_pytest.log; cat $O/r2h_config1.json; tail -3 $O/r2h_config1.err; tail -qn 1 $O/r2h_burgers*.log $O/r2h_weno*.log; cut -c1-2500 $O/r2h_bench.json; tail -3 $O/r2h_bench.err
