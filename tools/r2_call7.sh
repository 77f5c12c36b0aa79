#!/bin/bash
# round 2, call 7 (1 GPU): field-major staged records
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_size_parity.py -m gpu -q -x -p no:cacheprovider -k "nu or weno or burgers or config" > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2g_pytest.log
for c in "burgers2d_nu 4096" "weno1d_nu 4194304" "weno2d_nu 2048"; do
  set -- $c
  timeout 200 python tools/rhs_bench.py $1 $2 > $O/r2g_$1_$2.log 2>&1
done
for m in 3 2; do
  MOL_TILE_MINCTAS=$m timeout 200 python tools/rhs_bench.py burgers2d_nu 4096 > $O/r2g_burgers2d_nu_4096_ctas$m.log 2>&1
done
for tx in 256 1024 2048; do
  MOL_TILE_TX=$tx timeout 200 python tools/rhs_bench.py weno1d_nu 4194304 > $O/r2g_weno1d_nu_4194304_tx$tx.log 2>&1
done
MOL_TILE_FORCE_TMA=1 timeout 200 python tools/rhs_bench.py weno2d_nu 2048 > $O/r2g_weno2d_nu_2048_tma.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mol_rhs_tiled -c 2 -o $O/r2g_burgers2d_nu_full python tools/rhs_bench.py burgers2d_nu 4096 > $O/r2g_ncu.log 2>&1
tail -4 $O/r2g_pytest.log; tail -qn 1 $O/r2g_*_*.log
