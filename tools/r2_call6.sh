#!/bin/bash
# round 2, call 6 (1 GPU): staged per-node records (config 3, NU WENO)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2f_pytest.log
for c in "burgers2d_nu 4096" "burgers2d 4096" "weno1d_nu 1048576" "weno1d_nu 4194304" "weno2d_nu 2048" "nonlin1d 4194304"; do
  set -- $c
  timeout 200 python tools/rhs_bench.py $1 $2 > $O/r2f_$1_$2.log 2>&1
done
for m in 3 2; do
  MOL_TILE_MINCTAS=$m timeout 200 python tools/rhs_bench.py burgers2d_nu 4096 > $O/r2f_burgers2d_nu_4096_ctas$m.log 2>&1
done
for tx in 256 1024; do
  MOL_TILE_TX=$tx timeout 200 python tools/rhs_bench.py weno1d_nu 4194304 > $O/r2f_weno1d_nu_4194304_tx$tx.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mol_rhs_tiled -c 2 -o $O/r2f_burgers2d_nu_full python tools/rhs_bench.py burgers2d_nu 4096 > $O/r2f_ncu.log 2>&1
tail -12 $O/r2f_pytest.log; tail -qn 1 $O/r2f_*_*.log
