#!/bin/bash
# round 2, call 5 (1 GPU): suite after the WENO issue-slot work; WENO perf; ncu capture of the NU WENO kernel
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2e_pytest.log
for c in "weno1d_nu 1048576" "weno1d_nu 4194304" "weno1d 4194304" "weno2d 4096" "weno2d_nu 2048" "burgers2d_nu 4096"; do
  set -- $c
  timeout 200 python tools/rhs_bench.py $1 $2 > $O/r2e_$1_$2.log 2>&1
done
for m in 3 2; do
  MOL_TILE_MINCTAS=$m timeout 200 python tools/rhs_bench.py weno2d_nu 2048 > $O/r2e_weno2d_nu_2048_ctas$m.log 2>&1
  MOL_TILE_MINCTAS=$m timeout 200 python tools/rhs_bench.py weno2d 4096 > $O/r2e_weno2d_4096_ctas$m.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mol_rhs_tiled -c 2 -o $O/r2e_weno1d_nu_full python tools/rhs_bench.py weno1d_nu 4194304 > $O/r2e_ncu.log 2>&1
tail -12 $O/r2e_pytest.log; tail -qn 1 $O/r2e_*_*.log
