#!/bin/bash
# round 2, call 27 (1 GPU): ncu --set full of the uniform WENO5 kernels after the rewrite in differences (evidence for profiles/)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mol_rhs_tiled -s 10 -c 2 -o $O/r2ac_weno2d_full python tools/rhs_bench.py weno2d 4096 > $O/r2ac_weno2d_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mol_rhs_tiled -s 10 -c 2 -o $O/r2ac_weno1d_full python tools/rhs_bench.py weno1d 4194304 > $O/r2ac_weno1d_ncu.log 2>&1
ls -la $O/r2ac_*.ncu-rep
