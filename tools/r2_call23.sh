#!/bin/bash
# round 2, call 23 (1 GPU): WENO5 rewritten in differences (61 instead of 87 FP64 operations): parity and timing
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "weno or WENO or burgers_weno or bench_size" > $O/r2y_pytest.log 2>&1; echo "rc=$?" >> $O/r2y_pytest.log
for c in "weno1d 4194304" "weno2d 4096" "weno2d_nu 2048" "weno1d_nu 4194304" "weno1d_burgers 4194304"; do
  set -- $c
  timeout 200 python tools/rhs_bench.py $1 $2 > $O/r2y_$1.log 2>&1
done
tail -4 $O/r2y_pytest.log; tail -qn 1 $O/r2y_weno*.log
