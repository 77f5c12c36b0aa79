#!/bin/bash
# round 2, call 4 (1 GPU): suite after the NU WENO redesign + RK dense output; WENO perf
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2d_pytest.log
for c in "weno1d_nu 1048576" "weno1d_nu 4194304" "weno2d_nu 2048" "weno2d 4096" "weno1d 4194304"; do
  set -- $c
  timeout 200 python tools/rhs_bench.py $1 $2 > $O/r2d_$1_$2.log 2>&1
done
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2d_bench.json 2> $O/r2d_bench.err
tail -15 $O/r2d_pytest.log; tail -qn 1 $O/r2d_weno*.log; cut -c1-600 $O/r2d_bench.json
