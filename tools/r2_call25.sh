#!/bin/bash
# round 2, call 25 (1 GPU): cp.async staging loop without per-cell divisions: timing of the cp.async programs, full GPU suite, bench
set -u
mkdir -p gpurun_out
O=gpurun_out
for c in "burgers2d_nu 4096" "weno1d 4194304" "weno1d_nu 4194304" "nonlin1d 4194304" "burgers2d 4097"; do
  set -- $c
  timeout 200 python tools/rhs_bench.py $1 $2 > $O/r2aa_$1.log 2>&1
done
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2aa_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2aa_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2aa_bench.json 2> $O/r2aa_bench.err
tail -qn 1 $O/r2aa_burgers2d_nu.log $O/r2aa_weno1d.log $O/r2aa_weno1d_nu.log $O/r2aa_nonlin1d.log $O/r2aa_burgers2d.log; tail -3 $O/r2aa_pytest.log; cut -c1-330 $O/r2aa_bench.json
