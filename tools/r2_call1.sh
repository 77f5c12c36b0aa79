#!/bin/bash
# round 2, call 1 (1 GPU): suite, smoke, bench at the driver's flags, pending A/Bs
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2a_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2a_smoke.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2a_bench.json 2> $O/r2a_bench.err
timeout 100 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2a_bench_ref.json 2>> $O/r2a_bench.err
for v in 0 1; do
  MOL_WENO_RATIO=$v timeout 200 python tools/rhs_bench.py weno2d 4096 > $O/r2a_weno2d_ratio$v.log 2>&1
done
timeout 200 python tools/rhs_bench.py burgers2d_nu 4096 > $O/r2a_burgers2d_nu.log 2>&1
timeout 200 python tools/rhs_bench.py weno1d_nu 1048576 > $O/r2a_weno1d_nu.log 2>&1
tail -3 $O/r2a_pytest.log; cat $O/r2a_smoke.log | tail -2; cat $O/r2a_bench.json; tail -5 $O/r2a_bench.err; cat $O/r2a_bench_ref.json; tail -1 $O/r2a_weno2d_ratio*.log $O/r2a_burgers2d_nu.log $O/r2a_weno1d_nu.log
