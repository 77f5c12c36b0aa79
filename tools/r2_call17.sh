#!/bin/bash
# round 2, call 17 (1 GPU): both adaptive loops through the one controller kernel: queued vs host-driven, RK tests, timing
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python tools/queued_debug.py 64 > $O/r2r_debug.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_reference_acceptance.py -m gpu -q -x -p no:cacheprovider > $O/r2r_pytest.log 2>&1; echo "rc=$?" >> $O/r2r_pytest.log
for n in 128 512 1024; do timeout 300 python tools/solve_bench.py $n 400 > $O/r2r_solve_$n.log 2>&1; done
tail -14 $O/r2r_debug.log; tail -6 $O/r2r_pytest.log; tail -n 4 $O/r2r_solve_*.log
