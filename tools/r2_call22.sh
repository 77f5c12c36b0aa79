#!/bin/bash
# round 2, call 22 (1 GPU): memcheck after clamping the grid-coordinate reads of overhanging tile cells
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_zz_reference_acceptance.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "jvp or nonlinear or robin or late or scheme_parity" > $O/r2x_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r2x_memcheck.log
grep -c "Invalid" $O/r2x_memcheck.log; grep "at \|ERROR SUMMARY\|passed\|failed" $O/r2x_memcheck.log | sort | uniq -c | sort -rn | head -12; tail -3 $O/r2x_memcheck.log
