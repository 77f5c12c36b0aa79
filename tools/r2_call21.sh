#!/bin/bash
# round 2, call 21 (1 GPU): tiled J*v -- parity against the table-driven kernel, throughput, memcheck
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_zz_reference_acceptance.py -m gpu -q -x -p no:cacheprovider -k "jvp" > $O/r2w_pytest.log 2>&1; echo "rc=$?" >> $O/r2w_pytest.log
timeout 300 python tools/jvp_bench.py > $O/r2w_jvp_tiled.log 2>&1
MOL_JVP_GENERIC=1 timeout 300 python tools/jvp_bench.py > $O/r2w_jvp_generic.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_reference_acceptance.py -m gpu -q -x -p no:cacheprovider -k "tiled_jvp" > $O/r2w_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r2w_memcheck.log
tail -6 $O/r2w_pytest.log; cat $O/r2w_jvp_tiled.log | tail -4; cat $O/r2w_jvp_generic.log | tail -4; tail -3 $O/r2w_memcheck.log
