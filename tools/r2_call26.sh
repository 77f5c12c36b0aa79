#!/bin/bash
# round 2, call 26 (1 GPU): dual-number WENO5 in difference form with division-free weight ratios: J*v parity and throughput
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_zz_reference_acceptance.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "jvp" > $O/r2ab_pytest.log 2>&1; echo "rc=$?" >> $O/r2ab_pytest.log
timeout 200 python tools/jvp_bench.py > $O/r2ab_jvp.log 2>&1
tail -3 $O/r2ab_pytest.log; tail -3 $O/r2ab_jvp.log
