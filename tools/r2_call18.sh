#!/bin/bash
# round 2, call 18 (1 GPU): full GPU suite after the controller-kernel change of the adaptive loops
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2s_pytest.log 2>&1; echo "rc=$?" >> $O/r2s_pytest.log
timeout 200 python tools/config1_bench.py > $O/r2s_config1.log 2>&1
tail -8 $O/r2s_pytest.log; tail -4 $O/r2s_config1.log
