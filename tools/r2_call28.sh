#!/bin/bash
# round 2, call 28 (1 GPU): full GPU suite after the deterministic error norm (no -x: every test reports)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2ae_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2ae_pytest.log
tail -25 $O/r2ae_pytest.log | cut -c1-300
