#!/bin/bash
# First GPU call of the next round (single B200): everything that was written without a GPU, then the open measurements.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Outputs land in gpurun_out/r2a_*.  Nothing here changes clocks; every step has its own timeout.
set -u
mkdir -p gpurun_out
O=gpurun_out
# 1. the whole GPU suite (113 tests; the last ~50 of tests/test_zz_reference_acceptance.py have never run on a GPU)
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2a_pytest.log
#    ... and without -x, so that one failure does not hide the rest
timeout 900 python -m pytest tests/test_zz_reference_acceptance.py -m gpu -q -p no:cacheprovider > $O/r2a_pytest_zz_all.log 2>&1
# 2. smoke + bench (the driver's contract)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2a_smoke.log 2>&1
timeout 600 python bench.py > $O/r2a_bench.json 2> $O/r2a_bench.err
# 3. A/B of the opt-in WENO5 weights (NOTES.md): 2-D 4096^2 and 1-D 2^22
for v in 0 1; do
  MOL_WENO_RATIO=$v timeout 300 python tools/rhs_bench.py weno2d 4096 > $O/r2a_weno2d_ratio$v.log 2>&1
  MOL_WENO_RATIO=$v timeout 300 python tools/rhs_bench.py weno1d 4194304 > $O/r2a_weno1d_ratio$v.log 2>&1
done
# 4. the non-uniform 2-D kernel today (baseline for the shared-memory weight staging, profiles/r01_nu_tiled_ncu.md)
timeout 300 python tools/rhs_bench.py burgers2d_nu 4096 > $O/r2a_burgers2d_nu.log 2>&1
# 5. time to the first Tsit5 step with the threaded precompile (and with one thread)
for th in 0 1; do
  if [ $th = 1 ]; then export MOL_COMPILE_THREADS=1; else unset MOL_COMPILE_THREADS; fi
  timeout 300 python tools/solve_bench.py 512 2e-4 > $O/r2a_solve_threads$th.log 2>&1
done
tail -3 $O/r2a_pytest.log; cat $O/r2a_bench.json
