#!/bin/bash
# round 2, call 16 (1 GPU): queued adaptive Tsit5 (device-side step control): parity vs the host loop, timing, memcheck
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "queued or saveat or solve or step_to or persistent" > $O/r2q_pytest.log 2>&1; echo "rc=$?" >> $O/r2q_pytest.log
for n in 128 512 1024; do timeout 300 python tools/solve_bench.py $n 400 > $O/r2q_solve_$n.log 2>&1; done
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "queued" > $O/r2q_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r2q_memcheck.log
tail -15 $O/r2q_pytest.log; tail -n 5 $O/r2q_solve_*.log; tail -4 $O/r2q_memcheck.log
