#!/bin/bash
# round 2, call 19 (1 GPU): programmatic dependent launch in the queued adaptive solve (on / off), launch list, J*v throughput
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "queued" > $O/r2u_pytest.log 2>&1; echo "rc=$?" >> $O/r2u_pytest.log
for n in 128 512 1024; do
  MOL_RK_PDL=1 timeout 200 python tools/solve_bench.py $n 400 > $O/r2u_solve_${n}_pdl1.log 2>&1
  MOL_RK_PDL=0 timeout 200 python tools/solve_bench.py $n 400 > $O/r2u_solve_${n}_pdl0.log 2>&1
done
MOL_RK_QUEUED_BATCH=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 300 --csv --log-file $O/r2u_solve128_launches.csv python tools/solve_bench.py 128 40 > $O/r2u_ncu.log 2>&1
timeout 300 python tools/jvp_bench.py > $O/r2u_jvp.log 2>&1
timeout 200 python tools/rhs_bench.py nonlin1d 4194304 > $O/r2u_nonlin1d.log 2>&1
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "queued" > $O/r2u_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r2u_memcheck.log
tail -5 $O/r2u_pytest.log; tail -n 2 $O/r2u_solve_*.log; tail -4 $O/r2u_jvp.log; tail -2 $O/r2u_nonlin1d.log; tail -3 $O/r2u_memcheck.log; tail -2 $O/r2u_ncu.log
