#!/bin/bash
# round 2, final 1-GPU call: suite, smoke, bench at the driver's flags (both arms), launch list + ncu capture of the headline kernel
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2zz_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2zz_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2zz_smoke.log 2>&1
timeout 100 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2zz_bench_ref.json 2> $O/r2zz_bench_ref.err
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2zz_bench.json 2> $O/r2zz_bench.err
timeout 400 python bench.py > $O/r2zz_bench_default.json 2> $O/r2zz_bench_default.err
timeout 200 python tools/rk_bench.py 4096 tsit5,ssprk33 > $O/r2zz_rk.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2zz_launches.csv python bench.py --steps 20 --warmup 5 > $O/r2zz_bench_under_ncu.json 2> $O/r2zz_bench_under_ncu.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mol_rhs_tiled -c 3 -o $O/r2zz_tiled_full python bench.py --steps 3 --warmup 3 --no-extra > $O/r2zz_ncu.log 2>&1
tail -4 $O/r2zz_pytest.log; tail -1 $O/r2zz_smoke.log; cut -c1-400 $O/r2zz_bench_ref.json; cut -c1-3000 $O/r2zz_bench.json; cut -c1-300 $O/r2zz_bench_default.json; cat $O/r2zz_rk.log | tail -3
