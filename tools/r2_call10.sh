#!/bin/bash
# round 2, call 10 (1 GPU): full suite, persistent solver after the frame fix, config 3 default, smoke + bench
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2j_pytest.log
timeout 300 python tools/config1_bench.py > $O/r2j_config1.json 2> $O/r2j_config1.err
timeout 200 python tools/rhs_bench.py burgers2d_nu 4096 > $O/r2j_burgers2d_nu_4096.log 2>&1
timeout 200 python tools/rhs_bench.py weno2d_nu 2048 > $O/r2j_weno2d_nu_2048.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2j_smoke.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2j_bench.json 2> $O/r2j_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mol_rhs_tiled -c 2 -o $O/r2j_burgers2d_nu_full python tools/rhs_bench.py burgers2d_nu 4096 > $O/r2j_ncu.log 2>&1
tail -5 $O/r2j_pytest.log; cat $O/r2j_config1.json; tail -3 $O/r2j_config1.err; tail -qn 1 $O/r2j_burgers*.log $O/r2j_weno*.log $O/r2j_smoke.log; cut -c1-300 $O/r2j_bench.json; tail -3 $O/r2j_bench.err
