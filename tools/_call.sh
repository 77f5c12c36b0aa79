timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1n_pytest.log
run() { echo "== $*"; env "$@" 2>&1 | grep -E "us/RHS|rror" | tail -2; }
( run python tools/rhs_bench.py burgers2d_nu 1024
run python tools/rhs_bench.py burgers2d_nu 4096
run MOL_TILE_MINCTAS=4 python tools/rhs_bench.py burgers2d_nu 4096
run MOL_TILE_MINCTAS=3 python tools/rhs_bench.py burgers2d_nu 4096
run python tools/rhs_bench.py burgers2d 4096
run MOL_TILE_MINCTAS=4 python tools/rhs_bench.py burgers2d 4096
run MOL_TILE_MINCTAS=3 python tools/rhs_bench.py burgers2d 4096
run python tools/rhs_bench.py weno1d 4194304
run python tools/rhs_bench.py weno2d 4096
run MOL_TILE_MINCTAS=4 python tools/rhs_bench.py weno2d 4096
run python tools/rhs_bench.py bruss 4096 ) > gpurun_out/r1n_configs.log 2>&1
cat gpurun_out/r1n_pytest.log gpurun_out/r1n_configs.log
