timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1p_pytest.log
run() { echo "== $*"; env "$@" 2>&1 | grep -E "us/RHS|rror" | tail -2; }
( run python tools/rhs_bench.py burgers2d_nu 1024
run python tools/rhs_bench.py burgers2d_nu 4096
run MOL_TILE_NO_CPASYNC=1 python tools/rhs_bench.py burgers2d_nu 4096
run MOL_TILE_TX=128 MOL_TILE_TY=8 python tools/rhs_bench.py burgers2d_nu 4096
run MOL_TILE_STAGES=2 python tools/rhs_bench.py burgers2d_nu 4096
run python tools/rhs_bench.py weno1d 4194304
run MOL_TILE_NO_CPASYNC=1 python tools/rhs_bench.py weno1d 4194304
run MOL_TILE_STAGES=3 python tools/rhs_bench.py weno1d 4194304
run MOL_TILE_TX=1024 MOL_TILE_STAGES=3 MOL_TILE_MINCTAS=4 python tools/rhs_bench.py weno1d 4194304
run python tools/rhs_bench.py nonlin1d 4194304
run python tools/rhs_bench.py weno2d 4096 ) > gpurun_out/r1p_configs.log 2>&1
cat gpurun_out/r1p_pytest.log gpurun_out/r1p_configs.log
