#!/bin/bash
# round 2, call 15 (1 GPU): Tsit5 PRE-stage residency (3 vs 4 CTAs/SM), NVRTC 12.8 (PyTorch's) vs 12.9 (toolkit's), memcheck
set -u
mkdir -p gpurun_out
O=gpurun_out
N129=/usr/local/cuda/lib64/libnvrtc.so.12
MOL_DEBUG_SPILL=1 timeout 200 python tools/rk_bench.py 4096 tsit5,ssprk33 > $O/r2p_rk_default.log 2>&1
MOL_TILE_PRE_MINCTAS=4 timeout 200 python tools/rk_bench.py 4096 tsit5 > $O/r2p_rk_pre4.log 2>&1
MOL_TILE_PRE_MINCTAS=2 timeout 200 python tools/rk_bench.py 4096 tsit5 > $O/r2p_rk_pre2.log 2>&1
MOL_TILE_CAP_RESIDENCY=1 MOL_TILE_MINCTAS=3 timeout 200 python tools/rk_bench.py 4096 tsit5,ssprk33 > $O/r2p_rk_all3.log 2>&1
MOL_NVRTC_PATH=$N129 MOL_DEBUG_SPILL=1 timeout 200 python tools/rk_bench.py 4096 tsit5,ssprk33 > $O/r2p_rk_nvrtc129.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2p_bench.json 2> $O/r2p_bench.err
MOL_NVRTC_PATH=$N129 timeout 300 python bench.py --steps 20 --warmup 5 --no-extra > $O/r2p_bench_nvrtc129.json 2> $O/r2p_bench_nvrtc129.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra > $O/r2p_bench_rep2.json 2> $O/r2p_bench_rep2.err
for c in "burgers2d_nu 4096" "weno2d 4096" "weno2d_nu 2048" "weno1d 4194304"; do
  set -- $c
  timeout 200 python tools/rhs_bench.py $1 $2 > $O/r2p_$1_default.log 2>&1
  MOL_NVRTC_PATH=$N129 timeout 200 python tools/rhs_bench.py $1 $2 > $O/r2p_$1_nvrtc129.log 2>&1
done
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "persistent or saveat or nu_weno or staged or step_to" > $O/r2p_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r2p_memcheck.log
for f in default pre4 pre2 all3 nvrtc129; do echo "== $f"; grep -v "^\[mol\].*spill stores" $O/r2p_rk_$f.log | tail -12; done
for f in r2p_bench r2p_bench_nvrtc129 r2p_bench_rep2; do python - $O/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['per_step_ms']['median'], 'e2e %.3g'%d['e2e']['value'], json.dumps(d.get('extra',{}))[:600])
except Exception as e:
    print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1200:])
PY
done
tail -n 3 $O/r2p_*_default.log $O/r2p_*_nvrtc129.log | grep -v "^$"
tail -5 $O/r2p_memcheck.log
