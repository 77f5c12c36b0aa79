#!/bin/bash
# tile-configuration sweep for the Brusselator 4096^2 RHS (GPU box); prints ms/step + roofline frac per config
run() {
  echo "== $*"
  env "$@" python bench.py --steps 200 --warmup 10 --cpu-seconds 0.1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('ms_per_step', round(d['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4), 'launches', d['gpu_launches'])
    elif 'rror' in l: print(l)
"
}
run MOL_X=0
run MOL_TILE_STAGES=2
run MOL_TILE_STAGES=4
run MOL_TILE_STAGES=4 MOL_TILE_MINCTAS=3
run MOL_TILE_STAGES=3 MOL_TILE_MINCTAS=3
run MOL_TILE_STAGES=3 MOL_TILE_MINCTAS=5
run MOL_TILE_TY=8 MOL_TILE_STAGES=4 MOL_TILE_MINCTAS=6
run MOL_TILE_TY=8 MOL_TILE_STAGES=5 MOL_TILE_MINCTAS=5
run MOL_TILE_TX=128 MOL_TILE_TY=8 MOL_TILE_STAGES=3 MOL_TILE_MINCTAS=4
run MOL_TILE_TX=128 MOL_TILE_TY=16 MOL_TILE_STAGES=3 MOL_TILE_MINCTAS=2
run MOL_TILE_TX=128 MOL_TILE_TY=16 MOL_TILE_STAGES=2 MOL_TILE_MINCTAS=3
run MOL_TILE_TX=64 MOL_TILE_TY=32 MOL_TILE_STAGES=3 MOL_TILE_MINCTAS=2 MOL_TILE_THREADS=512
