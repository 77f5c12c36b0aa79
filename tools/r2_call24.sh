#!/bin/bash
# round 2, call 24 (8 GPUs): the driver's scaling protocol at N = 8 on the final tree
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2n8_bench.json 2> $O/r2n8_bench.err
python - $O/r2n8_bench.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('ms/step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], d['per_step_ms']['median'], (d.get('dist_parity') or d.get('parity'))['max_rel'], d['gpu_launches'], 'e2e %.3g'%d['e2e']['value'], json.dumps(d.get('extra',{}))[:900])
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
