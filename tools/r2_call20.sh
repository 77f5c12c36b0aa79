#!/bin/bash
# round 2, call 20 (2 GPUs): slab-mode tests and the bench line after the controller-kernel / PRE-residency changes
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x -p no:cacheprovider > $O/r2v_dist_pytest.log 2>&1; echo "rc=$?" >> $O/r2v_dist_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552"
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2v_bench_n2.json 2> $O/r2v_bench_n2.err
tail -4 $O/r2v_dist_pytest.log
python - $O/r2v_bench_n2.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('ms/step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], d['per_step_ms']['median'], (d.get('dist_parity') or d.get('parity'))['max_rel'], d['gpu_launches'], 'e2e %.3g'%d['e2e']['value'], json.dumps(d.get('extra',{}))[:900])
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
