#!/bin/bash
# round 2, call 13 (2 GPUs): fused ghost-plane wait on/off
set -u
mkdir -p gpurun_out
O=gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x -p no:cacheprovider > $O/r2m_dist_pytest.log 2>&1
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r2m_bench_n2_fused.json 2> $O/r2m_bench_n2_fused.err
timeout 300 $TR bench.py --gpus $N --steps 500 --warmup 10 --no-extra > $O/r2m_bench_n2_fused_k500.json 2> $O/r2m_bench_n2_fused_k500.err
MOL_DIST_FUSED=0 timeout 300 $TR bench.py --gpus $N --steps 500 --warmup 10 --no-extra > $O/r2m_bench_n2_unfused_k500.json 2> $O/r2m_bench_n2_unfused_k500.err
tail -3 $O/r2m_dist_pytest.log
for f in $O/r2m_bench_n2_fused.json $O/r2m_bench_n2_fused_k500.json $O/r2m_bench_n2_unfused_k500.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step %.4f'%d['ms_per_step'], d['per_step_ms'], d.get('dist_parity',{}).get('max_rel'), d['gpu_launches'], d.get('e2e',{}).get('value'), json.dumps(d.get('extra',{}))[:600])
except Exception as e:
    print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
