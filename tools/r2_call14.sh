#!/bin/bash
# round 2, call 14 (8 GPUs): the driver's scaling protocol at N = 8, 4 with the fused ghost-plane wait
set -u
mkdir -p gpurun_out
O=gpurun_out
for N in 8 4; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r2n_bench_n${N}.json 2> $O/r2n_bench_n${N}.err
timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-extra > $O/r2n_bench_n${N}_rep2.json 2> $O/r2n_bench_n${N}_rep2.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2n_bench_n2.json 2> $O/r2n_bench_n2.err
timeout 200 python bench.py --steps 20 --warmup 5 > $O/r2n_bench_n1.json 2> $O/r2n_bench_n1.err
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > $O/r2n_ref_n8.json 2> $O/r2n_ref_n8.err
timeout 500 python -m pytest tests/test_gpu_dist.py -m gpu -q -x -p no:cacheprovider > $O/r2n_dist_pytest.log 2>&1
tail -2 $O/r2n_dist_pytest.log
for f in n1 n2 n4 n4_rep2 n8 n8_rep2; do python - $O/r2n_bench_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], d['per_step_ms']['median'], (d.get('dist_parity') or d.get('parity'))['max_rel'], d['gpu_launches'], 'e2e %.3g'%d['e2e']['value'], json.dumps(d.get('extra',{}))[:420])
except Exception as e:
    print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1200:])
PY
done
cut -c1-400 $O/r2n_ref_n8.json
