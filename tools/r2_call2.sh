#!/bin/bash
# round 2, call 2 (N GPUs): the driver's scaling protocol + slab parity tests
set -u
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for rep in 1 2; do
  timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r2b_bench_n${N}_rep$rep.json 2> $O/r2b_bench_n${N}_rep$rep.err
done
timeout 300 $TR bench.py --gpus $N --steps 500 --warmup 10 --no-extra > $O/r2b_bench_n${N}_k500.json 2> $O/r2b_bench_n${N}_k500.err
timeout 100 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $O/r2b_ref_n${N}.json 2>> $O/r2b_ref_n${N}.err
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x -p no:cacheprovider > $O/r2b_dist_pytest.log 2>&1
for f in $O/r2b_bench_n${N}_rep1.json $O/r2b_bench_n${N}_rep2.json $O/r2b_bench_n${N}_k500.json $O/r2b_ref_n${N}.json; do cat $f; done
tail -n 3 $O/r2b_bench_n${N}_rep1.err; tail -n 3 $O/r2b_dist_pytest.log
