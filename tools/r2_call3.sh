#!/bin/bash
# round 2, call 3 (8 GPUs): the driver's scaling protocol at N=8 (and N=4), headline only
set -u
mkdir -p gpurun_out
O=gpurun_out
for N in 8 4; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r2c_bench_n${N}.json 2> $O/r2c_bench_n${N}.err
timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-extra > $O/r2c_bench_n${N}_rep2.json 2> $O/r2c_bench_n${N}_rep2.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 500 --warmup 10 --no-extra > $O/r2c_bench_n8_k500.json 2> $O/r2c_bench_n8_k500.err
nproc > $O/r2c_nproc.txt
cat $O/r2c_bench_n8.json $O/r2c_bench_n8_rep2.json $O/r2c_bench_n4.json $O/r2c_bench_n4_rep2.json $O/r2c_bench_n8_k500.json | cut -c1-900
