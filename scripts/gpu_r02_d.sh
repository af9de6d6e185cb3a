#!/bin/bash
# round 2, GPU call D (four GPUs): distributed layer with interior ranks (two neighbours each) + the bench line at N = 4
set -u
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 scripts/multi_check.py ) > gpurun_out/r02_multi4.log 2>&1
tail -6 gpurun_out/r02_multi4.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
tail -c 1500 gpurun_out/r02_bench_n4.json
tail -4 gpurun_out/r02_bench_n4.err
