#!/bin/bash
# round 2, GPU call L (eight GPUs): peer-memory halo exchange - checks + bench at N = 8
set -u
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 scripts/multi_check.py ) > gpurun_out/r02_multi8_peer.log 2>&1
tail -8 gpurun_out/r02_multi8_peer.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_n8_peer.json 2> gpurun_out/r02_bench_n8_peer.err
tail -c 1500 gpurun_out/r02_bench_n8_peer.json; tail -4 gpurun_out/r02_bench_n8_peer.err
