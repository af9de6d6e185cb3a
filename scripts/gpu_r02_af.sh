#!/bin/bash
# round 2, GPU call AF (four GPUs): multi-GPU checks and the N=4 bench line with the final kernels
set -u
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 scripts/multi_check.py ) > gpurun_out/r02_multi4_final.log 2>&1
tail -n 6 gpurun_out/r02_multi4_final.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 4 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_af_n4.json 2> gpurun_out/r02_bench_af_n4.err
tail -c 600 gpurun_out/r02_bench_af_n4.json; tail -n 4 gpurun_out/r02_bench_af_n4.err
