#!/bin/bash
# round 2, GPU call C (two GPUs): distributed layer checks + the bench line at N = 2
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_c_gpus.txt 2>&1
( time timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -q -k "two_gpus or emulated_ranks" ) > gpurun_out/r02_pytest_c.log 2>&1
tail -30 gpurun_out/r02_pytest_c.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -c 2500 gpurun_out/r02_bench_n2.json
tail -8 gpurun_out/r02_bench_n2.err
