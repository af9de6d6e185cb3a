#!/bin/bash
# round 2, GPU call Z (two GPUs): multi-GPU checks and the N=2 bench line with the final kernels
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_fused_variants_z.jsonl
: > $O
python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_z.err
CUDDH_B200_AFFINE=0 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_z.err
CUDDH_B200_AFFINE=0 CUDDH_B200_RING=-5 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_z.err
cat $O
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 scripts/multi_check.py ) > gpurun_out/r02_multi2_final.log 2>&1
tail -n 10 gpurun_out/r02_multi2_final.log
( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q ) > gpurun_out/r02_pytest_z.log 2>&1
tail -n 4 gpurun_out/r02_pytest_z.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_z_n2.json 2> gpurun_out/r02_bench_z_n2.err
tail -c 1200 gpurun_out/r02_bench_z_n2.json; tail -n 4 gpurun_out/r02_bench_z_n2.err
