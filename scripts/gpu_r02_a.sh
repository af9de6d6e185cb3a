#!/bin/bash
# round 2, GPU call A (one GPU): full GPU test suite, then the default bench line
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r02_gpu.txt 2>&1
nproc >> gpurun_out/r02_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -x --durations=15 ) > gpurun_out/r02_pytest.log 2>&1
tail -40 gpurun_out/r02_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
tail -c 3000 gpurun_out/r02_bench_a.json
tail -5 gpurun_out/r02_bench_a.err
