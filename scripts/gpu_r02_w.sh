#!/bin/bash
# round 2, GPU call W (one GPU): the whole GPU test-suite (no -x), n_basis 9 stored-metric stiffness on the lane-per-row kernel
set -u
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_w.log 2>&1
tail -n 12 gpurun_out/r02_pytest_w.log
CUDDH_B200_AFFINE=0 CUDDH_B200_PAIR=0 python scripts/time_high_order.py 1024 9 > gpurun_out/r02_high_order_w.jsonl 2>> gpurun_out/r02_w.err
cat gpurun_out/r02_high_order_w.jsonl | cut -c1-700
