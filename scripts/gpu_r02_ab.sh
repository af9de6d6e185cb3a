#!/bin/bash
# round 2, GPU call AB (one GPU): fused two-phase thread-pair kernel (Helmholtz composite at n_basis 6-8 on affine meshes): parity + timing
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "composite or affine_and_stored or ring_variants or operators_match or volume" ) > gpurun_out/r02_pytest_ab.log 2>&1
tail -n 8 gpurun_out/r02_pytest_ab.log
H=gpurun_out/r02_high_order_ab.jsonl
: > $H
for nb in 8 6 7; do
  python scripts/time_high_order.py 1024 $nb >> $H 2>> gpurun_out/r02_ab.err
done
CUDDH_B200_FUSED=0 python scripts/time_high_order.py 1024 8 >> $H 2>> gpurun_out/r02_ab.err
cat $H | cut -c1-1500; tail -n 5 gpurun_out/r02_ab.err
