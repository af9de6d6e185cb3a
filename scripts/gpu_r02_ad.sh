#!/bin/bash
# round 2, GPU call AD (one GPU): thread-pair kernel with 2 interior ids per element (n_basis 9: metric ring + fused composite): parity + timing
set -u
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -k "composite or affine_and_stored or operator or steady or volume or mass or stiff" ) > gpurun_out/r02_pytest_ad.log 2>&1
tail -n 8 gpurun_out/r02_pytest_ad.log
H=gpurun_out/r02_high_order_ad.jsonl
: > $H
for nb in 9 8 6; do
  python scripts/time_high_order.py 1024 $nb >> $H 2>> gpurun_out/r02_ad.err
done
cat $H | cut -c1-1800; tail -n 5 gpurun_out/r02_ad.err
