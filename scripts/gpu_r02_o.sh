#!/bin/bash
# round 2, GPU call O (one GPU): (1) fused affine kernel with the per-thread ring as default + split cluster barrier; (2) thread-pair
# kernel for n_basis 6-9: parity (reference kernels, oracle) and timings against the lane-per-row kernel
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_affine_tring2.jsonl
: > $O
for r in -4 -5 5; do
  CUDDH_B200_AFFINE_RING=$r python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_o.err
done
CUDDH_B200_AFFINE=0 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_o.err
python scripts/fused_variant.py 1024 4 >> $O 2>> gpurun_out/r02_o.err
cat $O
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -x ) > gpurun_out/r02_pytest_o.log 2>&1
tail -n 8 gpurun_out/r02_pytest_o.log
H=gpurun_out/r02_high_order.jsonl
: > $H
for nb in 8 9 6 7; do
  python scripts/time_high_order.py 1024 $nb >> $H 2>> gpurun_out/r02_o.err
  CUDDH_B200_AFFINE=0 python scripts/time_high_order.py 1024 $nb >> $H 2>> gpurun_out/r02_o.err
done
CUDDH_B200_PAIR=0 python scripts/time_high_order.py 1024 8 >> $H 2>> gpurun_out/r02_o.err
cat $H; tail -n 5 gpurun_out/r02_o.err
