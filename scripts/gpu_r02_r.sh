#!/bin/bash
# round 2, GPU call R (one GPU): thread-pair mass kernels with the per-thread metric ring; fused affine kernel with the mass scale folded
# into the back-contraction table and ring depth 5; parity of the operator tests
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_fused_variants_r.jsonl
: > $O
python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_r.err
CUDDH_B200_AFFINE_RING=-4 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_r.err
CUDDH_B200_AFFINE=0 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_r.err
python scripts/fused_variant.py 1024 4 >> $O 2>> gpurun_out/r02_r.err
CUDDH_B200_AFFINE=0 python scripts/fused_variant.py 1024 4 >> $O 2>> gpurun_out/r02_r.err
cat $O
H=gpurun_out/r02_high_order_r.jsonl
: > $H
for nb in 8 6 7; do
  CUDDH_B200_AFFINE=0 python scripts/time_high_order.py 1024 $nb >> $H 2>> gpurun_out/r02_r.err
done
cat $H
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -x -k "operator or steady or composite or volume or mass or stiff" ) > gpurun_out/r02_pytest_r.log 2>&1
tail -n 5 gpurun_out/r02_pytest_r.log
ncu --set full --clock-control none --import-source on -k regex:volume_action_pair -s 18 -c 1 -f -o gpurun_out/r02_prof_pair_m8 python scripts/time_high_order.py 1024 8 > gpurun_out/r02_ncu_r2.log 2>&1
tail -n 2 gpurun_out/r02_ncu_r2.log
tail -n 5 gpurun_out/r02_r.err
