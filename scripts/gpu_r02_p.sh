#!/bin/bash
# round 2, GPU call P (one GPU): parity of everything; cluster-barrier mode per instance; per-thread chunk ring (TRG) on the stored-metric
# fused / stand-alone instances; thread-pair kernel timings after the ILP restructure + ncu captures at n_basis 8
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -x ) > gpurun_out/r02_pytest_p.log 2>&1
tail -n 6 gpurun_out/r02_pytest_p.log
O=gpurun_out/r02_fused_variants_p.jsonl
: > $O
python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_p.err
CUDDH_B200_AFFINE_RING=5 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_p.err
CUDDH_B200_AFFINE_RING=-5 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_p.err
CUDDH_B200_AFFINE=0 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_p.err
CUDDH_B200_AFFINE=0 CUDDH_B200_RING=-5 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_p.err
CUDDH_B200_AFFINE=0 CUDDH_B200_RING=-5 CUDDH_B200_CSPLIT=0 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_p.err
CUDDH_B200_AFFINE=0 python scripts/fused_variant.py 1024 4 >> $O 2>> gpurun_out/r02_p.err
CUDDH_B200_AFFINE=0 CUDDH_B200_RING4=-5 python scripts/fused_variant.py 1024 4 >> $O 2>> gpurun_out/r02_p.err
cat $O
T=gpurun_out/r02_time_ops_p.jsonl
: > $T
CUDDH_B200_AFFINE=0 python scripts/time_ops.py 1024 >> $T 2>> gpurun_out/r02_p.err
CUDDH_B200_AFFINE=0 CUDDH_B200_RING1=-5 CUDDH_B200_RING4=-5 python scripts/time_ops.py 1024 >> $T 2>> gpurun_out/r02_p.err
cat $T
H=gpurun_out/r02_high_order_p.jsonl
: > $H
for nb in 8 9; do
  python scripts/time_high_order.py 1024 $nb >> $H 2>> gpurun_out/r02_p.err
  CUDDH_B200_AFFINE=0 python scripts/time_high_order.py 1024 $nb >> $H 2>> gpurun_out/r02_p.err
done
cat $H
ncu --set full --clock-control none --import-source on -k regex:volume_action_pair -s 4 -c 1 -f -o gpurun_out/r02_prof_pair_s8 python scripts/time_high_order.py 1024 8 > gpurun_out/r02_ncu_p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:volume_action_pair -s 18 -c 1 -f -o gpurun_out/r02_prof_pair_m8 python scripts/time_high_order.py 1024 8 > gpurun_out/r02_ncu_p2.log 2>&1
tail -n 2 gpurun_out/r02_ncu_p1.log gpurun_out/r02_ncu_p2.log; tail -n 5 gpurun_out/r02_p.err
