#!/bin/bash
# round 2, GPU call N (one GPU): fused affine kernel - per-thread cp.async ring for the mass rows (CUDDH_B200_AFFINE_RING=-3/-4/-5)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_affine_tring.jsonl
: > $O
for r in 5 -3 -4 -5 0; do
  CUDDH_B200_AFFINE_RING=$r python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_n.err
done
for r in 5 -4; do
  CUDDH_B200_AFFINE_RING=$r python scripts/fused_variant.py 1024 4 >> $O 2>> gpurun_out/r02_n.err
  CUDDH_B200_AFFINE_RING=$r python scripts/fused_variant.py 256 5 >> $O 2>> gpurun_out/r02_n.err
done
cat $O
CUDDH_B200_AFFINE_RING=-4 ncu --set full --clock-control none --import-source on -k regex:volume_action_ws -s 3 -c 1 -f -o gpurun_out/r02_prof_affine_tring python scripts/time_fused.py 5 > gpurun_out/r02_ncu_n1.log 2>&1
tail -n 2 gpurun_out/r02_ncu_n1.log; tail -n 5 gpurun_out/r02_n.err
