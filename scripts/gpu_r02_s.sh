#!/bin/bash
# round 2, GPU call S (one GPU): pair mass kernels with ring + U in registers; ncu: launch list of the bench command, full captures of
# the default fused kernel (affine, per-thread ring), of the stored-metric fused kernel with the per-thread chunk ring and of the
# stored-metric thread-pair stiffness kernel at n_basis 8
set -u
mkdir -p gpurun_out
H=gpurun_out/r02_high_order_s.jsonl
: > $H
for nb in 8 7; do
  CUDDH_B200_AFFINE=0 python scripts/time_high_order.py 1024 $nb >> $H 2>> gpurun_out/r02_s.err
done
cat $H
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-ddh"
$CMD > gpurun_out/r02_plain_s.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_s.csv $CMD > gpurun_out/r02_ncu_list_s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:volume_action_ws -s 4 -c 1 -f -o gpurun_out/r02_prof_fused_final $CMD > gpurun_out/r02_ncu_s1.log 2>&1
CUDDH_B200_AFFINE=0 CUDDH_B200_RING=-5 ncu --set full --clock-control none --import-source on -k regex:volume_action_ws -s 3 -c 1 -f -o gpurun_out/r02_prof_fused_stored_trg python scripts/time_fused.py 5 > gpurun_out/r02_ncu_s2.log 2>&1
CUDDH_B200_AFFINE=0 ncu --set full --clock-control none --import-source on -k regex:volume_action_pair -s 4 -c 1 -f -o gpurun_out/r02_prof_pair_s8_stored python scripts/time_high_order.py 1024 8 > gpurun_out/r02_ncu_s3.log 2>&1
tail -n 2 gpurun_out/r02_plain_s.log | cut -c1-600; tail -n 2 gpurun_out/r02_ncu_s1.log gpurun_out/r02_ncu_s2.log gpurun_out/r02_ncu_s3.log; tail -n 5 gpurun_out/r02_s.err
