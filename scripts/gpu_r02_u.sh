#!/bin/bash
# round 2, GPU call U (one GPU): full ncu captures of the final thread-pair kernels at n_basis 8 (affine stiffness, stored-metric
# stiffness, weighted mass) and stand-alone operator timings at n_basis 4 / 5
set -u
mkdir -p gpurun_out
python scripts/time_ops.py 1024 > gpurun_out/r02_time_ops_u.jsonl 2>> gpurun_out/r02_u.err
CUDDH_B200_AFFINE=0 python scripts/time_ops.py 1024 >> gpurun_out/r02_time_ops_u.jsonl 2>> gpurun_out/r02_u.err
cat gpurun_out/r02_time_ops_u.jsonl
ncu --set full --clock-control none -k regex:volume_action_pair -s 4 -c 1 -f -o gpurun_out/r02_prof_pair_final_s8 python scripts/time_high_order.py 1024 8 > gpurun_out/r02_ncu_u1.log 2>&1
CUDDH_B200_AFFINE=0 ncu --set full --clock-control none -k regex:volume_action_pair -s 4 -c 1 -f -o gpurun_out/r02_prof_pair_final_s8_stored python scripts/time_high_order.py 1024 8 > gpurun_out/r02_ncu_u2.log 2>&1
CUDDH_B200_AFFINE=0 ncu --set full --clock-control none -k regex:volume_action_pair -s 18 -c 1 -f -o gpurun_out/r02_prof_pair_final_m8 python scripts/time_high_order.py 1024 8 > gpurun_out/r02_ncu_u3.log 2>&1
tail -n 2 gpurun_out/r02_ncu_u1.log gpurun_out/r02_ncu_u2.log gpurun_out/r02_ncu_u3.log; tail -n 5 gpurun_out/r02_u.err; du -sh gpurun_out
