#!/bin/bash
# round 2, GPU call T (one GPU): pair kernels (mass: ring + U in registers; affine stiffness: U in registers) timings; launch list of
# the bench command and the full capture of the default fused kernel
set -u
mkdir -p gpurun_out
H=gpurun_out/r02_high_order_t.jsonl
: > $H
for nb in 8 7; do
  python scripts/time_high_order.py 1024 $nb >> $H 2>> gpurun_out/r02_t.err
  CUDDH_B200_AFFINE=0 python scripts/time_high_order.py 1024 $nb >> $H 2>> gpurun_out/r02_t.err
done
cat $H
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -x -k "operator or steady" ) > gpurun_out/r02_pytest_t.log 2>&1
tail -n 5 gpurun_out/r02_pytest_t.log
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-ddh"
$CMD > gpurun_out/r02_plain_t.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_t.csv $CMD > gpurun_out/r02_ncu_list_t.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:volume_action_ws -s 4 -c 1 -f -o gpurun_out/r02_prof_fused_final $CMD > gpurun_out/r02_ncu_t1.log 2>&1
tail -n 2 gpurun_out/r02_plain_t.log | cut -c1-600; tail -n 2 gpurun_out/r02_ncu_t1.log; tail -n 5 gpurun_out/r02_t.err; du -sh gpurun_out
