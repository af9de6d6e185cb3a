#!/bin/bash
# round 2, GPU call M (one GPU): fused affine kernel - mass rows from the ring (default) against straight from global memory
# (CUDDH_B200_AFFINE_RING=0): bitwise agreement, timings, and one full ncu capture of each variant
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_affine_direct.jsonl
: > $O
for nb in 5 4; do for nx in 256 1024; do
  python scripts/fused_variant.py $nx $nb >> $O 2>> gpurun_out/r02_m.err
  CUDDH_B200_AFFINE_RING=0 python scripts/fused_variant.py $nx $nb >> $O 2>> gpurun_out/r02_m.err
done; done
cat $O
( time CUDDH_B200_AFFINE_RING=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -x -k "helmholtz or Helmholtz or operator or fused or steady" ) > gpurun_out/r02_pytest_m.log 2>&1
tail -6 gpurun_out/r02_pytest_m.log
ncu --set full --clock-control none --import-source on -k regex:volume_action_ws -s 3 -c 1 -f -o gpurun_out/r02_prof_affine_ring python scripts/time_fused.py 5 > gpurun_out/r02_ncu_m1.log 2>&1
CUDDH_B200_AFFINE_RING=0 ncu --set full --clock-control none --import-source on -k regex:volume_action_ws -s 3 -c 1 -f -o gpurun_out/r02_prof_affine_direct python scripts/time_fused.py 5 > gpurun_out/r02_ncu_m2.log 2>&1
tail -2 gpurun_out/r02_ncu_m1.log gpurun_out/r02_ncu_m2.log; tail -5 gpurun_out/r02_m.err
