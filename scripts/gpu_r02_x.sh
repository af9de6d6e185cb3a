#!/bin/bash
# round 2, GPU call X (one GPU): ncu of the stored-metric fused kernel (TMA ring and per-thread chunk ring), summarised on the box;
# n_basis 9 routing check
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -k "operator or steady" ) > gpurun_out/r02_pytest_x.log 2>&1
tail -n 4 gpurun_out/r02_pytest_x.log
CUDDH_B200_AFFINE=0 python scripts/time_high_order.py 1024 9 > gpurun_out/r02_high_order_x.jsonl 2>> gpurun_out/r02_x.err
cut -c1-600 gpurun_out/r02_high_order_x.jsonl
T=/tmp/ncu_x; mkdir -p $T
CUDDH_B200_AFFINE=0 ncu --set full --clock-control none -k regex:volume_action_ws -s 3 -c 1 -f -o $T/fused_stored_tma python scripts/time_fused.py 5 > gpurun_out/r02_ncu_x1.log 2>&1
CUDDH_B200_AFFINE=0 CUDDH_B200_RING=-5 ncu --set full --clock-control none -k regex:volume_action_ws -s 3 -c 1 -f -o $T/fused_stored_trg python scripts/time_fused.py 5 > gpurun_out/r02_ncu_x2.log 2>&1
python scripts/ncu_report.py "CUDDH_B200_AFFINE=0 [CUDDH_B200_RING=-5] ncu --set full --clock-control none -k regex:volume_action_ws -s 3 -c 1 python scripts/time_fused.py 5   (stored-metric fused Helmholtz kernel, uniform_rect(1024), n_basis 5: TMA ring / per-thread chunk ring)" $T/fused_stored_tma.ncu-rep $T/fused_stored_trg.ncu-rep > gpurun_out/r02_fused_stored_ncu.txt 2>> gpurun_out/r02_x.err
python scripts/ncu_src.py $T/fused_stored_trg.ncu-rep volume_action_ws 60 > gpurun_out/r02_fused_stored_trg_src.txt 2>> gpurun_out/r02_x.err
python scripts/ncu_src.py $T/fused_stored_tma.ncu-rep volume_action_ws 60 > gpurun_out/r02_fused_stored_tma_src.txt 2>> gpurun_out/r02_x.err
grep -n "warp-stall\|^==\|gpu__time" gpurun_out/r02_fused_stored_ncu.txt | cut -c1-300; tail -n 3 gpurun_out/r02_x.err; du -sh gpurun_out
