#!/bin/bash
# round 2, GPU call AG (one GPU): full ncu capture of the fused two-phase thread-pair kernel (Helmholtz composite, n_basis 8), summarised on the box
set -u
mkdir -p gpurun_out
T=/tmp/ncu_ag; mkdir -p $T
ncu --set full --clock-control none -k regex:volume_action_pair -s 44 -c 1 -f -o $T/pair_fused8 python scripts/time_high_order.py 1024 8 > gpurun_out/r02_ncu_ag.log 2>&1
python scripts/ncu_report.py "ncu --set full --clock-control none -k regex:volume_action_pair -s 44 -c 1 python scripts/time_high_order.py 1024 8   (fused two-phase thread-pair kernel: S - w^2 M of one field, uniform_rect(1024), n_basis 8)" $T/pair_fused8.ncu-rep > gpurun_out/r02_pair_fused_ncu.txt 2>> gpurun_out/r02_ag.err
head -c 1200 gpurun_out/r02_pair_fused_ncu.txt; grep "warp-stall" gpurun_out/r02_pair_fused_ncu.txt | cut -c1-300; tail -n 3 gpurun_out/r02_ag.err
