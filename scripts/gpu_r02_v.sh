#!/bin/bash
# round 2, GPU call V (one GPU): the whole GPU test-suite, smoke, the default bench line and the reference arm; full ncu captures of
# the final thread-pair kernels at n_basis 8 summarised ON THE BOX (reports are too large to travel back three at a time)
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_pytest_v.log 2>&1
tail -n 6 gpurun_out/r02_pytest_v.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_v.log 2>&1; tail -n 2 gpurun_out/r02_smoke_v.log
( time python bench.py ) > gpurun_out/r02_bench_v_n1.json 2> gpurun_out/r02_bench_v_n1.err
tail -c 600 gpurun_out/r02_bench_v_n1.json; tail -n 4 gpurun_out/r02_bench_v_n1.err
( time python bench.py --impl reference ) > gpurun_out/r02_bench_v_ref.json 2> gpurun_out/r02_bench_v_ref.err
tail -c 400 gpurun_out/r02_bench_v_ref.json
python scripts/time_ops.py 1024 > gpurun_out/r02_time_ops_v.jsonl 2>> gpurun_out/r02_v.err
CUDDH_B200_AFFINE=0 python scripts/time_ops.py 1024 >> gpurun_out/r02_time_ops_v.jsonl 2>> gpurun_out/r02_v.err
T=/tmp/ncu_pair; mkdir -p $T
ncu --set full --clock-control none -k regex:volume_action_pair -s 4 -c 1 -f -o $T/pair_s8_affine python scripts/time_high_order.py 1024 8 > gpurun_out/r02_ncu_v1.log 2>&1
CUDDH_B200_AFFINE=0 ncu --set full --clock-control none -k regex:volume_action_pair -s 4 -c 1 -f -o $T/pair_s8_stored python scripts/time_high_order.py 1024 8 > gpurun_out/r02_ncu_v2.log 2>&1
CUDDH_B200_AFFINE=0 ncu --set full --clock-control none -k regex:volume_action_pair -s 18 -c 1 -f -o $T/pair_m8 python scripts/time_high_order.py 1024 8 > gpurun_out/r02_ncu_v3.log 2>&1
python scripts/ncu_report.py "ncu --set full --clock-control none -k regex:volume_action_pair -s 4 -c 1 [CUDDH_B200_AFFINE=0] python scripts/time_high_order.py 1024 8 (stiffness: 5th launch; weighted mass: -s 18)   uniform_rect(1024), n_basis 8" $T/pair_s8_affine.ncu-rep $T/pair_s8_stored.ncu-rep $T/pair_m8.ncu-rep > gpurun_out/r02_pair_kernel_ncu.txt 2>> gpurun_out/r02_v.err
head -c 1500 gpurun_out/r02_pair_kernel_ncu.txt; tail -n 5 gpurun_out/r02_v.err; du -sh gpurun_out
