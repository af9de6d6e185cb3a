#!/bin/bash
# round 2, GPU call E (one GPU): new tests, then the profiling recipe (B200_PROFILING.md): plain run, ncu launch list of the same
# command, full captures of the fused Helmholtz kernel and of the DDH kernel
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -k "fgmres or ddh_matches_reference or helmholtz_gmres_matches" ) > gpurun_out/r02_pytest_e.log 2>&1
tail -15 gpurun_out/r02_pytest_e.log
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-ddh"
$CMD > gpurun_out/r02_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:volume_action_ws -s 4 -c 1 -f -o gpurun_out/r02_prof_fused $CMD > gpurun_out/r02_ncu_fused.log 2>&1
python scripts/prof_ddh.py 128 > gpurun_out/r02_prof_ddh_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ddh_kernel_reg4 -s 1 -c 1 -f -o gpurun_out/r02_prof_ddh python scripts/prof_ddh.py 128 > gpurun_out/r02_ncu_ddh.log 2>&1
python scripts/prof_ddh.py 512 > /dev/null 2>&1 &&
ncu --set full --clock-control none -k regex:ddh_kernel_reg4 -s 1 -c 1 -f -o gpurun_out/r02_prof_ddh512 python scripts/prof_ddh.py 512 > gpurun_out/r02_ncu_ddh512.log 2>&1
ls -la gpurun_out | tail -12
