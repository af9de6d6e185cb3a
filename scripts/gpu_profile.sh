#!/bin/bash
# Run on the GPU box (under gpurun): plain bench first, then the ncu launch list and one full capture of the
# dominant kernel (B200_PROFILING.md recipe). Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:volume_action_kernel -s 6 -c 4 -o gpurun_out/prof_volume $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/plain.log
ls -la gpurun_out
