#!/bin/bash
# Run on the GPU box (under gpurun): plain bench first, then the ncu launch list of the same command and full captures of
# the dominant kernel (B200_PROFILING.md recipe). Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:volume_action_ws -s 4 -c 1 -f -o gpurun_out/prof_volume $CMD > gpurun_out/ncu_full.log 2>&1
# the stand-alone stiffness / weighted-mass instances of the same kernel
python scripts/prof_ops.py 1024 5 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:volume_action_ws -s 2 -c 2 -f -o gpurun_out/prof_ops python scripts/prof_ops.py 1024 5 > gpurun_out/ncu_ops.log 2>&1
tail -3 gpurun_out/plain.log
ls -la gpurun_out
