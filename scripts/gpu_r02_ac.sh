#!/bin/bash
# round 2, GPU call AC (one GPU): final state - the whole GPU test-suite, smoke, the default bench line, the reference arm and the
# ncu launch list of the bench command
set -u
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_ac.log 2>&1
tail -n 8 gpurun_out/r02_pytest_ac.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_ac.log 2>&1; tail -n 2 gpurun_out/r02_smoke_ac.log
( time python bench.py ) > gpurun_out/r02_bench_ac_n1.json 2> gpurun_out/r02_bench_ac_n1.err
tail -c 400 gpurun_out/r02_bench_ac_n1.json; tail -n 4 gpurun_out/r02_bench_ac_n1.err
( time python bench.py --impl reference ) > gpurun_out/r02_bench_ac_ref.json 2> gpurun_out/r02_bench_ac_ref.err
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-ddh"
$CMD > gpurun_out/r02_plain_ac.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_ac.csv $CMD > gpurun_out/r02_ncu_list_ac.log 2>&1
du -sh gpurun_out
