#!/bin/bash
# round 2, GPU call I (one GPU): full GPU suite + smoke + default bench + reference arm
set -u
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r02_pytest_i.log 2>&1
tail -22 gpurun_out/r02_pytest_i.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err
tail -c 600 gpurun_out/r02_bench_i.json; tail -4 gpurun_out/r02_bench_i.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
cut -c1-400 gpurun_out/r02_bench_ref.json; tail -3 gpurun_out/r02_bench_ref.err
