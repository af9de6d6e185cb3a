#!/bin/bash
# round 2, GPU call F (one GPU): lean ring producer - timing + parity, FGMRES test
set -u
mkdir -p gpurun_out
timeout 300 python scripts/time_ops.py 1024 > gpurun_out/r02_leanprod.jsonl 2> gpurun_out/r02_leanprod.err
cat gpurun_out/r02_leanprod.jsonl
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -k "steady_state or helmholtz_composite or stiffness_and_mass or full_size or fgmres or operator_actions" ) > gpurun_out/r02_pytest_f.log 2>&1
tail -12 gpurun_out/r02_pytest_f.log
