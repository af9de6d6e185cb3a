#!/bin/bash
# round 2, GPU call B (one GPU): full GPU test suite (no -x), per-warp ring A/B timing + its parity tests
set -u
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --durations=12 ) > gpurun_out/r02_pytest.log 2>&1
tail -25 gpurun_out/r02_pytest.log
for w in 0 1; do
  CUDDH_B200_WRING=$w timeout 300 python scripts/time_ops.py 1024 >> gpurun_out/r02_wring.jsonl 2>> gpurun_out/r02_wring.err
done
cat gpurun_out/r02_wring.jsonl
( CUDDH_B200_WRING=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "steady_state or helmholtz_composite or stiffness_and_mass or full_size" ) > gpurun_out/r02_pytest_wring.log 2>&1
tail -8 gpurun_out/r02_pytest_wring.log
