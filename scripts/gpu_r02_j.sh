#!/bin/bash
# round 2, GPU call J (one GPU): register-tiled n_basis 8 DDH kernel - parity + timing
set -u
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -k "ddh" ) > gpurun_out/r02_pytest_j.log 2>&1
tail -15 gpurun_out/r02_pytest_j.log
python - > gpurun_out/r02_ddh8.json 2> gpurun_out/r02_ddh8.err <<'PY'
import json, os, sys
sys.path.insert(0, os.getcwd())
import torch, cuddhelmholtz_b200 as cb, bench
drv = os.path.join(os.getcwd(), "oracle", "_ref", "ref_driver")
print(json.dumps(bench.ddh_high_order(cb, torch, drv if os.path.exists(drv) else None)))
print(json.dumps(bench.ddh_high_order(cb, torch, None, nx=256)))
PY
cat gpurun_out/r02_ddh8.json; tail -3 gpurun_out/r02_ddh8.err
