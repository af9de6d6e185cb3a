#!/bin/bash
# round 2, GPU call K (one GPU): block-32 variant of the n_basis 8 register-tiled DDH kernel - parity + timing
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ddh" ) > gpurun_out/r02_pytest_k.log 2>&1
tail -12 gpurun_out/r02_pytest_k.log
python - > gpurun_out/r02_ddh8b.json 2> gpurun_out/r02_ddh8b.err <<'PY'
import json, os, sys
sys.path.insert(0, os.getcwd())
import torch, cuddhelmholtz_b200 as cb, bench
print(json.dumps(bench.ddh_high_order(cb, torch, None)))
print(json.dumps(bench.ddh_high_order(cb, torch, None, nx=256)))
PY
cat gpurun_out/r02_ddh8b.json; tail -3 gpurun_out/r02_ddh8b.err
