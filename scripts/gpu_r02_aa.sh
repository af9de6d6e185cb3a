#!/bin/bash
# round 2, GPU call AA (one GPU): symmetric sqrt(w) tables in the affine stiffness phase (one table pair for forward and back-contraction)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_fused_variants_aa.jsonl
: > $O
python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_aa.err
CUDDH_B200_AFFINE=0 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_aa.err
python scripts/fused_variant.py 1024 4 >> $O 2>> gpurun_out/r02_aa.err
cat $O
python scripts/time_ops.py 1024 > gpurun_out/r02_time_ops_aa.jsonl 2>> gpurun_out/r02_aa.err
cat gpurun_out/r02_time_ops_aa.jsonl
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q ) > gpurun_out/r02_pytest_aa.log 2>&1
tail -n 6 gpurun_out/r02_pytest_aa.log; tail -n 3 gpurun_out/r02_aa.err
