#!/bin/bash
# round 2, GPU call AE (one GPU): late experiments on the default fused kernel - patch shapes (run-time knob) and an alternative build
# with the `zero` kernel parameter held in a register
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_fused_variants_ae.jsonl
: > $O
python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_ae.err
CUDDH_B200_LIBPATH=$PWD/cuddhelmholtz_b200/lib/variants/v1/libcuddh_b200.so python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_ae.err
CUDDH_B200_TPE_PX=16 CUDDH_B200_TPE_PY=8 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_ae.err
CUDDH_B200_TPE_PX=32 CUDDH_B200_TPE_PY=4 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_ae.err
CUDDH_B200_TPE_PX=4 CUDDH_B200_TPE_PY=32 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_ae.err
python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_ae.err
CUDDH_B200_LIBPATH=$PWD/cuddhelmholtz_b200/lib/variants/v1/libcuddh_b200.so python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_ae.err
cat $O; tail -n 3 gpurun_out/r02_ae.err
