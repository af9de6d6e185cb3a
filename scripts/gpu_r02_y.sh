#!/bin/bash
# round 2, GPU call Y (one GPU): element slot as an opaque register - timings of the fused variants and of the stand-alone operators
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_fused_variants_y.jsonl
: > $O
python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_y.err
CUDDH_B200_AFFINE=0 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_y.err
CUDDH_B200_AFFINE=0 CUDDH_B200_RING=-5 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_y.err
python scripts/fused_variant.py 1024 4 >> $O 2>> gpurun_out/r02_y.err
CUDDH_B200_AFFINE=0 python scripts/fused_variant.py 1024 4 >> $O 2>> gpurun_out/r02_y.err
cat $O
python scripts/time_ops.py 1024 > gpurun_out/r02_time_ops_y.jsonl 2>> gpurun_out/r02_y.err
CUDDH_B200_AFFINE=0 python scripts/time_ops.py 1024 >> gpurun_out/r02_time_ops_y.jsonl 2>> gpurun_out/r02_y.err
CUDDH_B200_AFFINE=0 CUDDH_B200_RING1=-5 CUDDH_B200_RING4=-5 python scripts/time_ops.py 1024 >> gpurun_out/r02_time_ops_y.jsonl 2>> gpurun_out/r02_y.err
cat gpurun_out/r02_time_ops_y.jsonl; tail -n 3 gpurun_out/r02_y.err
