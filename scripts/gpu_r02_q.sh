#!/bin/bash
# round 2, GPU call Q (one GPU): cluster-barrier mode compile-time again; per-thread chunk ring (TRG) against the TMA ring on the
# stored-metric fused / stand-alone instances; thread-pair kernels with U in registers
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_fused_variants_q.jsonl
: > $O
python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_q.err
CUDDH_B200_AFFINE_RING=5 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_q.err
CUDDH_B200_AFFINE=0 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_q.err
CUDDH_B200_AFFINE=0 CUDDH_B200_RING=-5 python scripts/fused_variant.py 1024 5 >> $O 2>> gpurun_out/r02_q.err
python scripts/fused_variant.py 1024 4 >> $O 2>> gpurun_out/r02_q.err
CUDDH_B200_AFFINE=0 python scripts/fused_variant.py 1024 4 >> $O 2>> gpurun_out/r02_q.err
CUDDH_B200_AFFINE=0 CUDDH_B200_RING4=-5 python scripts/fused_variant.py 1024 4 >> $O 2>> gpurun_out/r02_q.err
cat $O
T=gpurun_out/r02_time_ops_q.jsonl
: > $T
CUDDH_B200_AFFINE=0 python scripts/time_ops.py 1024 >> $T 2>> gpurun_out/r02_q.err
CUDDH_B200_AFFINE=0 CUDDH_B200_RING1=-5 CUDDH_B200_RING4=-5 python scripts/time_ops.py 1024 >> $T 2>> gpurun_out/r02_q.err
cat $T
H=gpurun_out/r02_high_order_q.jsonl
: > $H
for nb in 8 9 6 7; do
  python scripts/time_high_order.py 1024 $nb >> $H 2>> gpurun_out/r02_q.err
  CUDDH_B200_AFFINE=0 python scripts/time_high_order.py 1024 $nb >> $H 2>> gpurun_out/r02_q.err
done
cat $H
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -x -k "operator or steady or composite or volume or mass or stiff" ) > gpurun_out/r02_pytest_q.log 2>&1
tail -n 5 gpurun_out/r02_pytest_q.log
tail -n 5 gpurun_out/r02_q.err
