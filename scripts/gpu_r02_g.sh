#!/bin/bash
# round 2, GPU call G (one GPU): A/B of the ring-producer variants (separate library builds selected by CUDDH_B200_LIB)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02_producer_ab.jsonl
for v in base early late base early late; do
  echo "{\"variant\": \"$v\"}" >> gpurun_out/r02_producer_ab.jsonl
  CUDDH_B200_LIB=$PWD/cuddhelmholtz_b200/lib/libcuddh_b200_$v.so timeout 300 python scripts/time_ops.py 1024 >> gpurun_out/r02_producer_ab.jsonl 2>> gpurun_out/r02_producer_ab.err
done
cat gpurun_out/r02_producer_ab.jsonl
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fgmres" ) > gpurun_out/r02_pytest_g.log 2>&1
tail -12 gpurun_out/r02_pytest_g.log
