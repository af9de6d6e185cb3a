#!/bin/bash
# round 2, GPU call H (one GPU): affine stiffness path - parity, timing A/B, FGMRES test, bench
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02_affine_ab.jsonl
for a in 1 0 1 0; do
  echo "{\"affine\": $a}" >> gpurun_out/r02_affine_ab.jsonl
  CUDDH_B200_AFFINE=$a timeout 300 python scripts/time_ops.py 1024 >> gpurun_out/r02_affine_ab.jsonl 2>> gpurun_out/r02_affine_ab.err
done
cat gpurun_out/r02_affine_ab.jsonl
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -k "affine or steady_state or helmholtz_composite or stiffness_and_mass or full_size or fgmres or operator_actions or helmholtz_gmres" ) > gpurun_out/r02_pytest_h.log 2>&1
tail -15 gpurun_out/r02_pytest_h.log
( time timeout 900 python bench.py --steps 20 --warmup 5 --no-ddh ) > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err
tail -c 1200 gpurun_out/r02_bench_h.json; tail -3 gpurun_out/r02_bench_h.err
