"""Patch-kernel timings (median of 10, CUDA events) of the stand-alone stiffness / weighted-mass actions at n_basis nb on
uniform_rect(nx) with whatever kernels the environment selects (CUDDH_B200_PAIR=0: lane-per-row kernel; CUDDH_B200_AFFINE=0:
stored-metric stiffness); one JSON line.   python scripts/time_high_order.py 1024 8"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cuddhelmholtz_b200 as cb, bench
nx, nb = int(sys.argv[1]), int(sys.argv[2])
r = bench.high_order_ops(cb, torch, nx, nb, 6459.0)
r["n_basis"] = nb
r["env"] = {k: v for k, v in os.environ.items() if k.startswith("CUDDH_B200")}
print(json.dumps(r))
