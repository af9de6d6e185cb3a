import sys, os; sys.path.insert(0, "/root/repo")
import numpy as np, torch, cuddhelmholtz_b200 as cb
from cuddhelmholtz_b200.parallel import GpuSlabHelmholtz
import bench
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 5
slab = GpuSlabHelmholtz(1024, 1024, nb, 100.0, bench.coef, 0, 1)
x = torch.rand(2*slab.ndof, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
for _ in range(3): slab.apply(x, y)
torch.cuda.synchronize()
print(os.environ.get("CUDDH_B200_TPE_PX"), os.environ.get("CUDDH_B200_TPE_PY"), nb, slab.op.time_phases(x, y, 10))
