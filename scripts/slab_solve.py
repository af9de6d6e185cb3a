"""Distributed FP64 GMRES on the slab-partitioned Helmholtz operator (path A): one process per GPU, NCCL.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/slab_solve.py [nx] [nb] [m] [restarts]
Each rank holds a uniform_rect(nx, nx) slab (weak scaling). Unpreconditioned GMRES(m) does not converge on these problems in a
few restarts (neither does the reference's), so this reports the cost per Arnoldi step and the allreduce count."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import cuddhelmholtz_b200 as cb
from cuddhelmholtz_b200.parallel import GpuSlabHelmholtz
from bench import coef

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 5
m = int(sys.argv[3]) if len(sys.argv) > 3 else 30
restarts = int(sys.argv[4]) if len(sys.argv) > 4 else 3
world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cb.load()
slab = GpuSlabHelmholtz(nx, nx, nb, 100.0, coef, rank, world)
n = slab.ndof
b = torch.from_numpy(np.random.default_rng(3 + rank).uniform(-1, 1, 2 * n)).cuda()
slab.exchange(b)  # make b consistent on the mirrored rows (sum of both sides: any consistent right-hand side will do)
x = torch.zeros_like(b)
torch.cuda.synchronize()
t0 = time.perf_counter()
res = slab.solve(b, x, m=m, maxit=restarts + 1, tol=1e-12)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"n_gpus": world, "nx": nx, "n_basis": nb, "m": m, "ndof_per_gpu": n, "restarts": res["num_iter"] - 1, "matvecs": res["num_matvec"],
                      "allreduces": res["allreduces"], "seconds": dt, "ms_per_arnoldi_step": 1e3 * dt / max(res["num_matvec"], 1),
                      "res_norm_first_last": [res["res_norm"][0], res["res_norm"][-1]]}))
if world > 1:
    dist.destroy_process_group()
