"""DDH-GMRES solve (examples/DDH.cpp flow) sharded over the GPUs of one box; run under torch.distributed.run.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/ddh_multi.py [nx] [nb]
Prints one JSON line on rank 0: solve time (CUDA events, max over ranks), restarts, and the single-action time."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import cuddhelmholtz_b200 as cb
from cuddhelmholtz_b200.parallel import NeighbourDDH, ShardedDDH

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 4
omega_arg = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0     # 0: the examples' 2 pi nx / 10
only_actions = int(sys.argv[4]) if len(sys.argv) > 4 else 0        # > 0: time this many actions instead of a full solve
neighbour = int(sys.argv[5]) if len(sys.argv) > 5 else 0           # 1: NeighbourDDH (distributed vectors, send/recv of the slab-boundary traces)
world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
omega = omega_arg if omega_arg > 0 else 2 * np.pi * nx / 10
mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
fem = cb.H1Space(mesh, cb.Basis(nb))
xy = fem.physical_coordinates()
X, Y = xy[:, 0], xy[:, 1]
ha = np.where(X * X + Y * Y < 0.0625, 0.2, 1.0)
s = omega * omega
src = s / np.pi * np.exp(-s * ((X + 0.5) ** 2 + Y ** 2)) + s / np.pi * np.exp(-s * ((X - 0.5) ** 2 + (Y + 0.5) ** 2))
n = fem.size()
f = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
cb.MassMatrix(fem).action(torch.as_tensor(src, device="cuda"), f[:n])
D = cb.DDH(omega, ha, fem, nx, nx, 16)
A = NeighbourDDH(D, rank, world) if neighbour else ShardedDDH(D, rank, world)
m = D.size()
b = torch.empty(m, dtype=torch.float32, device="cuda")
A.rhs(f, b)
L = torch.zeros(m, dtype=torch.float32, device="cuda")
tmp = torch.empty_like(b)
A.action_tensors(b, tmp)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
A.action_tensors(b, tmp)
e1.record()
if only_actions > 0:
    class _O:  # timing-only mode
        num_iter = 0
        num_matvec = only_actions
        success = None
    out = _O()
    for _ in range(only_actions):
        A.action_tensors(b, tmp)
    L.copy_(tmp)
elif neighbour:
    res = A.solve(b, L, m=20, maxit=100, tol=1e-4)

    class _R:
        num_iter, num_matvec, success = res["num_iter"], res["num_matvec"], res["success"]
    out = _R()
else:
    out = cb.gmres(m, L, A, b, 20, 100, 1e-4)
e2.record()
torch.cuda.synchronize()
U = torch.empty(2 * n, dtype=torch.float64, device="cuda")
A.postprocess(L, f, U)
t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2)], dtype=torch.float64, device="cuda")
chk = torch.tensor([float(U.norm())], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    info = D.info()
    print(json.dumps({"n_gpus": world, "nx": nx, "n_basis": nb, "n_domains": info["n_domains"], "nt": info["nt"], "n_lambda": m,
                      "omega": omega, "action_ms": float(t[0]), "gmres_seconds": float(t[1]) / 1e3, "restarts": out.num_iter, "matvec": out.num_matvec,
                      "success": out.success, "exchange": ("send/recv %d B per action" % A.bytes_per_action) if neighbour else "allreduce %d B" % (4 * m), "u_norm": float(chk[0]), "action_fp32_tflops": D.flops() / (float(t[0]) * 1e-3) / 1e12}))
if world > 1:
    dist.destroy_process_group()
