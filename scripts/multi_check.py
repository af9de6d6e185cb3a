"""Multi-GPU checks of the library's distributed layer; run under torch.distributed.run on N >= 2 GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_check.py

(1) library communicator (NCCL through dlopen, id broadcast through torch.distributed); (2) path A: slab apply with the
exchange inside the library (pack fused into the face-mass launch, ncclSend/ncclRecv) == the torch.distributed exchange, bit for
bit, mirrored rows identical on both neighbours, distributed GMRES in the library == the torch implementation; (3) path B:
DDHDist action == NeighbourDDH action on the owned slots, distributed solve == single-GPU solve (restart count, solution).
Prints MULTI_OK on rank 0 when every rank passed (tests/test_multi_gpu.py looks for it)."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import cuddhelmholtz_b200 as cb
from cuddhelmholtz_b200.parallel import GpuSlabHelmholtz, NeighbourDDH

world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok, msg = True, ""
try:
    comm = cb.Comm(rank, world)
    t = torch.full((5,), float(rank + 1), dtype=torch.float64, device="cuda")
    comm.allreduce(t)
    torch.cuda.synchronize()
    assert float(t[0]) == world * (world + 1) / 2

    # ---------------- path A ----------------
    coef = lambda x, y: 1.0 + 0.5 * np.sin(np.pi * x) * np.cos(np.pi * y)
    nx, nb, omega = 128, 5, 20.0
    lib = GpuSlabHelmholtz(nx, nx, nb, omega, coef, rank, world, comm=comm)     # peer-memory exchange when CUDA IPC maps the neighbours
    os.environ["CUDDH_B200_PEER"] = "0"
    lib_nccl = GpuSlabHelmholtz(nx, nx, nb, omega, coef, rank, world, comm=comm)  # the ncclSend / ncclRecv path of the same library code
    del os.environ["CUDDH_B200_PEER"]
    assert not lib_nccl.lib.uses_peer_memory()
    ref = GpuSlabHelmholtz(nx, nx, nb, omega, coef, rank, world, comm=None)
    n = lib.ndof
    x = torch.as_tensor(np.random.default_rng(7 + rank).uniform(-1, 1, 2 * n), device="cuda")
    x2 = x.clone()
    lib.exchange(x)      # library stand-alone exchange: consistent input
    ref.exchange(x2)
    assert torch.equal(x, x2), "stand-alone exchange differs"
    y, y2, y3 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    for _ in range(5):   # repeated applies reuse the send / recv buffers (both parities of the peer path)
        lib.apply(x, y)
        ref.apply(x, y2)
        lib_nccl.apply(x, y3)
    torch.cuda.synchronize()
    assert torch.equal(y, y2), "library slab apply differs from the torch.distributed exchange: %g" % float((y - y2).abs().max())
    assert torch.equal(y3, y2), "library slab apply (NCCL path) differs from the torch.distributed exchange"
    # back-to-back applies without host synchronisation in between (epochs / parity buffers under load)
    ys = [torch.empty_like(x) for _ in range(8)]
    for k in range(8):
        lib.apply(x if k % 2 == 0 else y, ys[k])
    torch.cuda.synchronize()
    ref.apply(y, y2)
    assert torch.equal(ys[0], y) and torch.equal(ys[1], y2) and torch.equal(ys[7], ys[1]), "back-to-back applies differ"
    # mirrored rows: my top row == the bottom row of rank + 1 (bit for bit)
    rows = {}
    for which, peer in (("bottom", rank - 1), ("top", rank + 1)):
        if 0 <= peer < world:
            idx = torch.as_tensor(lib.interface_rows(which).astype(np.int64), device="cuda")
            rows[which] = torch.cat([y[idx], y[n + idx]])
    ops = []
    got = {}
    for which, peer in (("bottom", rank - 1), ("top", rank + 1)):
        if which in rows:
            got[which] = torch.empty_like(rows[which])
            ops.append(dist.P2POp(dist.isend, rows[which], peer))
            ops.append(dist.P2POp(dist.irecv, got[which], peer))
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    for which in rows:
        assert torch.equal(rows[which], got[which]), "mirrored interface row differs between neighbours"
    # distributed FP64 GMRES in the library against the torch implementation (CGS, one allreduce per step, both)
    b = torch.as_tensor(np.random.default_rng(99 + rank).uniform(-1, 1, 2 * n), device="cuda")
    lib.exchange(b)
    u1, u2 = torch.zeros_like(b), torch.zeros_like(b)
    o1 = lib.solve(b, u1, m=30, maxit=4, tol=1e-8)
    o2 = ref.solve(b.clone(), u2, m=30, maxit=4, tol=1e-8)
    assert o1.num_matvec == o2["num_matvec"], (o1.num_matvec, o2["num_matvec"])
    assert np.allclose(o1.res_norm, o2["res_norm"], rtol=1e-6), (o1.res_norm, o2["res_norm"])
    assert float((u1 - u2).norm() / u2.norm()) < 1e-6
    assert o1.allreduces <= o1.num_matvec + 2 * len(o1.res_norm) + o1.reorth + 2   # one allreduce per Arnoldi step (+ residual norms)

    # ---------------- path B ----------------
    nx, nb = 64, 4
    omega = 2 * np.pi * nx / 10
    mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
    fem = cb.H1Space(mesh, cb.Basis(nb))
    xy = fem.physical_coordinates()
    X, Y = xy[:, 0], xy[:, 1]
    ha = np.where(X * X + Y * Y < 0.0625, 0.2, 1.0)
    s = omega * omega
    src = s / np.pi * np.exp(-s * ((X + 0.5) ** 2 + Y ** 2)) + s / np.pi * np.exp(-s * ((X - 0.5) ** 2 + (Y + 0.5) ** 2))
    nd = fem.size()
    f = torch.zeros(2 * nd, dtype=torch.float64, device="cuda")
    cb.MassMatrix(fem).action(torch.as_tensor(src, device="cuda"), f[:nd])
    D = cb.DDH(omega, ha, fem, nx, nx, 16)
    A = cb.DDHDist(D, comm, rank, world)
    P = NeighbourDDH(D, rank, world)
    m = D.size()
    mask = A.mask()
    assert torch.equal(mask.float(), P.mask)
    b1, b2 = torch.empty(m, dtype=torch.float32, device="cuda"), torch.empty(m, dtype=torch.float32, device="cuda")
    A.rhs(f, b1)
    P.rhs(f, b2)
    assert torch.equal(b1, b2), "distributed rhs differs"
    lam = torch.as_tensor(np.random.default_rng(5).uniform(-1, 1, m).astype(np.float32), device="cuda") * P.mask
    y1, y2 = torch.empty_like(lam), torch.empty_like(lam)
    for _ in range(2):
        A.action(lam, y1)
        P.action_tensors(lam, y2)
    assert torch.equal(y1, y2), "distributed action differs"
    L = torch.zeros(m, dtype=torch.float32, device="cuda")
    o = A.solve(b1, L, m=20, maxit=100, tol=1e-4)
    U = torch.empty(2 * nd, dtype=torch.float64, device="cuda")
    A.postprocess(L, f, U)
    # the single-GPU solve of the same problem (every rank runs it redundantly)
    bs = torch.empty(m, dtype=torch.float32, device="cuda")
    D.rhs(f, bs)
    Ls = torch.zeros(m, dtype=torch.float32, device="cuda")
    os_ = cb.gmres(m, Ls, D, bs, 20, 100, 1e-4)
    Us = torch.empty(2 * nd, dtype=torch.float64, device="cuda")
    D.postprocess(Ls, f, Us)
    assert o.success and os_.success and abs(o.num_iter - os_.num_iter) <= 1, (o.num_iter, os_.num_iter)
    assert float((U - Us).norm() / Us.norm()) < 1e-3
    if rank == 0:
        print("multi_check: world %d  slab apply bitwise ok (peer memory: %s), slab gmres %d matvecs / %d allreduces, ddh solve %d restarts (single GPU %d), "
              "exchange %d B per action" % (world, lib.lib.uses_peer_memory(), o1.num_matvec, o1.allreduces, o.num_iter, os_.num_iter, A.info()["bytes_per_action"]))
except Exception:
    ok, msg = False, traceback.format_exc()
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if not ok:
    print("rank %d FAILED:\n%s" % (rank, msg), flush=True)
if rank == 0 and float(flag[0]) == 1.0:
    print("MULTI_OK", flush=True)
dist.destroy_process_group()
sys.exit(0 if float(flag[0]) == 1.0 else 1)
