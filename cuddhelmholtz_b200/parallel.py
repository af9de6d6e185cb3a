"""Multi-GPU layer for path A (operator apply): one process per GPU, `torch.distributed` for the plumbing.

The reference is single-GPU (no NCCL/MPI anywhere); this is the partitioning SURVEY §8(e) prescribes:
a uniform_rect(nx, ny_total) mesh is cut into `world` slabs of element rows, one per rank. A rank holds every DOF its
elements touch, so the node row on a slab interface is held (with identical values) by both neighbours. One operator
apply is: local matrix-free apply (partial sums on the interface rows) -> exchange of the interface rows with the
two neighbours (NCCL send/recv over NVLink, 8*(nx*(nb-1)+1) bytes per vector per neighbour) -> add. The sum
`own + received` is the same two-term FP addition on both sides (commutative), so both copies stay bitwise
identical and the result does not depend on the rank count.

The exchange is the only collective on this path. The logic here is backend-agnostic (the local operator is any
object with apply / restrict / prolong), which is how tests/test_parallel_cpu.py runs it on 2 gloo ranks with the
oracle as local operator.
"""
import numpy as np
import torch
import torch.distributed as dist


def slab_geometry(rank, world, ny_local, ay=-1.0, height=2.0):
    """y-range of rank's slab: slabs are stacked upwards, each `height` tall (weak scaling: fixed work per rank)."""
    return ay + height * rank, ay + height * (rank + 1)


def classify_boundary_edges(edges, boundary_edges, nx, ny):
    """Split the boundary edges of a uniform_rect(nx, ny) mesh (vertex id = i + (nx+1) j) into bottom / top / sides,
    each in edge-id order (= increasing x for bottom / top)."""
    n0 = edges[boundary_edges, 0].astype(np.int64)
    n1 = edges[boundary_edges, 1].astype(np.int64)
    row = nx + 1
    bottom = boundary_edges[(n0 < row) & (n1 < row)]
    top = boundary_edges[(n0 >= row * ny) & (n1 >= row * ny)]
    sides = boundary_edges[((n0 % row == 0) & (n1 % row == 0)) | ((n0 % row == nx) & (n1 % row == nx))]
    assert len(bottom) == nx and len(top) == nx and len(sides) == 2 * ny
    return bottom, top, sides


class SlabExchange:
    """Adds the neighbours' partial sums on the slab-interface rows of one or more stacked vectors.

    local: object with
        n_vec_rows()                      -> number of DOFs in one interface row
        restrict(which, y, buf)           -> buf <- y[interface row]     which in {"bottom", "top"}
        prolong(which, buf, y)            -> y[interface row] += buf
    Vectors may be stacks ([u; v]): `offsets` lists the start of each stacked block.
    """

    def __init__(self, local, rank, world, offsets=(0,), dtype=torch.float64, device="cpu", group=None):
        self.local, self.rank, self.world, self.offsets, self.group = local, rank, world, tuple(offsets), group
        m = local.n_vec_rows() * len(self.offsets)
        mk = lambda: torch.empty(m, dtype=dtype, device=device)
        self.has_bottom, self.has_top = rank > 0, rank < world - 1
        self.send_b, self.recv_b = (mk(), mk()) if self.has_bottom else (None, None)
        self.send_t, self.recv_t = (mk(), mk()) if self.has_top else (None, None)
        self.bytes_per_apply = 8 * m * (int(self.has_bottom) + int(self.has_top))

    def __call__(self, y):
        if self.world == 1:
            return
        r = self.local.n_vec_rows()
        ops = []
        for which, sb, rb, peer in (("bottom", self.send_b, self.recv_b, self.rank - 1), ("top", self.send_t, self.recv_t, self.rank + 1)):
            if sb is None:
                continue
            for k, off in enumerate(self.offsets):
                self.local.restrict(which, y, off, sb[k * r:(k + 1) * r])
            ops.append(dist.P2POp(dist.isend, sb, peer, group=self.group))
            ops.append(dist.P2POp(dist.irecv, rb, peer, group=self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for which, rb in (("bottom", self.recv_b), ("top", self.recv_t)):
            if rb is None:
                continue
            for k, off in enumerate(self.offsets):
                self.local.prolong(which, rb[k * r:(k + 1) * r], y, off)


class GpuSlabHelmholtz:
    """Helmholtz composite (examples/Helmholtz.hpp) on this rank's slab, on the GPU, plus the interface exchange."""

    def __init__(self, nx, ny_local, nb, omega, coef, rank, world, group=None, comm=None):
        """comm: a cuddhelmholtz_b200.Comm -> the interface exchange runs inside the library (HelmholtzSlab: pack fused into the
        face-mass launch, ncclSend/ncclRecv, 4 launches per apply); None -> the torch.distributed exchange below (any backend)."""
        import cuddhelmholtz_b200 as cb
        self.cb = cb
        self.rank, self.world = rank, world
        ay, by = slab_geometry(rank, world, ny_local)
        self.mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, ny_local, ay, by)
        self.fem = cb.H1Space(self.mesh, cb.Basis(nb))
        self.ndof = self.fem.size()
        edges, bnd = self.mesh.edges(), self.mesh.boundary_edges()
        bottom, top, sides = classify_boundary_edges(edges, bnd, nx, ny_local)
        phys = [sides] + ([bottom] if rank == 0 else []) + ([top] if rank == world - 1 else [])
        self.fs_phys = cb.FaceSpace(self.fem, np.sort(np.concatenate(phys)))   # the physical boundary only
        self.fs = {"bottom": cb.FaceSpace(self.fem, bottom), "top": cb.FaceSpace(self.fem, top)}
        xy = self.fem.physical_coordinates()
        c = coef(xy[:, 0], xy[:, 1])
        dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
        self._a2, self._af = dev(c * c), dev(c[self.fs_phys.global_indices()])
        self.op = cb.Helmholtz(omega, self._a2, self._af, self.fem, self.fs_phys)
        self.lib, self.comm, self._lanes = None, comm, {}
        if comm is not None and world > 1:
            self.lib = self._make_lib()
            self.exchange = self.lib.exchange
        else:
            self.exchange = SlabExchange(self, rank, world, offsets=(0, self.ndof), device="cuda", group=group)

    def _make_lib(self):
        none = np.zeros(0, np.int32)
        return self.cb.HelmholtzSlab(self.op, self.comm, self.rank, self.world, self.fem, self.fs_phys,
                                     self.fs["bottom"].global_indices() if self.rank > 0 else none,
                                     self.fs["top"].global_indices() if self.rank < self.world - 1 else none)

    def lane(self, k):
        """a further exchange handle on the same operator for requests kept in flight on stream k (the exchanges of one handle
        must be stream-ordered); collective: every rank must create its lanes in the same order"""
        if self.lib is None:
            return self
        if k not in self._lanes:
            self._lanes[k] = self.lib if k == 0 else self._make_lib()
        return self._lanes[k]

    def exchange_info(self):
        if self.lib is not None:
            how = ("NVLink stores into the neighbours' receive buffers (CUDA IPC) + epoch flags" if self.lib.uses_peer_memory()
                   else "grouped ncclSend/ncclRecv")
            return {"bytes_per_apply": self.lib.bytes_per_apply(), "peer_memory": self.lib.uses_peer_memory(),
                    "path": "library: pack fused into the face-mass launch, %s, add kernel" % how}
        return {"bytes_per_apply": getattr(self.exchange, "bytes_per_apply", 0), "path": "torch.distributed send/recv + restrict / prolong kernels"}

    def n_vec_rows(self):
        return self.fs["bottom"].size()

    def restrict(self, which, y, off, buf):
        self.fs[which].restrict(y[off:off + self.ndof], buf)

    def prolong(self, which, buf, y, off):
        self.fs[which].prolong(buf, y[off:off + self.ndof])

    def interface_rows(self, which):
        """slab-local DOF indices of the bottom / top interface node row"""
        return self.fs[which].global_indices()

    def apply(self, x, y):
        """y = A x on [u; v] with x consistent on the interface rows; y comes out consistent as well."""
        if self.lib is not None:
            self.lib.apply(x, y)
            return
        self.op.action(x, y)
        self.exchange(y)

    def solve(self, b, x, m=20, maxit=1000, tol=1e-6, group=None):
        """distributed GMRES(m) on this slab partition (one allreduce per Arnoldi step), see slab_gmres"""
        if self.lib is not None:
            return self.lib.solve(b, x, m=m, maxit=maxit, tol=tol, orth=1)
        mask = owned_mask(self, self.rank, self.world, (0, self.ndof), 2 * self.ndof, device="cuda")
        y = torch.empty_like(b)

        def A(v):
            v = v.contiguous()
            self.apply(v, y)
            return y.clone()

        return slab_gmres(A, x, b, mask, m, maxit, tol, group=group, world=self.world)


def subdomain_range(n_domains, rank, world):
    """contiguous block of subdomains of `rank` (subdomain ids are row-major over the subdomain grid -> row slabs)"""
    base, rem = divmod(n_domains, world)
    b = rank * base + min(rank, rem)
    return b, b + base + (1 if rank < rem else 0)


class ShardedDDH:
    """Path B across GPUs: the subdomain loop of DDH::{rhs, action, postprocess} is split into contiguous
    subdomain ranges, one per rank; vectors (lambda, f, u) are replicated. A rank's kernel writes only the lambda slots
    (resp. DOFs) its subdomains own and leaves zeros elsewhere; since every slot has exactly one writer, a sum-allreduce
    (NCCL over NVLink) reproduces T(x) exactly — bit for bit the single-GPU vector — and action = x - allreduce(T_r(x)).
    The Krylov vectors are replicated, so GMRES runs redundantly and identically on every rank with no further
    communication (its vector work is negligible next to the WaveHoltz solves: 1e5 flops per lambda entry per action).
    Objects of this class plug into cuddhelmholtz_b200.gmres as an operator over raw device pointers."""

    def __init__(self, ddh, rank, world, group=None):
        self.ddh, self.rank, self.world, self.group = ddh, rank, world, group
        self.n = ddh.size()
        self.range = subdomain_range(ddh.info()["n_domains"], rank, world)

    def size(self):
        return self.n

    def _sum(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def apply_T(self, x, t):
        self.ddh.apply_T_range(x, t, *self.range)
        self._sum(t)

    def action_tensors(self, x, y):
        self.apply_T(x, y)
        y.neg_().add_(x)  # y = x - T(x)   (source/DDH.cpp:638)

    def action(self, xp, yp):
        """operator interface used by gmres(): raw device addresses of float vectors of length size()"""
        x = _as_tensor(xp, self.n, torch.float32)
        y = _as_tensor(yp, self.n, torch.float32)
        self.action_tensors(x, y)

    def rhs(self, f, b):
        self.ddh.rhs_range(f, b, *self.range)
        self._sum(b)

    def postprocess(self, lam, f, u):
        self.ddh.postprocess_range(lam, f, u, *self.range)
        self._sum(u)


class NeighbourDDH:
    """Path B with DISTRIBUTED Krylov vectors and a neighbour-only trace exchange (north_star): subdomains
    are dealt to the ranks in contiguous row slabs as in ShardedDDH, but a lambda slot now lives on ONE rank - the rank of
    the subdomain that reads it (slots nobody reads: the writer's rank) - and after the local WaveHoltz solves only the
    traces written across a slab boundary travel: one NCCL send/recv pair per neighbour (~n_dx * 13 * 4 floats per
    neighbour for n_basis 4) instead of an allreduce of the whole vector. Inner products are masked to the owned slots and
    summed with one small allreduce per Arnoldi step (`solve` = slab_gmres in FP32 with FP64 reductions).

    The reader / writer of every slot come straight from the library's B table (source/DDH.cpp:425-440, including its
    cross-point overwrites): slot B(j,0,dom) is read and slot B(j,1,dom) written by subdomain dom."""

    def __init__(self, ddh, rank, world, group=None, device="cuda"):
        self.ddh, self.rank, self.world, self.group = ddh, rank, world, group
        info = ddh.info()
        nd, self.n = info["n_domains"], ddh.size()
        nl = self.n // 2
        self.range = subdomain_range(nd, rank, world)
        B = ddh.array("B").reshape(nd, 2, -1)
        dom_rank = np.empty(nd, np.int32)
        for r in range(world):
            a, b = subdomain_range(nd, r, world)
            dom_rank[a:b] = r
        who = np.broadcast_to(dom_rank[:, None], B[:, 0, :].shape)
        reader, writer = np.full(nl, -1, np.int32), np.full(nl, -1, np.int32)
        for c, arr in ((0, reader), (1, writer)):
            idx = B[:, c, :]
            ok = idx >= 0
            arr[idx[ok]] = who[ok]
        self.owner = np.where(reader >= 0, reader, np.where(writer >= 0, writer, 0)).astype(np.int32)
        both = lambda k: torch.as_tensor(np.concatenate([k, k + nl]), dtype=torch.long, device=device)  # lambda and mu halves
        mine = self.owner == rank
        self.mask = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.mask[both(np.nonzero(mine)[0])] = 1.0
        self.send_idx, self.recv_idx = {}, {}
        for q in range(world):
            if q == rank:
                continue
            snd = np.nonzero((writer == rank) & (self.owner == q))[0]
            rcv = np.nonzero(mine & (writer == q))[0]
            if len(snd):
                self.send_idx[q] = both(snd)
            if len(rcv):
                self.recv_idx[q] = both(rcv)
        self.bytes_per_action = 4 * sum(int(v.numel()) for v in self.send_idx.values())

    def size(self):
        return self.n

    def exchange(self, t):
        """t holds what this rank's subdomains wrote: ship the slots owned elsewhere, take delivery of the owned ones"""
        if self.world == 1:
            return
        ops, bufs = [], {}
        for q, idx in self.send_idx.items():
            sb = t[idx].contiguous()
            bufs[("s", q)] = sb
            ops.append(dist.P2POp(dist.isend, sb, q, group=self.group))
        for q, idx in self.recv_idx.items():
            rb = torch.empty(idx.numel(), dtype=t.dtype, device=t.device)
            bufs[("r", q)] = rb
            ops.append(dist.P2POp(dist.irecv, rb, q, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for q, idx in self.recv_idx.items():
            t[idx] = bufs[("r", q)]

    def apply_T(self, x, t):
        self.ddh.apply_T_range(x, t, *self.range)
        self.exchange(t)

    def action_tensors(self, x, y):
        self.apply_T(x, y)
        y.neg_().add_(x).mul_(self.mask)  # y = x - T(x) on the owned slots, 0 elsewhere

    def rhs(self, f, b):
        self.ddh.rhs_range(f, b, *self.range)
        self.exchange(b)
        b.mul_(self.mask)

    def postprocess(self, lam, f, u):
        """u from (f, lambda): every rank adds its subdomains' partition-of-unity contributions (one allreduce, once)"""
        self.ddh.postprocess_range(lam, f, u, *self.range)
        if self.world > 1:
            dist.all_reduce(u, op=dist.ReduceOp.SUM, group=self.group)

    def solve(self, b, x, m=20, maxit=100, tol=1e-4):
        y = torch.empty_like(b)

        def A(v):
            self.action_tensors(v.contiguous(), y)
            return y.clone()

        return slab_gmres(A, x, b, self.mask, m, maxit, tol, group=self.group, world=self.world)


class _RawVec:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _as_tensor(ptr, n, dtype):
    return torch.as_tensor(_RawVec(ptr, n, "<f4" if dtype == torch.float32 else "<f8"), device="cuda")


# ---- distributed FP64 GMRES for the slab-partitioned operator (path A) ----------------------------------
def owned_mask(local, rank, world, offsets, n, device="cpu"):
    """1.0 on the DOFs this rank owns, 0.0 on the copies it only mirrors: a slab-interface node row is held by both
    neighbours, the lower rank owns it, so rank > 0 drops its bottom row from every inner product.
    `local.interface_rows(which)` gives the slab-local DOF indices of that row."""
    w = torch.ones(n, dtype=torch.float64, device=device)
    if world > 1 and rank > 0:
        idx = torch.as_tensor(np.asarray(local.interface_rows("bottom")), dtype=torch.long, device=device)
        for off in offsets:
            w[off + idx] = 0.0
    return w


def slab_gmres(apply, x, b, mask, m, maxit, tol=1e-6, group=None, world=1):
    """Restarted GMRES(m) with the reference's control flow (source/gmres.cpp:91-235: restart counter from 1, inner break
    on |eta_{k+1}| < tol*||b||, true residual after every restart) on vectors that are partitioned into slabs with
    mirrored interface rows. `apply(v) -> A v` must return a vector that is consistent on the mirrored rows (operator +
    SlabExchange). Inner products count every DOF once (`mask`) and cost ONE allreduce per Arnoldi step: classical
    Gram-Schmidt, [V_k^T w ; w.w] reduced together, ||w - V_k h||^2 = w.w - h.h (recomputed explicitly, one more
    allreduce, only when that difference cancels badly). Every rank runs the same small Hessenberg / Givens update.
    Returns dict(success, num_iter, num_matvec, res_norm, allreduces); x is updated in place."""
    n = x.numel()
    out = dict(success=False, num_iter=0, num_matvec=0, res_norm=[], allreduces=0)

    def reduce_(t):
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            out["allreduces"] += 1
        return t

    def dot(a, c):  # local partial sums in FP64 whatever the vector type (the FP32 path of DDH included)
        return float(reduce_(torch.dot((mask * a).double(), c.double()).reshape(1))[0])

    bnrm = np.sqrt(dot(b, b))
    V = torch.zeros(m + 1, n, dtype=x.dtype, device=x.device)
    H = np.zeros((m + 1, m))
    cs, sn = np.zeros(m), np.zeros(m)
    r = b - apply(x)
    out["num_matvec"] += 1
    r_nrm = np.sqrt(dot(r, r))
    out["res_norm"].append(r_nrm)
    if r_nrm < tol * bnrm:
        out["success"] = True
        return out
    it = 1
    while it < maxit:
        V[0] = r / r_nrm
        eta = np.zeros(m + 1)
        eta[0] = r_nrm
        k1 = 0
        for k in range(m):
            k1 = k + 1
            w = apply(V[k])
            out["num_matvec"] += 1
            mw = mask * w
            red = torch.empty(k1 + 1, dtype=torch.float64, device=x.device)  # partial sums travel in FP64 (FP32 path too)
            mw64 = mw.double()
            red[:k1] = V[:k1].double() @ mw64 if x.dtype != torch.float64 else V[:k1] @ mw64
            red[k1] = torch.dot(mw64, w.double())
            red = reduce_(red).cpu().numpy()
            h, ww = red[:k1], red[k1]
            w = w - torch.as_tensor(h, dtype=x.dtype, device=x.device) @ V[:k1]
            nrm2 = ww - float(np.dot(h, h))
            if nrm2 < 1e-3 * ww:  # cancellation: take the norm of the orthogonalised vector itself
                nrm2 = dot(w, w)
            hk1 = np.sqrt(max(nrm2, 0.0))
            H[:k1, k] = h
            H[k1, k] = hk1
            if hk1 == 0.0:
                break
            V[k1] = w / hk1
            for j in range(k):  # Givens rotations, source/gmres.cpp:7-23
                t = cs[j] * H[j, k] + sn[j] * H[j + 1, k]
                H[j + 1, k] = -sn[j] * H[j, k] + cs[j] * H[j + 1, k]
                H[j, k] = t
            d = np.hypot(H[k, k], H[k1, k])
            cs[k], sn[k] = H[k, k] / d, H[k1, k] / d
            H[k, k] = cs[k] * H[k, k] + sn[k] * H[k1, k]
            H[k1, k] = 0.0
            eta[k1] = -sn[k] * eta[k]
            eta[k] = cs[k] * eta[k]
            if abs(eta[k1]) < tol * bnrm:
                break
        yk = np.linalg.solve(np.triu(H[:k1, :k1]), eta[:k1]) if k1 > 0 else np.zeros(0)
        x += torch.as_tensor(yk, dtype=x.dtype, device=x.device) @ V[:k1]
        r = b - apply(x)
        out["num_matvec"] += 1
        r_nrm = np.sqrt(dot(r, r))
        out["res_norm"].append(r_nrm)
        if r_nrm < tol * bnrm:
            out["success"] = True
            break
        it += 1
    out["num_iter"] = it
    return out
