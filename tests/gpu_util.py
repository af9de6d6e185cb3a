"""helpers for the -m gpu tests"""
import numpy as np
import torch


def dev(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda")


def host(t):
    return t.detach().cpu().numpy()


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


class RawVec:
    """view a raw device address as a torch tensor (for operator callbacks that receive raw pointers)"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def as_tensor(ptr, n, dtype=torch.float64):
    return torch.as_tensor(RawVec(ptr, n, "<f8" if dtype == torch.float64 else "<f4"), device="cuda")


def warped_mesh(nx, ny=None):
    """smoothly perturbed structured grid on [-1,1]^2 as (xy, elems): every element gets its own (non-affine) metric, the
    boundary stays put. Vertex i + (nx+1) j, elements CCW from the lower-left corner (Mesh2D::uniform_rect's convention)."""
    ny = ny or nx
    X0, Y0 = np.meshgrid(np.linspace(-1.0, 1.0, nx + 1), np.linspace(-1.0, 1.0, ny + 1), indexing="xy")
    X = X0 + 0.06 * np.sin(np.pi * X0) * np.sin(np.pi * Y0)
    Y = Y0 + 0.05 * np.sin(2 * np.pi * X0) * np.sin(np.pi * Y0)
    xy = np.stack([X.ravel(), Y.ravel()], 1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (i + (nx + 1) * j).ravel()
    el = np.stack([v0, v0 + 1, v0 + nx + 2, v0 + nx + 1], 1).astype(np.int32)
    return xy, el


def write_mesh_file(path, xy, el):
    """the text format of tests/golden/unstructured_square.txt (ref_driver `file:` meshes); repr() round-trips doubles exactly"""
    with open(path, "w") as f:
        f.write("%d %d\n" % (len(xy), len(el)))
        f.write("".join("%r %r\n" % (float(a), float(b)) for a, b in xy))
        f.write("".join("%d %d %d %d\n" % tuple(int(v) for v in e) for e in el))


class max_ctas:
    """context manager: cap the CTAs of the persistent operator kernels (forces many patches per CTA on small meshes)"""

    def __init__(self, n):
        self.n = n

    def __enter__(self):
        import cuddhelmholtz_b200 as cb
        cb.set_option("max_ctas", self.n)

    def __exit__(self, *a):
        import cuddhelmholtz_b200 as cb
        cb.set_option("max_ctas", 0)
