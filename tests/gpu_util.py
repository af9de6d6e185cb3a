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
