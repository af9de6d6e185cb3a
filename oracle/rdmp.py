"""TEST INFRASTRUCTURE. Reader for the flat "RDMP" container written by oracle/ref_driver.cu,
plus the 64-bit FNV-1a hash SURVEY.md §8(c) pins the reference's host-side arrays with."""
import numpy as np

_DT = {0: np.int32, 1: np.float64, 2: np.float32}


def read_rdmp(path):
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    off = 0
    while off < len(buf):
        name = buf[off:off + 48].split(b"\0", 1)[0].decode()
        dt = int(np.frombuffer(buf, np.int32, 1, off + 48)[0])
        cnt = int(np.frombuffer(buf, np.int64, 1, off + 52)[0])
        off += 60
        dtype = np.dtype(_DT[dt])
        out[name] = np.frombuffer(buf, dtype, cnt, off).copy()
        off += cnt * dtype.itemsize
    return out


def fnv1a64(arr):
    """h = 0xcbf29ce484222325; for e: h ^= bits(e); h *= 0x100000001b3 (ints as uint32, doubles as raw 64-bit)."""
    a = np.ascontiguousarray(arr)
    if a.dtype == np.float64:
        bits = a.view(np.uint64).ravel()
    elif a.dtype in (np.int32, np.uint32):
        bits = a.view(np.uint32).astype(np.uint64).ravel()
    else:
        raise TypeError(a.dtype)
    h = 0xcbf29ce484222325
    M = 0x100000001b3
    mask = (1 << 64) - 1
    for e in bits.tolist():
        h = ((h ^ e) * M) & mask
    return "%016x" % h
