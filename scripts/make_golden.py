"""Generate tests/golden/*.npz from the UNMODIFIED reference's own host objects (oracle/_ref/ref_driver,
built by `make -C oracle ref` in the build container) and check the oracle restatement against them while
doing so. Small cases are stored as full arrays, larger ones as 64-bit FNV-1a hashes (oracle/rdmp.py).
Run in the build container only: /root/reference does not exist on the GPU box.

    python scripts/make_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.rdmp import fnv1a64, read_rdmp  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
GOLD = os.path.join(ROOT, "tests", "golden")
MESH = os.path.join(GOLD, "unstructured_square.txt")


def run(*args):
    with tempfile.NamedTemporaryFile(suffix=".bin") as f:
        subprocess.check_call([REF, *[str(a) for a in args], f.name])
        return read_rdmp(f.name)


def main():
    out = {}
    # 1-D tables: every (nb, nq) pair the operators instantiate + the dsteqr sizes
    pairs = [(2, 3), (3, 4), (3, 5), (3, 6), (4, 5), (4, 6), (4, 8), (5, 6), (5, 7), (5, 9), (6, 7), (6, 8), (6, 11), (7, 8), (7, 9),
             (7, 12), (8, 9), (8, 10), (8, 14), (9, 10), (9, 15)]
    for nb, nq in pairs:
        t = run("tables", nb, nq)
        for k, v in t.items():
            out["tables_%d_%d_%s" % (nb, nq, k)] = v
    np.savez_compressed(os.path.join(GOLD, "tables.npz"), **out)

    # index maps: full arrays for small meshes
    out = {}
    for spec, tag, nb in [("rect:2", "rect2", 3), ("rect:3", "rect3", 2), ("rect:10", "rect10", 4), ("rect:10", "rect10", 5),
                          ("rect:6", "rect6", 8), ("file:" + MESH, "unstr", 2), ("file:" + MESH, "unstr", 4),
                          ("file:" + MESH, "unstr", 5), ("file:" + MESH, "unstr", 9)]:
        h = run("h1", spec, nb)
        for k in ("edges", "boundary_edges", "ndof", "I", "xy", "fdof", "face_I", "face_proj", "min_h", "max_h", "n_edges"):
            out["h1_%s_%d_%s" % (tag, nb, k)] = h[k]
    np.savez_compressed(os.path.join(GOLD, "h1.npz"), **out)

    # index maps: hashes for larger meshes
    hashes = {}
    for spec, tag, nb in [("rect:64", "rect64", 5), ("rect:128", "rect128", 4), ("rect:256", "rect256", 5), ("rect:64", "rect64", 9),
                          ("rect:512", "rect512", 4), ("rect:1024", "rect1024", 5)]:  # 1024^2 = BASELINE config 2 (30 s, 6 GB here)
        h = run("h1", spec, nb)
        hashes["h1_%s_%d" % (tag, nb)] = dict(ndof=int(h["ndof"][0]), n_edges=int(h["n_edges"][0]), I=fnv1a64(h["I"]),
                                              xy=fnv1a64(h["xy"]), edges=fnv1a64(h["edges"]), face_I=fnv1a64(h["face_I"]),
                                              face_proj=fnv1a64(h["face_proj"]))
    out = {}
    for nx, nb in [(8, 4), (16, 4), (8, 8)]:
        e = run("ensemble", nx, nb, 16)
        for k, v in e.items():
            out["ens_%d_%d_%s" % (nx, nb, k)] = v
    np.savez_compressed(os.path.join(GOLD, "ensemble.npz"), **out)
    for nx, nb in [(32, 4), (64, 4), (32, 8)]:
        e = run("ensemble", nx, nb, 16)
        hashes["ens_%d_%d" % (nx, nb)] = dict(n_domains=int(e["n_domains"][0]), n_shared=len(e["cmap"]) // 4, cmap=fnv1a64(e["cmap"]),
                                              gI=fnv1a64(e["gI"]), sI=fnv1a64(e["sI"]), fI=fnv1a64(e["fI"]), pI=fnv1a64(e["pI"]))
    import json
    with open(os.path.join(GOLD, "hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
