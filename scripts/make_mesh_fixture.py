"""Regenerate tests/golden/unstructured_square.txt from the reference's mesh DATA fixture
(/root/reference/meshes/unstructured_square/{info,coordinates,elements}.txt): a single text file
"nv nel" / nv lines "x y" / nel lines "a b c d", the format oracle/ref_driver.cu and the tests read.
Run in the build container only (the reference tree does not exist on the GPU box)."""
import numpy as np, os
src = "/root/reference/meshes/unstructured_square"
nv, nel = (int(v) for v in open(src + "/info.txt").read().split())
xy = np.loadtxt(src + "/coordinates.txt").reshape(nv, 2)
el = np.loadtxt(src + "/elements.txt", dtype=np.int64).reshape(nel, 4)
out = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "unstructured_square.txt")
with open(out, "w") as f:
    f.write("%d %d\n" % (nv, nel))
    for x, y in xy:
        f.write("%s %s\n" % (repr(float(x)), repr(float(y))))
    for e in el:
        f.write("%d %d %d %d\n" % tuple(e))
print("wrote", out)
