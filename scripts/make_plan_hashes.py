"""Regenerate tests/golden/plan_hashes.json: FNV-1a hashes of every array of the assembly plans (both layouts) for a set of meshes.
    python scripts/make_plan_hashes.py tests/golden/plan_hashes.json
Pins the plan builder bit for bit (the GPU kernels consume these arrays verbatim); host only."""
import sys, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cuddhelmholtz_b200 as cb
from conftest import load_mesh_file
out = {}
xy, el = load_mesh_file()
meshes = {"unstr": lambda: cb.Mesh2D.from_vertices(xy, el), "r37": lambda: cb.Mesh2D.uniform_rect(37,-1.0,1.0,29,0.0,2.0), "r10": lambda: cb.Mesh2D.uniform_rect(10,-1.0,1.0,10,-1.0,1.0), "r256": lambda: cb.Mesh2D.uniform_rect(256,-1.0,1.0,256,-1.0,1.0), "r3": lambda: cb.Mesh2D.uniform_rect(3,-1.0,1.0,2,-1.0,1.0)}
for tag, mk in meshes.items():
    for nb in ((2,3,4,5,8,9) if tag != "r256" else (4,5)):
        mesh = mk(); fem = cb.H1Space(mesh, cb.Basis(nb))
        for k in (0,1):
            st = fem.check_plan(k)
            assert st["mismatches"] == 0
            out["%s_%d_%d" % (tag, nb, k)] = st["hash"]
json.dump(out, open(sys.argv[1], "w"), indent=0)
print(len(out), "hashes")
