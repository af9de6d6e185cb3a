"""Kernel timings (median of 30 launches, CUDA events) of the stand-alone stiffness / weighted-mass kernels and of the fused Helmholtz
kernel at uniform_rect(nx), for n_basis 4 and 5; one JSON line. Environment knobs (CUDDH_B200_WRING, ...) are read by the library."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cuddhelmholtz_b200 as cb
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
peak = 6459.0
out = {"nx": nx, "wring": os.environ.get("CUDDH_B200_WRING")}
mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
for nb in (5, 4):
    fem = cb.H1Space(mesh, cb.Basis(nb))
    fs = cb.FaceSpace(fem, mesh.boundary_edges())
    n = fem.size()
    x = torch.rand(2 * n, dtype=torch.float64, device="cuda") - 0.5
    y = torch.empty_like(x)
    a = torch.rand(n, dtype=torch.float64, device="cuda") + 0.5
    af = torch.ones(fs.size(), dtype=torch.float64, device="cuda")
    r = {}
    for name, op, xx, yy in (("S", cb.StiffnessMatrix(fem), x[:n], y[:n]), ("M", cb.MassMatrix(a, fem), x[:n], y[:n]),
                             ("H", cb.Helmholtz(100.0, a, af, fem, fs), x, y)):
        op.action(xx, yy)
        p, s = op.time_phases(xx, yy, 30)
        r[name] = {"kernel_ms": round(p, 4), "rest_ms": round(s, 4), "hbm_frac": round(op.algorithmic_bytes() / (p * 1e-3) / 1e9 / peak, 4)}
        del op
    out["nb%d" % nb] = r
    del fem, fs, x, y, a, af
    torch.cuda.empty_cache()
print(json.dumps(out))
