"""BASELINE configs[2] size on one GPU: uniform_rect(2048), n_basis 4, omega 100, DDH block 16 (262 144 subdomains,
nt = 10 295) — setup, rhs, ONE action (a full GMRES(30) solve is ~600 actions), plus the FP64 operator apply at 2048^2.
The reference cannot build this mesh (32-bit edge key). Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import cuddhelmholtz_b200 as cb

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
out = {"nx": nx}
t0 = time.perf_counter()
mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
out["mesh_s"] = time.perf_counter() - t0
assert mesh.n_edges() == 2 * nx * (nx + 1)
# ---- path A at n_basis 5
t0 = time.perf_counter()
fem5 = cb.H1Space(mesh, cb.Basis(5))
out["h1space_nb5_s"] = time.perf_counter() - t0
n5 = fem5.size()
assert n5 == (4 * nx + 1) ** 2
x = torch.rand(n5, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty_like(x)
t0 = time.perf_counter()
S = cb.StiffnessMatrix(fem5)
out["stiffness_setup_s"] = time.perf_counter() - t0
S.action(x, y)
p, s = S.time_phases(x, y, 10)
out["stiffness_nb5"] = {"ndof": n5, "ms": p + s, "gdofs": n5 / ((p + s) * 1e-3) / 1e9, "algorithmic_gbs": S.algorithmic_bytes() / (p * 1e-3) / 1e9}
one = torch.ones_like(x)
S.action(one, y)
out["stiffness_nb5"]["max_abs_S_times_one"] = float(y.abs().max())
del S, x, y, one, fem5
torch.cuda.empty_cache()
# ---- path B at n_basis 4
omega = 100.0
t0 = time.perf_counter()
fem = cb.H1Space(mesh, cb.Basis(4))
n = fem.size()
assert n == (3 * nx + 1) ** 2
xy = fem.physical_coordinates()
ha = np.where(xy[:, 0] ** 2 + xy[:, 1] ** 2 < 0.0625, 0.2, 1.0)
D = cb.DDH(omega, ha, fem, nx, nx, 16)
out["ddh_setup_s"] = time.perf_counter() - t0
info = D.info()
m = D.size()
out["ddh"] = {"n_domains": info["n_domains"], "nt": info["nt"], "n_lambda": m, "ndof": n}
s2 = omega * omega
src = s2 / np.pi * np.exp(-s2 * ((xy[:, 0] + 0.5) ** 2 + xy[:, 1] ** 2)) + s2 / np.pi * np.exp(-s2 * ((xy[:, 0] - 0.5) ** 2 + (xy[:, 1] + 0.5) ** 2))
f = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
cb.MassMatrix(fem).action(torch.as_tensor(src, device="cuda"), f[:n])
b = torch.empty(m, dtype=torch.float32, device="cuda")
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
D.rhs(f, b)
e1.record()
t = torch.empty_like(b)
D.action(b, t)
e2.record()
torch.cuda.synchronize()
out["ddh"].update({"rhs_s": e0.elapsed_time(e1) / 1e3, "action_s": e1.elapsed_time(e2) / 1e3,
                   "action_fp32_tflops": D.flops() / (e1.elapsed_time(e2) * 1e-3) / 1e12,
                   "rhs_norm": float(b.norm()), "finite": bool(torch.isfinite(t).all())})
print(json.dumps(out))
