"""GPU diagnostic: DDH product vs reference driver vs oracle, stage by stage (nx, nb from argv)."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cuddhelmholtz_b200 as cb
from oracle.rdmp import read_rdmp
nx, nb = int(sys.argv[1]), int(sys.argv[2])
omega = 2 * np.pi * nx / 10
drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
with tempfile.NamedTemporaryFile(suffix=".bin") as f:
    subprocess.check_call([drv, "ddh", str(nx), str(nb), repr(float(omega)), "20", "100", "1e-4", "2024", f.name])
    r = read_rdmp(f.name)
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, float) - np.asarray(b, float)) / max(np.linalg.norm(np.asarray(b, float)), 1e-300))
dev = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")
mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
fem = cb.H1Space(mesh, cb.Basis(nb))
D = cb.DDH(omega, r["a"], fem, nx, nx, 16)
n = D.size(); ndof = fem.size()
f = dev(r["b"])
b = torch.empty(n, dtype=torch.float32, device="cuda"); D.rhs(f, b)
print("rhs rel", rel(b.cpu().numpy(), r["rhs"]), "ref self-spread", rel(r["act_y2"], r["act_y"]))
L = torch.zeros(n, dtype=torch.float32, device="cuda")
out = cb.gmres(n, L, D, b, 20, 100, 1e-4)
print("iters mine", out.num_iter, out.num_matvec, "ref", r["num_iter"][0], r["num_matvec"][0], "success", out.success, r["success"][0])
print("res hist mine", np.array(out.res_norm)[:8], "ref", r["res_norm"][:8])
print("lambda rel", rel(L.cpu().numpy(), r["lambda"]))
U = torch.empty(2 * ndof, dtype=torch.float64, device="cuda")
D.postprocess(L, f, U)
Um, Ur = U.cpu().numpy(), r["U"]
print("U rel", rel(Um, Ur), "u part", rel(Um[:ndof], Ur[:ndof]), "v part", rel(Um[ndof:], Ur[ndof:]))
# postprocess with the REFERENCE lambda: isolates postprocess from the solve
U2 = torch.empty_like(U); D.postprocess(dev(r["lambda"], torch.float32), f, U2)
U2 = U2.cpu().numpy()
print("postprocess(ref lambda) rel", rel(U2, Ur))
d = np.abs(U2 - Ur); k = np.argsort(-d)[:10]
print("worst idx", k, "mine", U2[k], "ref", Ur[k])
print("norms", np.linalg.norm(Um), np.linalg.norm(Ur), "max|ref|", np.abs(Ur).max(), "nan?", np.isnan(Ur).any(), np.isnan(Um).any())
# residual of the reference lambda under my operator and vice versa
y = torch.empty(n, dtype=torch.float32, device="cuda")
D.action(dev(r["lambda"], torch.float32), y); print("my residual of ref lambda", rel(y.cpu().numpy(), b.cpu().numpy()))
D.action(L, y); print("my residual of my lambda", rel(y.cpu().numpy(), b.cpu().numpy()))
