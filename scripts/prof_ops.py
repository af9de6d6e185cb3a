"""tiny driver for ncu: a few stiffness / weighted-mass applies at the headline size"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cuddhelmholtz_b200 as cb
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
fem = cb.H1Space(mesh, cb.Basis(nb))
n = fem.size()
x = torch.rand(n, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty_like(x)
a = torch.rand(n, dtype=torch.float64, device="cuda") + 0.5
S, M = cb.StiffnessMatrix(fem), cb.MassMatrix(a, fem)
for _ in range(3):
    S.action(x, y)
    M.action(x, y)
torch.cuda.synchronize()
print("ok")
