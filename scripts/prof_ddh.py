"""tiny driver for ncu: a few DDH actions at the reference example's size (uniform_rect(128), n_basis 4, omega = 2 pi 12.8)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cuddhelmholtz_b200 as cb
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 128
omega = 2 * np.pi * nx / 10
mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
fem = cb.H1Space(mesh, cb.Basis(4))
D = cb.DDH(omega, np.ones(fem.size()), fem, nx, nx, 16)
x = torch.rand(D.size(), dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
for _ in range(3):
    D.action(x, y)
torch.cuda.synchronize()
print("ok", D.info())
