"""One fused Helmholtz apply on uniform_rect(nx) for n_basis nb with whatever kernel variant the environment selects
(CUDDH_B200_AFFINE_RING, CUDDH_B200_AFFINE, ...): prints a JSON line with a hash of the result (variants that only move the metric
data differently must agree bit for bit), the result norm and the kernel time (median of 30 launches).
   python scripts/fused_variant.py 1024 5"""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cuddhelmholtz_b200 as cb
if os.environ.get("CUDDH_B200_LIBPATH"):  # experiments: an alternative build of the library (scripts only, not a product knob)
    from cuddhelmholtz_b200 import capi as _capi
    _capi.LIB_PATH = os.environ["CUDDH_B200_LIBPATH"]
nx, nb = int(sys.argv[1]), int(sys.argv[2])
mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
fem = cb.H1Space(mesh, cb.Basis(nb))
fs = cb.FaceSpace(fem, mesh.boundary_edges())
n = fem.size()
g = torch.Generator(device="cuda").manual_seed(7)
x = torch.rand(2 * n, dtype=torch.float64, device="cuda", generator=g) - 0.5
a = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) + 0.5
af = torch.ones(fs.size(), dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
H = cb.Helmholtz(100.0, a, af, fem, fs)
H.action(x, y)
torch.cuda.synchronize()
h = hashlib.sha1(y.cpu().numpy().tobytes()).hexdigest()[:16]
p, s = H.time_phases(x, y, 30)
print(json.dumps({"nx": nx, "nb": nb, "env": {k: v for k, v in os.environ.items() if k.startswith("CUDDH_B200")}, "sha1": h,
                  "norm": float(y.norm()), "kernel_ms": round(p, 4), "rest_ms": round(s, 4),
                  "hbm_frac": round(H.algorithmic_bytes() / (p * 1e-3) / 1e9 / 6459.0, 4)}))
