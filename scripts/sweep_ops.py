"""GPU: time the stiffness / weighted-mass patch kernels at the headline size under the experiment knobs
(CUDDH_B200_MAXR register cap, CUDDH_B200_WARPS warps per CTA). One subprocess per setting (knobs are read once)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import torch, cuddhelmholtz_b200 as cb
nx, nb = int(sys.argv[1]), int(sys.argv[2])
mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
fem = cb.H1Space(mesh, cb.Basis(nb))
n = fem.size()
x = torch.rand(n, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty_like(x)
a = torch.rand(n, dtype=torch.float64, device="cuda") + 0.5
out = {}
for name, op in (("S", cb.StiffnessMatrix(fem)), ("Mw", cb.MassMatrix(a, fem))):
    op.action(x, y)
    p, s = op.time_phases(x, y, 20)
    out[name] = dict(ms=p, shared_ms=s, gbs=op.algorithmic_bytes() / p / 1e6)
print(json.dumps(out))
''' % ROOT


def run(env, nx=1024, nb=5):
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    r = subprocess.run([sys.executable, "-c", CHILD, str(nx), str(nb)], capture_output=True, text=True, env=e)
    try:
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        return {"error": r.stderr[-400:]}


if __name__ == "__main__":
    settings = [{}, {"CUDDH_B200_WARPS": 4}, {"CUDDH_B200_WARPS": 3}]
    for s in settings:
        print(json.dumps(s), json.dumps(run(s)), flush=True)
