"""DDH-GMRES solve-time ladder on one GPU (examples/DDH.cpp flow, omega = 2 pi nx / 10), next to the unmodified
reference library (oracle/_ref/ref_driver ddh) where it is run. Prints one JSON line per size."""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.rdmp import read_rdmp
sizes = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [128, 256, 512]
ref_sizes = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 and sys.argv[2] else []
drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
for nx in sizes:
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ddh_multi.py"), str(nx), "4"], capture_output=True, text=True)
    mine = json.loads(r.stdout.strip().splitlines()[-1])
    line = {"nx": nx, "n_domains": mine["n_domains"], "ours": {k: mine[k] for k in ("action_ms", "gmres_seconds", "restarts", "matvec", "success")}}
    if nx in ref_sizes and os.path.exists(drv):
        import numpy as np
        with tempfile.NamedTemporaryFile(suffix=".bin") as f:
            subprocess.check_call([drv, "ddh", str(nx), "4", repr(float(2 * np.pi * nx / 10)), "20", "100", "1e-4", "2024", f.name])
            d = read_rdmp(f.name)
        line["reference"] = {"gmres_seconds": float(d["gmres_seconds"][0]), "restarts": int(d["num_iter"][0]), "matvec": int(d["num_matvec"][0]),
                             "success": bool(d["success"][0])}
        line["solve_speedup"] = line["reference"]["gmres_seconds"] / mine["gmres_seconds"]
    print(json.dumps(line), flush=True)
