"""GPU: warp-specialised / thread-per-element / lane-per-row patch kernels at the headline size (one subprocess per setting)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sweep_ops import run
if __name__ == "__main__":
    sets = [json.loads(a) for a in sys.argv[1:]] or [{}, {"CUDDH_B200_WS": 0}, {"CUDDH_B200_TPE": 0}]
    for nb in (5, 4):
        for s in sets:
            print("nb", nb, json.dumps(s), json.dumps(run(s, 1024, nb)), flush=True)
