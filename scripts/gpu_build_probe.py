"""Development probe: build only (twice: cold pool, warm pool) at one size."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodylib_b200 import KDTree
from nbodylib_b200.synth import clustered_box
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = ng ** 3
pos, vel, mass = clustered_box(ng, seed=2025, nhalo=max(8, min(8192, n // 16384)), device="cuda")
for rep in range(reps):
    torch.cuda.synchronize(); t0 = time.time()
    t = KDTree(pos, vel, mass, Period=np.ones(3), device=0)
    torch.cuda.synchronize()
    i = t.info
    print("ng %d rep %d: create wall %.1f ms build_ms %.1f (%.1f Mpart/s) stage %.1f ms" % (ng, rep, (time.time() - t0) * 1e3, i.build_ms, n / i.build_ms / 1e3, i.h2d_ms), flush=True)
    t.close()
