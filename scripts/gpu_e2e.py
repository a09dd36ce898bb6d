"""Development probe: where does the host-buffer (e2e) path spend its time?"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodylib_b200 import KDTree
from nbodylib_b200.synth import clustered_box
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = ng ** 3
pos, vel, mass = clustered_box(ng, seed=2025, nhalo=max(8, min(8192, n // 16384)), device="cuda")
hp, hv, hm = (x.cpu().pin_memory().numpy() for x in (pos, vel, mass))
del pos, vel, mass
torch.cuda.empty_cache()
for it in range(3):
    t0 = time.perf_counter()
    t = KDTree(hp, hv, hm, Period=np.ones(3), device=0)
    t1 = time.perf_counter()
    rho = t.CalcDensity(64)
    t2 = time.perf_counter()
    i = t.info
    t.close()
    t3 = time.perf_counter()
    print("iter %d: create %.3f (stage %.1f ms, build %.1f ms)  CalcDensity call %.3f (kernel %.1f ms)  close %.3f  total %.3f" % (
        it, t1 - t0, i.h2d_ms, i.build_ms, t2 - t1, i.last_kernel_ms, t3 - t2, t3 - t0), flush=True)
