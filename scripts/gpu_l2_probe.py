"""Development probe: does the persisting-L2 carve-out of the density kernel slow the kernels that run after it?
build / FOF timings before and after a CalcDensity call at one size."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodylib_b200 import KDTree
from nbodylib_b200.synth import clustered_box
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = ng ** 3
pos, vel, mass = clustered_box(ng, seed=2025, nhalo=max(8, min(8192, n // 16384)), device="cuda")
per = np.ones(3)
g = torch.empty(n, dtype=torch.int32, device="cuda")
rho = torch.empty(n, dtype=torch.float64, device="cuda")
def build():
    t = KDTree(pos, vel, mass, Period=per, device=0)
    return t, t.info.build_ms
def fof(t):
    t.FOF(0.2 / ng, 20, 1, out=g)
    i = t.info
    return i.last_kernel_ms, i.last_call_ms
for rep in range(3):
    t, b = build()
    print("before density: build %.1f ms, fof link %.1f call %.1f ms" % ((b,) + fof(t)), flush=True)
    t.close()
t, b = build()
for rep in range(2):
    t.CalcDensity(64, out=rho)
    print("density: kernel %.1f ms call %.1f ms" % (t.info.last_kernel_ms, t.info.last_call_ms), flush=True)
print("after density: fof link %.1f call %.1f ms" % fof(t), flush=True)
for rep in range(3):
    t2, b = build()
    print("after density: build %.1f ms, fof link %.1f call %.1f ms" % ((b,) + fof(t2)), flush=True)
    t2.close()
t.close()
