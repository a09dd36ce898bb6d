"""Development probe: timings of build / CalcDensity / CalcVelDensity / FOF on the clustered box at growing sizes."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodylib_b200 import KDTree, FOF6D  # noqa: E402
from nbodylib_b200.synth import clustered_box  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [128, 256]
for ng in sizes:
    n = ng ** 3
    nh = max(8, min(8192, n // 4096))
    t0 = time.time()
    pos, vel, mass = clustered_box(ng, seed=2025, nhalo=nh, device="cuda")
    torch.cuda.synchronize()
    print("\n== ng=%d n=%d nhalo=%d gen %.2fs mem %.1f GB" % (ng, n, nh, time.time() - t0, torch.cuda.memory_allocated() / 1e9), flush=True)
    t0 = time.time()
    t = KDTree(pos, vel, mass, Period=np.ones(3), device=0)
    torch.cuda.synchronize()
    i = t.info
    print(" create wall %.3fs  build_ms %.1f (%.1f Mpart/s)  stage %.1f ms  nodes %d depth %d dev %.2f GB" % (
        time.time() - t0, i.build_ms, n / i.build_ms / 1e3, i.h2d_ms, i.num_nodes, i.depth, i.device_bytes / 1e9), flush=True)
    rho = torch.empty(n, dtype=torch.float64, device="cuda")
    for k in (32, 64):
        t.CalcDensity(k, out=rho)
        i = t.info
        print(" CalcDensity(%d): kernel %.1f ms  -> %.1f Mpart/s ; call %.1f ms ; mean rho %.4g" % (k, i.last_kernel_ms, n / i.last_kernel_ms / 1e3, i.last_call_ms, rho.mean().item() / n), flush=True)
    t.CalcVelDensity(64, 64, out=rho)
    i = t.info
    print(" CalcVelDensity(64,64): kernel %.1f ms -> %.1f Mpart/s" % (i.last_kernel_ms, n / i.last_kernel_ms / 1e3), flush=True)
    g = torch.empty(n, dtype=torch.int32, device="cuda")
    ll = 0.2 / ng
    _, ng_ = t.FOF(ll, 20, 1, out=g)
    i = t.info
    print(" FOF(0.2): link kernel %.1f ms, call %.1f ms -> %.1f Mpart/s ; groups %d grouped frac %.3f" % (
        i.last_kernel_ms, i.last_call_ms, n / i.last_call_ms / 1e3, ng_, (g > 0).float().mean().item()), flush=True)
    t.close()
    del pos, vel, mass, rho, g
    torch.cuda.empty_cache()
