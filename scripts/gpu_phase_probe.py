"""Timing probe of the phase-space kNN (nbk_knn_phase_particles) next to the position kNN on the same particles.
Usage (GPU box): python scripts/gpu_phase_probe.py [n] [k]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import nbodylib_b200 as nb  # noqa: E402
from nbodylib_b200.synth import clustered_small  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
pos, vel, mass = clustered_small(n, seed=1)
for vscale in (0.02, 0.2):
    v = (vel * (vscale / vel.std())).astype(np.float32)
    with nb.KDTree(pos.astype(np.float32), v, None, TreeType=nb.TPHS, Aniso=-1) as t:
        for rep in range(2):
            t0 = time.time()
            nn, d2 = t.FindNearestPhase(k)
            wall = time.time() - t0
        print("phase kNN n=%d k=%d sigma_v=%.2f: kernel %.2f ms, call %.1f ms wall (host copies included)" % (n, k, vscale, t.info.last_kernel_ms, wall * 1e3), flush=True)
with nb.KDTree(pos.astype(np.float32), None, None) as t:
    for rep in range(2):
        nn, d2 = t.FindNearestPos(k)
    print("position kNN n=%d k=%d: kernel %.2f ms" % (n, k, t.info.last_kernel_ms), flush=True)
