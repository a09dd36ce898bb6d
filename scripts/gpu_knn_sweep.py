"""Development probe: CalcDensity kernel time at one size; knobs come from the environment."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodylib_b200 import KDTree
from nbodylib_b200.synth import clustered_box
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ks = [int(a) for a in sys.argv[2:]] or [64]
n = ng ** 3
pos, vel, mass = clustered_box(ng, seed=2025, nhalo=max(8, min(8192, n // 16384)), device="cuda")
if os.environ.get("PROBE_N"):
    n = int(os.environ["PROBE_N"])
    sel = torch.randperm(ng ** 3, device="cuda")[:n]
    pos, vel, mass = pos[sel].contiguous(), vel[sel].contiguous(), mass[sel].contiguous()
t = KDTree(pos, vel, mass, Period=np.ones(3), device=0)
rho = torch.empty(n, dtype=torch.float64, device="cuda")
if os.environ.get("PROBE_TRIM"):
    # before an ncu capture: hand cached blocks back to the driver (torch's allocator, the library's stream-ordered pool), so
    # that the profiler's per-pass save / restore of device memory covers the live arrays only
    del pos, vel, mass
    torch.cuda.empty_cache()
    t._lib.nbk_release_cached_memory(0)
for k in ks:
    for rep in range(2):
        t.CalcDensity(k, out=rho)
    i = t.info
    print("ng %d k %d env %s: kernel %.1f ms -> %.1f Mpart/s flagged %d sum %.6e" % (ng, k, {a: os.environ[a] for a in os.environ if a.startswith("NBK_")}, i.last_kernel_ms, n / i.last_kernel_ms / 1e3, i.last_flagged, rho.sum().item()), flush=True)
