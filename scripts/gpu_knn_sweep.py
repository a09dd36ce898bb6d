"""Development probe: CalcDensity kernel time at one size for a list of option settings.
usage: python scripts/gpu_knn_sweep.py NG K [opt=val,opt=val ...]   (NBK_LIB_FILE selects an alternative build)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodylib_b200 import _lib
if os.environ.get("NBK_LIB_FILE"):
    _lib.LIB_PATH = os.path.join(ROOT, "nbodylib_b200", os.environ["NBK_LIB_FILE"])
from nbodylib_b200 import KDTree, set_option
from nbodylib_b200.synth import clustered_box
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 256
k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
settings = sys.argv[3:] or [""]
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
reps = int(os.environ.get("PROBE_REPS", "2"))
for st in settings:
    opts = dict(kv.split("=") for kv in st.split(",") if kv)
    for name, v in opts.items():
        set_option(name, int(v))
    for rep in range(reps):
        t.CalcDensity(k, out=rho)
    i = t.info
    print("ng %d k %d opts %s: kernel %.1f ms -> %.1f Mpart/s flagged %d sum %.9e" % (ng, k, opts, i.last_kernel_ms, n / i.last_kernel_ms / 1e3, i.last_flagged, rho.sum().item()), flush=True)
    for name in opts:
        set_option(name, -1 if name == "knn_transpose" else 0)
