"""Development probe for ncu: builds the 256^3 (or NG^3) clustered box once, then runs FOF(0.2 spacings), FOFCriterion(FOF6d) and
one more build, so that `ncu -k regex:...` can pick the kernels of each."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodylib_b200 import _lib
if os.environ.get("NBK_LIB_FILE"):
    _lib.LIB_PATH = os.path.join(ROOT, "nbodylib_b200", os.environ["NBK_LIB_FILE"])
from nbodylib_b200 import KDTree
from nbodylib_b200.synth import clustered_box
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = ng ** 3
pos, vel, mass = clustered_box(ng, seed=2025, nhalo=max(8, min(8192, n // 16384)), device="cuda")
t = KDTree(pos, vel, mass, Period=np.ones(3), device=0)
g = torch.empty(n, dtype=torch.int32, device="cuda")
_, ng3 = t.FOF(0.2 / ng, 20, 1, out=g)
print("fof3d groups", ng3, "link kernel ms", t.info.last_kernel_ms, flush=True)
sv2 = float(((vel - vel.mean(0)) ** 2).sum(1).mean().item() / 3.0)
params = np.zeros(10)
params[1] = params[6] = (0.2 / ng) ** 2
params[2] = params[7] = (1.25 ** 2) * sv2
_, ng6 = t.FOFCriterion(2, params, 20, 1, out=g)
print("fof6d groups", ng6, "link kernel ms", t.info.last_kernel_ms, flush=True)
t.close()
t = KDTree(pos, vel, mass, Period=np.ones(3), device=0)
print("second build ms", t.info.build_ms, flush=True)
