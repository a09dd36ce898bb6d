"""Generates tests/golden/ref_small.npz from the UNMODIFIED reference (oracle/_ref/libnbref.so, built by
oracle/Makefile from /root/reference).  Run in the build container only:  python tests/golden/make_golden.py
The fixture pins the oracle port (and the CUDA path) on machines where /root/reference is absent."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nbodylib_b200.synth import clustered_small  # noqa: E402
from oracle.pyoracle import Ref, build_ref  # noqa: E402

assert build_ref() is not None, "needs /root/reference"
n, k = 2500, 16
pos, vel, mass = clustered_small(n, seed=42, nhalo=10)
out = dict(pos=pos, vel=vel, mass=mass, k=np.int32(k))
ll = 0.25 / n ** (1.0 / 3)
sv = 0.6 * np.sqrt(((vel - vel.mean(0)) ** 2).sum(1).mean() / 3)
params = np.zeros(10)
params[1] = params[6] = (1.4 * ll) ** 2
params[2] = params[7] = sv ** 2
out["ll"] = np.float64(ll)
out["params"] = params
rng = np.random.default_rng(7)
xq = rng.random((300, 3))
out["xq"] = xq
for tag, period in (("np", None), ("p", np.ones(3))):
    R = Ref(pos, vel, mass, period=period, kerntype=Ref.KEPAN, kernres=1000)
    out["kernnorm_epan"] = np.float64(R.kernnorm)
    out["nodes"] = np.array([R.num_nodes, R.num_leaves], dtype=np.int64)
    for which in (0, 1):
        ids, d2 = R.knn_particles(k, which=which)
        out["knn%d_ids_%s" % (which, tag)] = ids
        out["knn%d_d2_%s" % (which, tag)] = d2
    ids, d2 = R.knn_points(xq, k)
    out["knnx_ids_" + tag] = ids
    out["knnx_d2_" + tag] = d2
    out["rho_" + tag] = R.calc_density(k)
    out["vrho_" + tag] = R.calc_veldensity(8, k)
    g, ng = R.fof(ll, 5, 0)
    out["fof_" + tag] = g
    g, ng = R.fof(ll, 5, 1)
    out["fof_ord_" + tag] = g
    g, ng = R.fof_criterion(2, params, 5, 0)
    out["fof6d_" + tag] = g
    g, ng = R.fof_criterion(0, params, 5, 0)
    out["fof3d_" + tag] = g
    off, idx = R.ball_points(xq, (3 * ll) ** 2)
    out["ball_off_" + tag] = off
    out["ball_idx_" + tag] = np.concatenate([np.sort(idx[off[i]:off[i + 1]]) for i in range(len(xq))]) if len(idx) else idx
    R.close()
R = Ref(pos, vel, mass, kerntype=Ref.KSPH, kernres=1000)
out["kernnorm_sph"] = np.float64(R.kernnorm)
out["rho_sph_np"] = R.calc_density(k)
R.close()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_small.npz"), **out)
print("wrote ref_small.npz", {k_: getattr(v, "shape", None) for k_, v in out.items()})
