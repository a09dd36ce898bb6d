"""Generates tests/golden/ref_phase.npz from the UNMODIFIED reference (oracle/_ref/libnbref.so): phase-space nearest
neighbours -- KDTree::FindNearestPhase(Int_t tt, ...) and FindNearestPhase(Double_t *x, Double_t *v, ...) on a TPHS tree
built with Aniso = -1 (KDFindNearest.cxx:347-361,543-555), non periodic and periodic -- on the particles of ref_small.npz
with the velocities rescaled so that both halves of the 6D distance matter.  The script also asserts that FindNearest(tt)
on that tree and FindNearestPhase(tt) on a TPHYS tree return the same rows (the result does not depend on the tree the
reference walks).  Run in the build container only:  python tests/golden/make_golden_phase.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref, build_ref  # noqa: E402

assert build_ref() is not None, "needs /root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "ref_small.npz"))
pos, xq = G["pos"], G["xq"]
n = len(pos)
rng = np.random.default_rng(23)
# velocity spread of about a tenth of the box: the 12 nearest in 6D are neither the nearest in position nor in velocity
vel = (G["vel"] * 0.25).astype(np.float32).astype(np.float64)
vq = (rng.normal(size=xq.shape) * vel.std()).astype(np.float32).astype(np.float64)
qsel = np.arange(0, n, 2, dtype=np.int32)
k = 12
out = dict(vel=vel, vq=vq, qsel=qsel, k=np.int32(k))
for tag, period in (("np", None), ("p", np.ones(3))):
    R = Ref(pos, vel, None, treetype=Ref.TPHS, period=period, aniso=-1)
    ids, d2 = R.knn_phase_particles(qsel, k, which=0)
    ids1, d21 = R.knn_phase_particles(qsel, k, which=1)
    assert np.array_equal(ids, ids1) and np.array_equal(d2, d21)
    idx, d2x = R.knn_phase_points(xq, vq, k)
    R.close()
    R = Ref(pos, vel, None, treetype=Ref.TPHYS, period=period)
    ids2, d22 = R.knn_phase_particles(qsel, k, which=0)
    idx2, d2x2 = R.knn_phase_points(xq, vq, k)
    R.close()
    assert np.array_equal(ids, ids2) and np.array_equal(d2, d22) and np.array_equal(idx, idx2) and np.array_equal(d2x, d2x2)
    assert (ids >= 0).all() and (np.diff(d2, axis=1) >= 0).all()
    # the 6D neighbours are not the position neighbours
    pid, _ = Ref(pos, vel, None, period=period).knn_particle_list(qsel, k, which=1)
    assert np.mean([len(set(a) & set(b)) for a, b in zip(ids, pid)]) < 0.8 * k
    out["phase_ids_" + tag], out["phase_d2_" + tag] = ids, d2
    out["phasex_ids_" + tag], out["phasex_d2_" + tag] = idx, d2x
np.savez_compressed(os.path.join(HERE, "ref_phase.npz"), **out)
print("wrote ref_phase.npz:", {a: out[a].shape for a in out})
