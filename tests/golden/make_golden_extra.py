"""Generates tests/golden/ref_extra.npz from the UNMODIFIED reference (oracle/_ref/libnbref.so): the single-target
estimators (CalcDensityParticle / CalcVelDensityParticle / Calc*Position / CalcSmoothLocalValue), criterion search
(SearchCriterionTagged, dense SearchCriterion), dense SearchBallPos and the node getters (FindLeafNode, cut values), on
the particles of ref_small.npz.  Run in the build container only:  python tests/golden/make_golden_extra.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref, build_ref  # noqa: E402

assert build_ref() is not None, "needs /root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "ref_small.npz"))
pos, vel, mass, k = G["pos"], G["vel"], G["mass"], int(G["k"])
n = len(pos)
rng = np.random.default_rng(11)
# unequal masses so that the mass weighting of the gather sum is visible (fp32 representable like every input)
mass2 = (mass * (1.0 + rng.random(n))).astype(np.float32).astype(np.float64)
qsel = np.arange(1, n, 3, dtype=np.int32)
xq = G["xq"]
vq = rng.normal(size=xq.shape).astype(np.float32).astype(np.float64) * float(np.std(vel))
params = G["params"]
ll = float(G["ll"])
# criterion searches around particles that are NOT in the tree: displaced copies of tree particles, so that rows are not empty
jn = rng.integers(0, n, 200)
xn = (pos[jn] + 0.4 * ll * rng.normal(size=(200, 3))).astype(np.float32).astype(np.float64)
vn = (vel[jn] + 0.3 * np.sqrt(params[7]) * rng.normal(size=(200, 3))).astype(np.float32).astype(np.float64)
# FindNearestCheck: a third of the particles fail the check (Particle::type != 0 in the driver's FOFcheckfunc)
types = (rng.random(n) < 0.3).astype(np.int32)
kf = 8
out = dict(mass2=mass2, qsel=qsel, vq=vq, kv=np.int32(k // 2), xn=xn, vn=vn, types=types, kf=np.int32(kf))
for tag, period in (("np", None), ("p", np.ones(3))):
    R = Ref(pos, vel, mass2, period=period, kerntype=Ref.KEPAN, kernres=1000)
    out["dens_part_" + tag] = R.calc_density_particles(qsel, k)
    out["vdens_part_" + tag] = R.calc_veldensity_particles(qsel, k // 2, k)
    out["dens_pos_" + tag] = R.calc_density_points(xq, k)
    out["vdens_pos_" + tag] = R.calc_veldensity_points(xq, vq, k // 2, k)
    for crit, name in ((0, "c3"), (2, "c6")):
        off, idx = R.search_criterion_particles(qsel, crit, params)
        out["%s_off_%s" % (name, tag)] = off
        out["%s_idx_%s" % (name, tag)] = np.concatenate([np.sort(idx[off[i]:off[i + 1]]) for i in range(len(qsel))])
        off, idx = R.search_criterion_points(xn, vn, crit, params)
        assert len(idx) > 100
        out["%sx_off_%s" % (name, tag)] = off
        out["%sx_idx_%s" % (name, tag)] = np.concatenate([np.sort(idx[off[i]:off[i + 1]]) for i in range(len(xn))])
    # filtered kNN: FindNearestCheck(tt | Coordinate), FindNearestCriterion(tt | Particle) (KDFindNearest.cxx:363-441)
    R.set_types(types)
    out["nnchk_ids_" + tag], out["nnchk_d2_" + tag] = R.knn_filtered(kf)
    out["nnchkx_ids_" + tag], out["nnchkx_d2_" + tag] = R.knn_filtered(kf, x=xq)
    for crit, name in ((0, "c3"), (2, "c6")):
        out["nn%s_ids_%s" % (name, tag)], out["nn%s_d2_%s" % (name, tag)] = R.knn_filtered(kf, crit=crit, params=params)
        out["nn%sx_ids_%s" % (name, tag)], out["nn%sx_d2_%s" % (name, tag)] = R.knn_filtered(kf, crit=crit, params=params, x=xn, v=vn)
    R.set_types(np.zeros(n, dtype=np.int32))
    # dense forms for a handful of targets: marks accumulate over calls like a caller's tagging loop would
    dq = qsel[:: max(1, len(qsel) // 12)][:12]
    out["dense_q"] = dq
    nn = np.zeros(n, dtype=np.int32)
    d2 = np.zeros(n)
    for j, q in enumerate(dq):
        R.search_ball_dense(int(q), (2.5 * ll) ** 2, j + 1, nn, d2)
    out["dense_ball_nn_" + tag], out["dense_ball_d2_" + tag] = nn.copy(), d2.copy()
    nn[:] = 0
    d2[:] = 0
    for j, x in enumerate(xq[:12]):
        R.search_ball_dense(x, (2.5 * ll) ** 2, j + 1, nn, d2)
    out["dense_ballx_nn_" + tag], out["dense_ballx_d2_" + tag] = nn.copy(), d2.copy()
    nn[:] = 0
    d2[:] = 0
    for j, q in enumerate(dq):
        R.search_criterion_dense(int(q), 2, params, j + 1, nn, d2)
    out["dense_c6_nn_" + tag], out["dense_c6_d2_" + tag] = nn.copy(), d2.copy()
    if tag == "np":
        out["sm_rho"], out["sm_vel"], out["sm_disp"] = R.calc_smooth_vel(k)
        out["sm_skew"], out["sm_kurt"] = R.calc_smooth_higher(k)
        ids, d2k = R.knn_particles(k)
        dist = np.sqrt(d2k[7][::-1]).copy()
        w = mass2[ids[7][::-1]].copy()
        out["slv_dist"], out["slv_weight"] = dist, w
        out["slv_value"] = np.float64(R.smooth_local_value(dist, w))
        leaves = [R.find_leaf(int(q)) for q in dq]
        out["leaf_off"] = np.cumsum([0] + [len(x) for x in leaves]).astype(np.int64)
        out["leaf_ids"] = np.concatenate(leaves).astype(np.int32)
        leavesx = [R.find_leaf(x) for x in xq[:12]]
        out["leafx_off"] = np.cumsum([0] + [len(x) for x in leavesx]).astype(np.int64)
        out["leafx_ids"] = np.concatenate(leavesx).astype(np.int32)
        cid, cdim, cval, _ = R.dump_cuts()
        out["cut_ids"], out["cut_dims"], out["cut_vals"] = cid, cdim, cval
    R.close()
np.savez_compressed(os.path.join(HERE, "ref_extra.npz"), **out)
print("wrote ref_extra.npz", {k_: getattr(v, "shape", None) for k_, v in out.items()})
