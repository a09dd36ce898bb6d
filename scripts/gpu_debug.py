"""Development probe run on the GPU box: compares every device entry point with the oracle port (and the
reference build when oracle/_ref is present) and prints diagnostics instead of stopping at the first
mismatch.  Not part of the test-suite."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodylib_b200 import KDTree, FOF3D, FOF6D, TPHS  # noqa: E402
from nbodylib_b200.synth import clustered_small, uniform_box  # noqa: E402
from oracle.pyoracle import Port, Ref, canonical_groups, have_ref  # noqa: E402

P = Port()


def section(name):
    print("\n==== " + name, flush=True)


def check_tree(t, pos, bucket):
    order = t.order()
    n = len(pos)
    ok = np.array_equal(np.sort(order), np.arange(n))
    print(" order is a permutation:", ok)
    s, e, c, b = t.nodes()
    info = t.info
    present = s >= 0
    leaf = present & ((e - s) <= bucket)
    print(" nodes", info.num_nodes, "present", present.sum(), "leaves", info.num_leaves, "leafcount", leaf.sum(), "depth", info.depth)
    tp = pos[order]
    bad = 0
    for i in np.nonzero(present)[0][:200000]:
        q = tp[s[i]:e[i]]
        lo, hi = q.min(0), q.max(0)
        if not (np.all(b[i, 0::2] <= lo) and np.all(b[i, 1::2] >= hi) and np.allclose(b[i, 0::2], lo, rtol=1e-6, atol=1e-7) and np.allclose(b[i, 1::2], hi, rtol=1e-6, atol=1e-7)):
            bad += 1
            if bad < 4:
                print("  bad bbox node", i, s[i], e[i], b[i], lo, hi)
    print(" bad bboxes:", bad)
    # children partition & split rule
    badsplit = 0
    for i in np.nonzero(present & ~leaf)[0][:100000]:
        l, r = 2 * i + 1, 2 * i + 2
        m = s[i] + (e[i] - s[i] - 1) // 2
        if not (s[l] == s[i] and e[l] == m + 1 and s[r] == m + 1 and e[r] == e[i]):
            badsplit += 1
            continue
        d = c[i]
        q = tp[s[i]:e[i]]
        ext = q.max(0) - q.min(0)
        if d != int(np.argmax(ext)) and ext[d] < ext.max():
            badsplit += 1
        if tp[s[l]:e[l], d].max() > tp[s[r]:e[r], d].min():
            badsplit += 1
    print(" bad splits:", badsplit)
    return order


def run(name, fn):
    try:
        fn()
    except Exception:
        print("!! %s raised" % name)
        traceback.print_exc()


def small_suite(n, period, store_flag=0, seed=3):
    pos, vel, mass = clustered_small(n, seed=seed)
    section("n=%d period=%s flags=%d" % (n, period is not None, store_flag))
    t0 = time.time()
    t = KDTree(pos, vel, mass, bucket_size=16, Period=period, device=0, flags=store_flag)
    print(" create %.3fs build_ms %.2f h2d_ms %.2f store %d inexact %d" % (time.time() - t0, t.info.build_ms, t.info.h2d_ms, t.info.store_bytes, t.info.inexact_coords))
    order = check_tree(t, pos, 16)

    def knn():
        for k in (8, 32, 64):
            for tree_form in (False, True):
                nn, d2 = t.FindNearestPos(k, ids=True, tree_form=tree_form)
                io, do = P.knn_particles(pos, k, period=period, which=1 if tree_form else 0)
                io, do = io[order], do[order]
                print(" knn k=%d tree_form=%d: d2 equal %s, id rows equal %.6f, sets equal %.6f, ms %.3f" % (
                    k, tree_form, np.array_equal(d2, do), (nn == io).all(1).mean(), (np.sort(nn, 1) == np.sort(io, 1)).all(1).mean(), t.info.last_kernel_ms))
                if not np.array_equal(d2, do):
                    bad = np.nonzero((d2 != do).any(1))[0]
                    print("   first bad rows", bad[:5], "count", len(bad))
                    r = bad[0]
                    print("   dev", d2[r][:6], nn[r][:6], "\n   orc", do[r][:6], io[r][:6])
    run("knn", knn)

    def knn_points():
        rng = np.random.default_rng(5)
        x = rng.random((3000, 3))
        nn, d2 = t.FindNearestPosPoints(x, 16, ids=True)
        io, do = P.knn_points(pos, x, 16, period=period)
        print(" knn points: d2 equal", np.array_equal(d2, do), "sets equal", (np.sort(nn, 1) == np.sort(io, 1)).all(1).mean())
        if period is not None:
            nn, d2 = t.FindNearestPosPoints(x, 16, ids=True, strict=True)
            io, do = P.knn_points(pos, x, 16, period=period, strict=1)
            print(" knn points strict: d2 equal", np.array_equal(d2, do))
    run("knn_points", knn_points)

    def dens():
        for k in (16, 64):
            rho, h = t.CalcDensity(k, want_h=True)
            ro, ho = P.density(pos, mass, k)
            print(" density k=%d: max rel err %.3e  h max rel err %.3e  ms %.3f" % (k, np.abs(rho / ro - 1).max(), np.abs(h / ho - 1).max(), t.info.last_kernel_ms))
        for kv, kx in ((16, 32), (32, 32)):
            rv = t.CalcVelDensity(kv, kx)
            rvo = P.veldensity(pos, vel, kv, kx)
            print(" veldensity %d/%d: max rel err %.3e  exact %s" % (kv, kx, np.abs(rv / rvo - 1).max(), np.array_equal(rv, rvo)))
        hh = t.CalcSmoothingScale(32)
        _, ho = P.density(pos, mass, 32)
        print(" smoothing scale equal:", np.array_equal(hh, ho))
    run("dens", dens)

    def fof():
        ll = 0.2 / n ** (1.0 / 3)
        for order_flag in (0, 1):
            g, ng = t.FOF(ll, 8, order_flag)
            go, ngo = P.fof(pos, None, 0, [ll * ll], period, 8, order_flag)
            same = np.array_equal(canonical_groups(g), canonical_groups(go))
            print(" FOF order=%d: ng %d vs %d, partition equal %s, grouped %d, ms %.3f" % (order_flag, ng, ngo, same, (g > 0).sum(), t.info.last_kernel_ms))
            if order_flag:
                print("   sizes descending:", bool(np.all(np.diff(np.bincount(g)[1:]) <= 0)), "labels identical:", np.array_equal(g, go))
        sv = np.sqrt(((vel - vel.mean(0)) ** 2).sum(1).mean() / 3)
        params = np.zeros(10)
        params[1] = params[6] = (1.5 * ll) ** 2
        params[2] = params[7] = (0.5 * sv) ** 2
        g, ng = t.FOFCriterion(FOF6D, params, 8, 0)
        go, ngo = P.fof(pos, vel, 4, params, period, 8, 0)
        print(" FOF6d: ng %d vs %d, partition equal %s, grouped %d" % (ng, ngo, np.array_equal(canonical_groups(g), canonical_groups(go)), (g > 0).sum()))
        g, ng = t.FOFCriterion(FOF3D, params, 8, 0)
        go, ngo = P.fof(pos, vel, 2, params, period, 8, 0)
        print(" FOF3d: ng %d vs %d, partition equal %s" % (ng, ngo, np.array_equal(canonical_groups(g), canonical_groups(go))))
        pre = (np.arange(n) % 7 == 0).astype(np.int32)
        g, ng = t.FOF(ll, 4, 0, precheck=pre)
        print(" FOF precheck: excluded all zero:", bool((g[pre != 0] == 0).all()), "ng", ng)
    run("fof", fof)

    def ball():
        rng = np.random.default_rng(9)
        x = rng.random((2000, 3))
        r2 = (1.5 / n ** (1.0 / 3)) ** 2
        off, idx = t.SearchBallPosTaggedPoints(x, r2, ids=True)
        oo, io = P.ball_points(pos, x, r2, period)
        same = np.array_equal(off, oo) and all(np.array_equal(np.sort(idx[off[i]:off[i + 1]]), io[oo[i]:oo[i + 1]]) for i in range(len(x)))
        print(" ball points: total %d vs %d equal %s" % (off[-1], oo[-1], same))
        tt = rng.integers(0, n, 2000).astype(np.int32)
        off, idx = t.SearchBallPosTagged(tt, r2, ids=True)
        oo, io = P.ball_points(pos, pos[order[tt]], r2, period)
        cnt_dev = np.diff(off)
        cnt_o = np.diff(oo) - (0 if period is not None else 1)
        print(" ball particles: counts equal", np.array_equal(cnt_dev, cnt_o))
    run("ball", ball)
    t.close()


def tphs_suite(n):
    section("TPHS FOF form A vs FOF6d form B, n=%d" % n)
    pos, vel, mass = clustered_small(n, seed=11)
    ll = 0.3 / n ** (1.0 / 3)
    sv = 0.5 * np.sqrt(((vel - vel.mean(0)) ** 2).sum(1).mean() / 3)
    for period in (None, np.ones(3)):
        ps, vs = pos / ll, vel / sv           # what Particle::ScalePhase does (x *= 1/ll)
        ps, vs = pos * (1.0 / ll), vel * (1.0 / sv)
        pA = None if period is None else period * (1.0 / ll)
        tA = KDTree(ps, vs, mass, TreeType=TPHS, Period=pA, device=0)
        print(" store bytes (scaled coords):", tA.info.store_bytes)
        gA, ngA = tA.FOF(1.0, 8, 0)
        goA, ngoA = P.fof(ps, vs, 1, [1.0], pA, 8, 0)
        print(" form A periodic=%s: ng %d vs oracle %d equal %s" % (period is not None, ngA, ngoA, np.array_equal(canonical_groups(gA), canonical_groups(goA))))
        tA.close()
        tB = KDTree(pos, vel, mass, Period=period, device=0)
        params = np.zeros(10)
        params[1] = params[6] = ll * ll
        params[2] = params[7] = sv * sv
        gB, ngB = tB.FOFCriterion(FOF6D, params, 8, 0)
        goB, ngoB = P.fof(pos, vel, 4, params, period, 8, 0)
        print(" form B: ng %d vs oracle %d equal %s ; A==B partitions: %s" % (ngB, ngoB, np.array_equal(canonical_groups(gB), canonical_groups(goB)),
                                                                          np.array_equal(canonical_groups(gA), canonical_groups(gB))))
        tB.close()


def ref_suite(n):
    if not have_ref():
        print("no oracle/_ref")
        return
    section("vs reference library n=%d" % n)
    pos, vel, mass = uniform_box(n)
    period = np.ones(3)
    t0 = time.time()
    t = KDTree(pos, vel, mass, Period=period, device=0)
    print(" create %.3fs build_ms %.2f h2d %.2f nodes %d leaves %d" % (time.time() - t0, t.info.build_ms, t.info.h2d_ms, t.info.num_nodes, t.info.num_leaves))
    order = t.order()
    R = Ref(pos, vel, mass, period=period)
    print(" ref build %.3fs nodes %d leaves %d threads %d" % (R.build_seconds, R.num_nodes, R.num_leaves, Ref.max_threads()))
    t0 = time.time()
    nn, d2 = t.FindNearestPos(32, ids=True)
    tdev = time.time() - t0
    ir, dr = R.knn_particles(32, which=0)
    print(" kNN k=32 periodic: dev call %.3fs kernel %.2f ms ; ref %.3fs" % (tdev, t.info.last_kernel_ms, R.last_seconds))
    ir, dr = ir[order], dr[order]
    print("   d2 equal", np.array_equal(d2, dr), "sets equal", (np.sort(nn, 1) == np.sort(ir, 1)).all(1).mean())
    t0 = time.time()
    rho = t.CalcDensity(32)
    tdev = time.time() - t0
    rr = R.calc_density(32)
    print(" CalcDensity(32): dev call %.3fs kernel %.2f ms ; ref serial %.3fs ; max rel err %.3e" % (tdev, t.info.last_kernel_ms, R.last_seconds, np.abs(rho / rr - 1).max()))
    rv = t.CalcVelDensity(32, 32)
    rvr = R.calc_veldensity(32, 32)
    print(" CalcVelDensity: kernel %.2f ms ; ref %.3fs ; max rel err %.3e" % (t.info.last_kernel_ms, R.last_seconds, np.abs(rv / rvr - 1).max()))
    ll = 0.2 / n ** (1.0 / 3)
    g, ng = t.FOF(ll, 2, 1)
    gr, ngr = R.fof(ll, 2, 1)
    print(" FOF: kernel %.2f ms call %.2f ms ; ref %.3fs ; ng %d vs %d ; equal %s" % (t.info.last_kernel_ms, t.info.last_call_ms, R.last_seconds, ng, ngr, np.array_equal(canonical_groups(g), canonical_groups(gr))))
    t.close()
    R.close()


if __name__ == "__main__":
    import torch
    print(torch.cuda.get_device_name(0))
    small_suite(20000, None)
    small_suite(20000, np.ones(3))
    small_suite(5000, np.ones(3), store_flag=1 << 4)   # forced fp64 storage
    small_suite(777, None)
    small_suite(13, None)
    tphs_suite(12000)
    ref_suite(1000000)
