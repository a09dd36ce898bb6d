"""GPU parity suite (-m gpu): every call goes through the C ABI (nbodylib_b200.KDTree -> libnbk.so) and is
compared with (a) the golden vectors produced by the reference, (b) the brute-force oracle port on seeded
inputs, (c) the reference library itself at BASELINE config 1 when oracle/_ref travelled with the repo, and
(d) size-independent properties at sizes the CPU checkers cannot reach.

Bars: neighbour sets, d2 values, FOF partitions, ball-search sets: bit-exact.  Densities: 1e-10 relative here
(the north star asks for 1e-5; the device accumulates in fp64, only the summation order differs)."""
import numpy as np
import pytest

from tests.util import canon, csr_rows_sorted, load_golden, load_golden_extra, rows_equal_as_sets

pytestmark = pytest.mark.gpu

RTOL_RHO = 1e-10


@pytest.fixture(scope="module")
def nb(built):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import nbodylib_b200
    return nbodylib_b200


@pytest.fixture(scope="module")
def G():
    return load_golden()


def by_id(order, rows):
    """rows indexed by tree index -> indexed by ID"""
    out = np.empty_like(rows)
    out[order] = rows
    return out


@pytest.mark.parametrize("tag", ["np", "p"])
def test_golden_knn(nb, G, tag):
    period = None if tag == "np" else np.ones(3)
    k = int(G["k"])
    with nb.KDTree(G["pos"], G["vel"], G["mass"], Period=period) as t:
        assert (t.GetNumNodes(), t.GetNumLeafNodes()) == tuple(G["nodes"])
        assert t.GetKernNorm() == float(G["kernnorm_epan"])
        assert t.info.store_bytes == 4 and t.info.inexact_coords == 0
        order = t.order()
        for which in (0, 1):
            nn, d2 = t.FindNearestPos(k, ids=True, tree_form=bool(which))
            assert np.array_equal(by_id(order, d2), G["knn%d_d2_%s" % (which, tag)])
            assert rows_equal_as_sets(by_id(order, nn), G["knn%d_ids_%s" % (which, tag)])
            assert np.all(np.diff(d2, axis=1) >= 0)
        # tree-index outputs map through order() to the same IDs (KDTree.h:295 index convention)
        nn_t, _ = t.FindNearestPos(k)
        nn_i, _ = t.FindNearestPos(k, ids=True)
        assert np.array_equal(order[nn_t], nn_i)
        nn, d2 = t.FindNearestPosPoints(G["xq"], k, ids=True)
        assert np.array_equal(d2, G["knnx_d2_" + tag]) and rows_equal_as_sets(nn, G["knnx_ids_" + tag])


@pytest.mark.parametrize("tag", ["np", "p"])
def test_golden_density(nb, G, tag):
    period = None if tag == "np" else np.ones(3)
    k = int(G["k"])
    with nb.KDTree(G["pos"], G["vel"], G["mass"], Period=period) as t:
        rho, h = t.CalcDensity(k, want_h=True)
        np.testing.assert_allclose(rho, G["rho_" + tag], rtol=RTOL_RHO)
        d2k = G["knn0_d2_np"][:, -1]                      # Calc* ignore the period (quirk Q2)
        assert np.array_equal(h, 0.5 * np.sqrt(d2k))
        assert np.array_equal(t.CalcSmoothingScale(k), h)
        np.testing.assert_allclose(t.CalcVelDensity(8, k), G["vrho_" + tag], rtol=RTOL_RHO)
    if tag == "np":
        with nb.KDTree(G["pos"], G["vel"], G["mass"], KernType=nb.KSPH) as t:
            assert t.GetKernNorm() == float(G["kernnorm_sph"])
            np.testing.assert_allclose(t.CalcDensity(k), G["rho_sph_np"], rtol=RTOL_RHO)


@pytest.mark.parametrize("tag", ["np", "p"])
def test_golden_fof_and_ball(nb, G, tag):
    period = None if tag == "np" else np.ones(3)
    ll = float(G["ll"])
    with nb.KDTree(G["pos"], G["vel"], G["mass"], Period=period) as t:
        g, ng = t.FOF(ll, 5, 0)
        assert ng == G["fof_" + tag].max() and np.array_equal(canon(g), canon(G["fof_" + tag]))
        g, ng, plen = t.FOF(ll, 5, 1, want_len=True)
        assert np.array_equal(canon(g), canon(G["fof_ord_" + tag]))
        assert np.array_equal(np.bincount(g)[1:], np.bincount(G["fof_ord_" + tag])[1:])
        assert np.array_equal(plen[1:], np.bincount(g)[1:]) and np.all(np.diff(plen[1:]) <= 0)
        g, ng = t.FOFCriterion(nb.FOF6D, G["params"], 5, 0)
        assert ng > 0 and np.array_equal(canon(g), canon(G["fof6d_" + tag]))
        g, ng = t.FOFCriterion(nb.FOF3D, G["params"], 5, 0)
        assert np.array_equal(canon(g), canon(G["fof3d_" + tag]))
        off, idx = t.SearchBallPosTaggedPoints(G["xq"], (3 * ll) ** 2, ids=True)
        assert np.array_equal(off, G["ball_off_" + tag])
        assert np.array_equal(np.concatenate(csr_rows_sorted(off, idx)), G["ball_idx_" + tag])


@pytest.mark.parametrize("n,period,k,flags", [(20011, None, 32, 0), (20011, 1, 64, 0), (4099, 1, 17, 1 << 4), (31, None, 8, 0), (16, None, 4, 0), (17, 1, 5, 0)])
def test_port_parity_seeded(nb, port, n, period, k, flags):
    """ragged sizes (non power of two, single leaf, one split), fp32 and forced fp64 storage"""
    from nbodylib_b200.synth import clustered_small
    pos, vel, mass = clustered_small(n, seed=n)
    period = None if period is None else np.ones(3)
    with nb.KDTree(pos, vel, mass, Period=period, flags=flags) as t:
        order = t.order()
        assert np.array_equal(np.sort(order), np.arange(n))
        nn, d2 = t.FindNearestPos(k, ids=True)
        oi, od = port.knn_particles(pos, k, period=period, which=0)
        assert np.array_equal(by_id(order, d2), od) and rows_equal_as_sets(by_id(order, nn), oi)
        if period is not None:
            nn, d2 = t.FindNearestPos(k, ids=True, strict=True)
            oi, od = port.knn_particles(pos, k, period=period, which=0, strict=1)
            assert np.array_equal(by_id(order, d2), od)
        rho = t.CalcDensity(k)
        orho, oh = port.density(pos, mass, k)
        np.testing.assert_allclose(rho, orho, rtol=RTOL_RHO)
        kv = max(1, k // 2)
        np.testing.assert_allclose(t.CalcVelDensity(kv, k), port.veldensity(pos, vel, kv, k), rtol=RTOL_RHO)
        ll = 0.25 / n ** (1.0 / 3)
        for order_flag in (0, 1):
            g, ng = t.FOF(ll, 3, order_flag)
            og, ong = port.fof(pos, None, 0, [ll * ll], period, 3, order_flag)
            assert ng == ong and np.array_equal(canon(g), canon(og))


def test_empty_and_bad_arguments(nb):
    pos = np.random.default_rng(0).random((100, 3))
    with pytest.raises(nb.NbkError):
        nb.KDTree(pos[:0])
    with pytest.raises(nb.NbkError):
        nb.KDTree(pos, TreeType=7)                       # reference: "Error in type of tree specified"
    with pytest.raises(nb.NbkError):
        nb.KDTree(pos, TreeType=nb.TPROJ)                # valid reference call, no device implementation
    with nb.KDTree(pos) as t:
        with pytest.raises(nb.NbkError):
            t.CalcDensity(100)                           # needs Nsmooth < numparts
        with pytest.raises(nb.NbkError):
            t.CalcVelDensity(4, 8)                       # no velocities given
        with pytest.raises(nb.NbkError):
            t.FOFCriterion(nb.FOFVEL, np.zeros(10))
        nn, d2 = t.FindNearestPos(8, q0=10, q1=10)       # empty query range
        assert nn.shape == (0, 8)
        off, idx = t.SearchBallPosTaggedPoints(np.zeros((0, 3)), 0.1)
        assert len(idx) == 0
        # fewer than k candidates: reference fills with -1 / MAXVALUE (KDFindNearest.cxx:16-19)
        nn, d2 = t.FindNearestPos(120, ids=True)
        assert np.all(nn[:, 99:] == -1) and np.all(d2[:, 99:] == 1e32) and np.all(nn[:, :99] >= 0)


def test_duplicates_and_ties(nb, port):
    """coincident particles are never neighbours in the target form (KDLeafNode.cxx:21: dist2>0); FOF links them"""
    rng = np.random.default_rng(4)
    pos = rng.random((3000, 3)).astype(np.float32).astype(np.float64)
    pos[100:200] = pos[0:100]                            # 100 exact duplicates
    with nb.KDTree(pos) as t:
        order = t.order()
        nn, d2 = t.FindNearestPos(8, ids=True)
        oi, od = port.knn_particles(pos, 8)
        assert np.array_equal(by_id(order, d2), od) and np.all(d2 > 0)
        g, ng = t.FOF(1e-6, 2, 0)
        og, ong = port.fof(pos, None, 0, [1e-12], None, 2, 0)
        assert ng == ong == 100 and np.array_equal(canon(g), canon(og))


def test_fp32_key_ties_fall_back_to_exact_heap(nb, port):
    """The density kernels rank candidates by the fp32 rounding of the exact fp64 d2; queries whose k-th and
    (k+1)-th keys coincide must be re-run by the exact fp64 heap.  Quadruplets of points 1e-13 apart force that."""
    rng = np.random.default_rng(12)
    base = rng.random((1500, 3))
    pos = np.concatenate([base + np.array([j * 1e-13, 0, 0]) for j in range(4)])
    mass = 1.0 + rng.random(len(pos))
    vel = rng.standard_normal((len(pos), 3))
    with nb.KDTree(pos, vel, mass) as t:
        assert t.info.store_bytes == 8
        rho, h = t.CalcDensity(18, want_h=True)
        assert t.info.last_flagged > 1000
        orho, oh = port.density(pos, mass, 18)
        np.testing.assert_allclose(rho, orho, rtol=RTOL_RHO)
        assert np.array_equal(h, oh)
        np.testing.assert_allclose(t.CalcVelDensity(7, 18), port.veldensity(pos, vel, 7, 18), rtol=RTOL_RHO)


@pytest.mark.parametrize("opt", [{"knn_leaf": 16}, {"knn_leaf": 64}, {"knn_exact": 1}, {}, {"knn_transpose": 0}, {"knn_transpose": 32}])
def test_density_kernel_variants(nb, port, opt):
    """Every configuration of the density kernel gives the oracle's answer: unmerged 16-particle leaves, two-tile leaves, the
    exact fp64 heap only, the defaults, tiles always / never screened in the transposed form.  Both storage widths."""
    from nbodylib_b200.synth import clustered_small
    n, k = 30011, 40
    pos, vel, mass = clustered_small(n, seed=77)
    orho, oh = port.density(pos, mass, k)
    ovd = port.veldensity(pos, vel, 13, k)
    try:
        for name, v in opt.items():
            nb.set_option(name, v)
        for flags in (0, 1 << 4):
            with nb.KDTree(pos, vel, mass, flags=flags) as t:
                rho, h = t.CalcDensity(k, want_h=True)
                np.testing.assert_allclose(rho, orho, rtol=RTOL_RHO)
                assert np.array_equal(h, oh)
                np.testing.assert_allclose(t.CalcVelDensity(13, k), ovd, rtol=RTOL_RHO)
                np.testing.assert_allclose(t.CalcVelDensity(k, k), port.veldensity(pos, vel, k, k), rtol=RTOL_RHO)
    finally:
        for name in opt:
            nb.set_option(name, -1 if name == "knn_transpose" else 0)
    with pytest.raises(nb.NbkError):
        nb.set_option("no_such_option", 1)


def test_density_kernel_degenerate_inputs(nb, port):
    """Inputs that defeat the fp32 keys of the density kernel go to the exact kernel and still match the oracle: many exactly
    equal distances (a lattice), distances that underflow fp32 (coordinates 1e-25 apart), fewer candidates than k because of
    coincident particles."""
    g = np.arange(12, dtype=np.float64) / 16.0
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(3)
    tiny = (rng.random((700, 3)) * 1e-25).astype(np.float32).astype(np.float64)
    dup = np.repeat(rng.random((40, 3)).astype(np.float32).astype(np.float64), 30, axis=0)
    for pos, k in ((lat, 20), (tiny, 12), (dup, 45)):
        mass = 1.0 + rng.random(len(pos))
        orho, oh = port.density(pos, mass, k)
        with nb.KDTree(pos, None, mass) as t:
            rho, h = t.CalcDensity(k, want_h=True)
            assert np.array_equal(h, oh)
            np.testing.assert_allclose(rho, orho, rtol=RTOL_RHO)


def test_tphs_form_a_equals_fof6d_form_b(nb, port):
    """BASELINE config 4: ScalePhase + TPHS tree + FOF(1.0) == FOFCriterion(FOF6d); scaled coordinates are not
    fp32-representable, so the tree must keep fp64 coordinates to stay bit-exact."""
    from nbodylib_b200.synth import clustered_small
    n = 9000
    pos, vel, mass = clustered_small(n, seed=21)
    ll = 0.3 / n ** (1.0 / 3)
    sv = 0.5 * np.sqrt(((vel - vel.mean(0)) ** 2).sum(1).mean() / 3)
    for period in (None, np.ones(3)):
        ps, vs = pos * (1.0 / ll), vel * (1.0 / sv)
        pA = None if period is None else period * (1.0 / ll)
        with nb.KDTree(ps, vs, mass, TreeType=nb.TPHS, Period=pA) as tA:
            assert tA.info.store_bytes == 8
            gA, ngA = tA.FOF(1.0, 8, 0)
        oA, _ = port.fof(ps, vs, 1, [1.0], pA, 8, 0)
        assert np.array_equal(canon(gA), canon(oA))
        params = np.zeros(10)
        params[1] = params[6] = ll * ll
        params[2] = params[7] = sv * sv
        with nb.KDTree(pos, vel, mass, Period=period) as tB:
            gB, ngB = tB.FOFCriterion(nb.FOF6D, params, 8, 0)
        oB, _ = port.fof(pos, vel, 4, params, period, 8, 0)
        assert ngB > 3 and np.array_equal(canon(gB), canon(oB))
        assert np.array_equal(canon(gA), canon(gB))


def test_config1_against_reference_library(nb):
    """BASELINE config 1 (1M uniform, periodic unit box, b=16, k=32 + CalcDensity) against the reference itself."""
    from oracle.pyoracle import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libnbref.so did not travel with this checkout")
    from nbodylib_b200.synth import uniform_box
    n = 1000000
    pos, vel, mass = uniform_box(n)
    period = np.ones(3)
    R = Ref(pos, vel, mass, period=period)
    with nb.KDTree(pos, vel, mass, Period=period) as t:
        assert (t.GetNumNodes(), t.GetNumLeafNodes()) == (R.num_nodes, R.num_leaves) == (131071, 65536)
        order = t.order()
        nn, d2 = t.FindNearestPos(32, ids=True)
        ri, rd = R.knn_particles(32, which=0)
        assert np.array_equal(by_id(order, d2), rd) and rows_equal_as_sets(by_id(order, nn), ri)
        del nn, d2, ri, rd
        np.testing.assert_allclose(t.CalcDensity(32), R.calc_density(32), rtol=RTOL_RHO)
        ll = 0.2 / n ** (1.0 / 3)
        g, ng = t.FOF(ll, 2, 1)
        rg, rng_ = R.fof(ll, 2, 1)
        assert ng == rng_ and np.array_equal(canon(g), canon(rg))
    R.close()


def test_properties_at_scale(nb):
    """256^3 clustered periodic box (BASELINE config 2 size): properties that need no CPU checker."""
    import torch
    from nbodylib_b200.synth import clustered_box
    ng = 256
    n = ng ** 3
    pos, vel, mass = clustered_box(ng, seed=2024, nhalo=4096, device="cuda")
    with nb.KDTree(pos, vel, mass, Period=np.ones(3)) as t:
        i = t.info
        assert i.store_bytes == 4 and i.num_leaves == 2 ** 20 and i.num_nodes == 2 ** 21 - 1
        order = torch.from_numpy(t.order()).cuda().long()
        assert torch.equal(torch.sort(order).values, torch.arange(n, device="cuda"))
        # tree order really is the permuted input
        dev_pos = pos[order]
        # kNN on a slice: ascending rows, k-th distance consistent with the smoothing scale
        q0, q1 = 5000000, 5000000 + 65536
        nn, d2 = t.FindNearestPos(32, q0=q0, q1=q1)
        assert np.all(np.diff(d2, axis=1) >= 0) and np.all(d2[:, 0] == 0) and np.array_equal(nn[:, 0], np.arange(q0, q1))
        # recompute the distances from the returned indices with minimum image in numpy fp64
        P = dev_pos[q0:q1].double().cpu().numpy()
        Q = dev_pos[torch.from_numpy(nn.astype(np.int64)).cuda()].double().cpu().numpy()
        d = P[:, None, :] - Q
        d -= np.round(d)
        assert np.allclose((d ** 2).sum(-1), d2, rtol=1e-12, atol=1e-18)
        # FOF: linking length 0.2 spacings; labels 1..ng by decreasing size; idempotent; every member of a group
        # has a friend inside the group (checked on a sample through the ball search)
        g, ngrp = t.FOF(0.2 / ng, 20, 1)
        g2, ngrp2 = t.FOF(0.2 / ng, 20, 1)
        assert ngrp == ngrp2 and np.array_equal(g, g2) and g.max() == ngrp and g.min() == 0
        sizes = np.bincount(g)[1:]
        assert sizes.min() >= 20 and np.all(np.diff(sizes) <= 0)
        # mass conservation style check of the density estimator: sum(m/rho) ~ volume (loose statistical bound)
        rho = t.CalcDensity(32)
        assert np.all(rho > 0) and np.isfinite(rho).all()
        vol = (1.0 / rho).sum()
        assert 0.5 < vol < 2.0, vol


def test_cxx_shim_program(built, tmp_path):
    """A C++ program written against the reference's NBody::KDTree interface (examples/shim_demo.cxx, modelled on the
    reference's src/tests/test_kdtree.cxx) compiles against nbodylib_b200/shim/KDTree.h, links libnbk.so and runs."""
    import os
    import shutil
    import subprocess
    from tests.util import ROOT
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if gxx is None:
        pytest.skip("no host C++ compiler")
    exe = str(tmp_path / "shim_demo")
    lib = os.path.join(ROOT, "nbodylib_b200")
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-fopenmp", "-I" + os.path.join(lib, "shim"), os.path.join(ROOT, "examples", "shim_demo.cxx"),
                           "-L" + lib, "-lnbk", "-Wl,-rpath," + lib, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "shim demo ok" in out.stdout, out.stdout + out.stderr
    assert "nodes 32767 leaves 16384" in out.stdout
    assert "per-particle loop vs whole system: 0 differences" in out.stdout
    assert "0 rows differ" in out.stdout and "0 mismatches" in out.stdout
    print(out.stdout)


def test_reference_harness_runs_unchanged_on_the_shim(built):
    """The reference's own test program, src/tests/test_kdtree.cxx, compiled UNCHANGED with the reference's src/NBody and
    src/Math headers against nbodylib_b200/shim/KDTree.h (oracle/Makefile `harness`, built where /root/reference exists; the
    binary travels): all five tree types of its TreeTypes() table (Physical, Physical Rdist, Physical Rdist Adaptfac, Velocity,
    Phase) run on the device through oracle/harness_main.cxx, which also checks what the harness only prints (neighbours and
    ball counts against brute force, one box-sized FOF group, order restored by the destructor)."""
    import os
    import subprocess
    from oracle.pyoracle import HARNESS
    if not os.path.exists(HARNESS):
        pytest.skip("oracle/_ref/test_kdtree_shim did not travel with this checkout")
    out = subprocess.run([HARNESS, "20000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "HARNESS OK" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
    for name in ("Physical", "Velocity", "Phase", "Physical Rdist", "Physical Rdist Adaptfac"):
        assert "==== %s tree\n" % name in out.stdout


@pytest.mark.parametrize("bucket,k", [(1, 3), (4, 9), (8, 33), (32, 16), (64, 40), (100, 7)])
def test_bucket_sizes_and_odd_k(nb, port, bucket, k):
    """leaf sizes below and above the 32-wide tile, k that is neither a power of two nor a multiple of four"""
    from nbodylib_b200.synth import clustered_small
    n = 7001
    pos, vel, mass = clustered_small(n, seed=bucket * 100 + k)
    for period in (None, np.ones(3)):
        with nb.KDTree(pos, vel, mass, bucket_size=bucket, Period=period) as t:
            order = t.order()
            nn, d2 = t.FindNearestPos(k, ids=True)
            oi, od = port.knn_particles(pos, k, period=period, which=0)
            assert np.array_equal(by_id(order, d2), od) and rows_equal_as_sets(by_id(order, nn), oi)
            rho, h = t.CalcDensity(k, want_h=True)
            orho, oh = port.density(pos, mass, k)
            np.testing.assert_allclose(rho, orho, rtol=RTOL_RHO)
            assert np.array_equal(h, oh)
            kv = max(1, k - 2)
            np.testing.assert_allclose(t.CalcVelDensity(kv, k), port.veldensity(pos, vel, kv, k), rtol=RTOL_RHO)
            ll = 0.3 / n ** (1.0 / 3)
            g, ng = t.FOF(ll, 4, 1)
            og, ong = port.fof(pos, None, 0, [ll * ll], period, 4, 1)
            assert ng == ong and np.array_equal(canon(g), canon(og))
            x = np.random.default_rng(k).random((257, 3))
            off, idx = t.SearchBallPosTaggedPoints(x, (2.5 * ll) ** 2, ids=True)
            oo, oidx = port.ball_points(pos, x, (2.5 * ll) ** 2, period)
            assert np.array_equal(off, oo) and np.array_equal(np.concatenate(csr_rows_sorted(off, idx) + [np.zeros(0, np.int32)]), oidx)


def test_input_layouts(nb, port):
    """the same particles through every input path of nbk_create: fp64 / fp32 host arrays, the reference's 88-byte AoS
    Particle records (strided view), device tensors, no masses (NOMASS build), no velocities"""
    import ctypes as C
    import torch
    from nbodylib_b200 import _lib as L
    from nbodylib_b200.synth import clustered_small
    n, k = 5003, 12
    pos, vel, mass = clustered_small(n, seed=3)
    orho, _ = port.density(pos, mass, k)
    orho1, _ = port.density(pos, None, k)

    def rho_of(tree):
        with tree as t:
            return t.CalcDensity(k), t.info
    r, i = rho_of(nb.KDTree(pos, vel, mass))
    np.testing.assert_allclose(r, orho, rtol=RTOL_RHO)
    r, i = rho_of(nb.KDTree(pos.astype(np.float32), vel.astype(np.float32), mass.astype(np.float32)))
    np.testing.assert_allclose(r, orho, rtol=RTOL_RHO)
    assert i.store_bytes == 4
    r, i = rho_of(nb.KDTree(torch.from_numpy(pos).cuda(), torch.from_numpy(vel).cuda(), torch.from_numpy(mass).cuda()))
    np.testing.assert_allclose(r, orho, rtol=RTOL_RHO)
    r, i = rho_of(nb.KDTree(pos, None, None))
    np.testing.assert_allclose(r, orho1, rtol=RTOL_RHO)
    # reference Particle layout (Particle.h:264-354, default build): mass@0 position@8 velocity@32 pid@56 id@60 type@64 rho@72 phi@80
    rec = np.dtype({"names": ["mass", "pos", "vel", "pid", "id", "type", "rho", "phi"],
                    "formats": ["f8", ("f8", 3), ("f8", 3), "i4", "i4", "i4", "f8", "f8"],
                    "offsets": [0, 8, 32, 56, 60, 64, 72, 80], "itemsize": 88})
    parts = np.zeros(n, dtype=rec)
    parts["mass"], parts["pos"], parts["vel"] = mass, pos, vel
    lib = L.load()
    p = L.NbkParticles()
    base = parts.ctypes.data
    p.pos, p.pos_stride = base + 8, 88
    p.vel, p.vel_stride = base + 32, 88
    p.mass, p.mass_stride = base + 0, 88
    p.real_bytes, p.on_device = 8, 0
    h = C.c_void_p()
    L.check(lib.nbk_create(C.byref(p), n, 16, 0, 2, 1000, 0, None, 0, -1, C.byref(h)))
    rho = np.empty(n)
    L.check(lib.nbk_calc_density(h, k, rho.ctypes.data, None, 0))
    lib.nbk_destroy(h)
    np.testing.assert_allclose(rho, orho, rtol=RTOL_RHO)


def test_tvel_tree_knn(nb, port):
    """TVEL trees search in velocity space (KDFindNearest.cxx:336-346) and never reflect (KDSplitNode.cxx:1082-1085)"""
    from nbodylib_b200.synth import clustered_small
    pos, vel, mass = clustered_small(6007, seed=8)
    with nb.KDTree(pos, vel, mass, TreeType=nb.TVEL, Period=np.ones(3)) as t:
        order = t.order()
        nn, d2 = t.FindNearestPos(10, ids=True)
        oi, od = port.knn_particles(vel, 10)
        assert np.array_equal(by_id(order, d2), od) and rows_equal_as_sets(by_id(order, nn), oi)


def test_fof_linked_lists(nb):
    """pHead / pNext / pTail / pLen (KDFOF.cxx:52-67): walking a group's chain visits exactly its members"""
    from nbodylib_b200.synth import clustered_small
    n = 30011
    pos, vel, mass = clustered_small(n, seed=17)
    with nb.KDTree(pos, vel, mass, Period=np.ones(3)) as t:
        order = t.order()
        g, ng, ls = t.FOF(0.25 / n ** (1 / 3), 6, 1, want_lists=True)
        gt = g[order]                                   # group by tree index
        head, nxt, tail, plen = ls["pHead"], ls["pNext"], ls["pTail"], ls["pLen"]
        assert ng > 10 and np.array_equal(plen[1:], np.bincount(g)[1:])
        un = np.nonzero(gt == 0)[0]
        assert np.array_equal(head[un], un) and np.array_equal(tail[un], un) and np.all(nxt[un] == -1)
        for gid in list(range(1, min(ng, 40) + 1)) + [ng]:
            members = np.nonzero(gt == gid)[0]
            i = head[members[0]]
            chain = []
            while i != -1:
                chain.append(i)
                i = nxt[i]
            assert np.array_equal(np.array(chain), members)          # ascending tree index
            assert np.all(head[members] == members[0]) and np.all(tail[members] == members[-1])


@pytest.mark.parametrize("n,bucket,flags", [(1, 16, 0), (15, 16, 0), (4096, 16, 0), (4097, 16, 0), (8193, 1, 0), (50021, 16, 0), (50021, 3, 1 << 4),
                                            (300007, 16, 0), (300007, 100, 1 << 4), (131072, 1024, 0),
                                            # warp-aligned shape (NBK_WARP_ALIGNED = 1 << 7): same build rule, split at a multiple of 32
                                            (33, 16, 1 << 7), (4097, 16, 1 << 7), (4129, 16, 1 << 7), (50021, 16, 1 << 7), (50021, 3, (1 << 7) | (1 << 4)),
                                            (300007, 40, 1 << 7), (1048576 + 7, 16, 1 << 7), (131072, 16, 1 << 7)])
def test_build_structure(nb, n, bucket, flags):
    """The rank-space build (global levels + one shared-memory kernel for nodes <= 4096 particles) obeys the reference's
    build rule node by node (KDTree.cxx:459-504, 994, 1012): ranges from left = ceil(size / 2), leaf iff size <= bucket, cut
    dimension = largest extent (ties -> lowest dimension), tight bounds, every left particle <= every right particle in the
    cut dimension.  Sizes straddle the shared-memory capacity, ties come from a coarse coordinate grid."""
    rng = np.random.default_rng(n + bucket)
    pos = rng.random((n, 3))
    pos[: n // 3] = np.round(pos[: n // 3] * 64) / 64           # many exactly equal coordinates -> ties at the medians
    pos = pos.astype(np.float32).astype(np.float64)
    with nb.KDTree(pos, None, None, bucket_size=bucket, flags=flags) as t:
        order, (s, e, c, b), nn, nl = t.order(), t.nodes(), t.GetNumNodes(), t.GetNumLeafNodes()
    assert np.array_equal(np.sort(order), np.arange(n))
    P = pos[order]
    present = s >= 0
    assert int(present.sum()) == nn
    stack, seen, leaves = [(0, 0, n)], 0, 0
    while stack:
        node, lo, hi = stack.pop()
        assert (s[node], e[node]) == (lo, hi)
        seen += 1
        Q = P[lo:hi]
        assert np.array_equal(b[node, 0::2], Q.min(0).astype(np.float32)) and np.array_equal(b[node, 1::2], Q.max(0).astype(np.float32))
        if hi - lo <= bucket:
            assert c[node] < 0
            leaves += 1
            continue
        ext = Q.max(0) - Q.min(0)
        cd = int(np.argmax(ext))                                # first maximum = lowest dimension on ties
        assert c[node] == cd
        mid = lo + (hi - lo + 1) // 2
        if (flags & (1 << 7)) and hi - lo > 32:                 # include/nbk.h NBK_WARP_ALIGNED: balanced split of the node's 32-particle units
            assert lo % 32 == 0
            mid = lo + 32 * (((hi - lo + 31) // 32 + 1) // 2)
        assert P[lo:mid, cd].max() <= P[mid:hi, cd].min()
        stack.append((2 * node + 1, lo, mid))
        stack.append((2 * node + 2, mid, hi))
    assert (seen, leaves) == (nn, nl)


@pytest.mark.parametrize("n", [1000, 33333, 200003, (1 << 20) + 77])
def test_warp_aligned_tree_gives_the_same_results(nb, n):
    """NBK_WARP_ALIGNED changes the tree's shape (and with it the tree order), never a result: kNN distances and neighbour ids,
    densities, velocity densities, FOF / FOF6d partitions and ball rows equal those of the reference-shaped tree."""
    from nbodylib_b200 import _lib as L
    from nbodylib_b200.synth import clustered_small
    pos, vel, mass = clustered_small(n, seed=n % 97)
    ll = 0.25 / n ** (1 / 3)
    sv2 = ((vel - vel.mean(0)) ** 2).sum(1).mean() / 3.0
    params = np.zeros(10)
    params[1] = params[6] = ll * ll
    params[2] = params[7] = sv2
    res = []
    for flags in (0, L.WARP_ALIGNED):
        for period in (None, np.ones(3)):
            with nb.KDTree(pos, vel, mass, Period=period, flags=flags) as t:
                assert t.info.warp_aligned == (1 if flags else 0)
                order = t.order()
                inv = np.empty(n, dtype=np.int64)
                inv[order] = np.arange(n)
                nn, d2 = t.FindNearestPos(20, ids=True)
                rho, h = t.CalcDensity(48, want_h=True)
                vd = t.CalcVelDensity(32, 48)
                g, ng = t.FOF(ll, 5, 1)
                g6, ng6 = t.FOFCriterion(nb.FOF6D, params, 5, 1)
                off, idx = t.SearchBallPosTagged(np.arange(min(n, 4000), dtype=np.int32), (1.5 * ll) ** 2, ids=True)[:2]
                rows = [np.sort(idx[off[i]:off[i + 1]]) for i in range(len(off) - 1)]
                # per-particle results are by ID; neighbour lists and ball rows are per tree position: bring them to ID order
                res.append(dict(nn=nn[inv], d2=d2[inv], rho=rho, h=h, vd=vd, g=g, ng=ng, g6=g6, ng6=ng6, rows={int(order[i]): r for i, r in enumerate(rows)}))
    for a, b in ((res[0], res[2]), (res[1], res[3])):
        assert np.array_equal(a["d2"], b["d2"]) and np.array_equal(a["h"], b["h"])
        same = (a["nn"] == b["nn"]).all(1)
        for i in np.nonzero(~same)[0]:                      # ties in d2 may be listed in a different order
            assert np.array_equal(np.sort(a["nn"][i]), np.sort(b["nn"][i])) or len(np.unique(a["d2"][i])) < 20
        np.testing.assert_allclose(a["rho"], b["rho"], rtol=1e-12)
        np.testing.assert_allclose(a["vd"], b["vd"], rtol=1e-10)
        assert a["ng"] == b["ng"] and np.array_equal(canon(a["g"]), canon(b["g"]))
        assert a["ng6"] == b["ng6"] and np.array_equal(canon(a["g6"]), canon(b["g6"]))
        common = set(a["rows"]) & set(b["rows"])
        assert len(common) > 0 or n > 4000
        for i in common:
            assert np.array_equal(a["rows"][i], b["rows"][i])


@pytest.mark.parametrize("periodic", [False, True])
def test_checked_fof_against_reference(nb, periodic):
    """FOF / FOFCriterion with a FOFcheckfunc mask (ipcheckflag) and FOFCriterionSetBasisForLinks against the live
    reference with the same mask (reference semantics pinned by tests/test_oracle_cpu.py::test_reference_checked_fof_semantics)."""
    from oracle.pyoracle import Ref, have_ref
    from nbodylib_b200.synth import clustered_small
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    n = 40000
    pos, vel, mass = clustered_small(n, seed=33)
    rng = np.random.default_rng(8)
    types = (rng.random(n) < 0.3).astype(np.int32)
    chk = np.where(types != 0, -1, 0).astype(np.int32)
    basis = types == 0
    period = np.ones(3) if periodic else None
    ll = 0.3 / n ** (1 / 3)
    sv2 = float(((vel - vel.mean(0)) ** 2).sum(1).mean() / 3)
    params = np.zeros(10)
    params[1] = params[6] = ll * ll
    params[2] = params[7] = 4.0 * sv2
    R = Ref(pos, vel, mass, period=period)
    R.set_types(types)
    with nb.KDTree(pos, vel, mass, Period=period) as t:
        for minnum, order in ((5, 0), (5, 1)):
            g, ng = t.FOF(ll, minnum, order, precheck=chk)
            rg, rng_ = R.fof_checked(0, 0, np.array([ll] + [0.0] * 9), minnum, order)
            assert ng == rng_ and np.array_equal(canon(g), canon(np.maximum(rg, 0)))
            for crit in (0, 2):
                g, ng = t.FOFCriterion(crit, params, minnum, order, precheck=chk)
                rg, rng_ = R.fof_checked(1, crit, params, minnum, order)
                assert ng == rng_ and np.array_equal(canon(g), canon(np.maximum(rg, 0)))
                if order:
                    assert np.array_equal(np.bincount(g)[1:], np.bincount(np.maximum(rg, 0))[1:])
        for crit in (0, 2):
            for minnum, order in ((1, 0), (6, 0), (6, 1)):
                g, ng = t.FOFCriterionSetBasisForLinks(crit, params, chk, minnum, order)
                rg, rng_ = R.fof_checked(2, crit, params, minnum, order)

                def label_by_basis(grp):      # group -> smallest ID among its basis members; 0 stays 0
                    lab = np.full(grp.max() + 1, n, dtype=np.int64)
                    np.minimum.at(lab, grp[basis], np.nonzero(basis)[0])
                    lab[0] = -1
                    return lab[grp]
                same = label_by_basis(g) == label_by_basis(rg)
                if minnum == 1:
                    # no group is dissolved: the basis particles' partition is identical, and so is the group count
                    assert ng == rng_ and same[basis].all()
                    assert np.array_equal(canon(np.where(basis, g, 0)), canon(np.where(basis, rg, 0)))
                # a checked particle linked by several groups goes to the one discovered first; the two implementations can
                # order two groups differently only when their first members share a leaf (order inside a leaf is free), and
                # with a size threshold such a particle can tip a group over it -- both rare
                assert same.mean() > 0.999 and abs(ng - rng_) <= 2
    R.close()


@pytest.mark.parametrize("flags", [0, 1 << 4])
def test_attached_halo_tree_equals_masked_single_tree(nb, flags):
    """nbk_attach_halo: main particles in one tree, halo particles in a second one; CalcDensity queries the main particles
    against both.  Must equal the single tree over main + halo with the queries masked to the main particles
    (nbk_calc_density_subset), which is itself checked against the oracle by the sharded tests."""
    from nbodylib_b200.synth import clustered_small
    n, k = 60000, 48
    pos, vel, mass = clustered_small(n, seed=91)
    halo_sel = pos[:, 0] > 0.93                       # a slab face worth of "ghosts"
    main_idx, halo_idx = np.nonzero(~halo_sel)[0], np.nonzero(halo_sel)[0]
    allpos = np.concatenate([pos[main_idx], pos[halo_idx]])
    allmass = np.concatenate([mass[main_idx], mass[halo_idx]])
    n1 = len(main_idx)
    active = np.zeros(n, dtype=np.uint8)
    active[:n1] = 1
    rho_ref, h_ref = np.zeros(n), np.zeros(n)
    with nb.KDTree(allpos, None, allmass, flags=flags) as t:
        t.CalcDensitySubset(k, active, rho_ref, h_ref)
    t1 = nb.KDTree(pos[main_idx], None, mass[main_idx], flags=flags)
    t2 = nb.KDTree(pos[halo_idx], None, mass[halo_idx], flags=flags)
    t1.attach_halo(t2)
    assert t1.n == n and t1.n_main == n1
    rho, h = np.zeros(n), np.zeros(n)
    t1.CalcDensityInto(k, rho, h)
    assert np.array_equal(h[:n1], h_ref[:n1])
    np.testing.assert_allclose(rho, rho_ref, rtol=RTOL_RHO, atol=1e-300)
    assert rho[n1:].max() > 0                          # scatter terms did land on halo particles
    with pytest.raises(nb.NbkError):
        t1.FOF(0.01, 8, 0)
    t1.close()


# ---- single-target estimators, criterion search, dense forms, node mirror (a11, a13; 8f rank 2) ---------
@pytest.fixture(scope="module")
def X():
    return load_golden_extra()


def _where(order):
    w = np.empty_like(order)
    w[order] = np.arange(len(order), dtype=order.dtype)
    return w


@pytest.mark.parametrize("tag", ["np", "p"])
def test_golden_single_target_estimators(nb, G, X, tag):
    """CalcDensityParticle / CalcVelDensityParticle / CalcDensityPosition / CalcVelDensityPosition against the reference's
    values (tests/golden/ref_extra.npz).  Same operation order as the reference; pow() is the only library call that may
    round differently, hence 1e-13 instead of bit equality."""
    period = None if tag == "np" else np.ones(3)
    k, kv = int(G["k"]), int(X["kv"])
    with nb.KDTree(G["pos"], G["vel"], X["mass2"], Period=period) as t:
        where = _where(t.order())
        tt = where[X["qsel"]]
        np.testing.assert_allclose(t.CalcDensityParticle(tt, k), X["dens_part_" + tag], rtol=1e-13)
        np.testing.assert_allclose(t.CalcVelDensityParticle(tt, kv, k), X["vdens_part_" + tag], rtol=1e-13)
        np.testing.assert_allclose(t.CalcDensityPosition(G["xq"], k), X["dens_pos_" + tag], rtol=1e-13)
        np.testing.assert_allclose(t.CalcVelDensityPosition(G["xq"], X["vq"], kv, k), X["vdens_pos_" + tag], rtol=1e-13)
        # whole-system form (no list: tree indices 0..n-1, 32 neighbouring queries per warp) and the scalar form
        alln = t.CalcDensityParticle(None, k)
        np.testing.assert_allclose(alln[tt], X["dens_part_" + tag], rtol=1e-13)
        allv = t.CalcVelDensityParticle(None, kv, k)
        np.testing.assert_allclose(allv[tt], X["vdens_part_" + tag], rtol=1e-13)
        one = t.CalcDensityParticle(int(tt[5]), k)
        assert isinstance(one, float) and one == pytest.approx(float(X["dens_part_" + tag][5]), rel=1e-13)
        # the gather-only density is the whole-system CalcVelDensity's single-target twin: identical numbers
        np.testing.assert_allclose(allv[where], t.CalcVelDensity(kv, k), rtol=1e-13)
        assert t.CalcSmoothLocalValue(k, X["slv_dist"], X["slv_weight"]) == pytest.approx(float(X["slv_value"]), rel=1e-14)
        with pytest.raises(nb.NbkError) as e:
            t.CalcDensityParticle(np.array([0, len(where)]), k)
        assert e.value.code == -1


@pytest.mark.parametrize("tag", ["np", "p"])
def test_golden_criterion_search(nb, G, X, tag):
    """SearchCriterionTagged(Int_t tt | Particle&, FOF3d | FOF6d): rows identical to the reference's"""
    period = None if tag == "np" else np.ones(3)
    with nb.KDTree(G["pos"], G["vel"], G["mass"], Period=period) as t:
        order = t.order()
        tt = _where(order)[X["qsel"]]
        for crit, name in ((nb.FOF3D, "c3"), (nb.FOF6D, "c6")):
            off, idx = t.SearchCriterionTagged(tt, crit, G["params"], ids=True)
            assert np.array_equal(off, X["%s_off_%s" % (name, tag)])
            assert np.array_equal(np.concatenate(csr_rows_sorted(off, idx)), X["%s_idx_%s" % (name, tag)])
            off2, idx2 = t.SearchCriterionTagged(tt, crit, G["params"])                 # tree indices map to the same IDs
            assert np.array_equal(off2, off) and np.array_equal(order[idx2], idx)
            off, idx = t.SearchCriterionTaggedPoints(X["xn"], X["vn"], crit, G["params"], ids=True)
            assert np.array_equal(off, X["%sx_off_%s" % (name, tag)])
            assert np.array_equal(np.concatenate(csr_rows_sorted(off, idx)), X["%sx_idx_%s" % (name, tag)])
        with pytest.raises(nb.NbkError) as e:
            t.SearchCriterionTagged(tt[:4], nb.FOFVEL, G["params"])
        assert e.value.code == -3
        with pytest.raises(nb.NbkError) as e:
            t.SearchBallPosTagged([-1], 0.01)
        assert e.value.code == -1


@pytest.mark.parametrize("tag", ["np", "p"])
def test_golden_dense_search_forms(nb, G, X, tag):
    """dense SearchBallPos(tt | x, fdist2, imark, nn, dist2) and SearchCriterion(tt, cmp, params, imark, nn, dist2): the
    caller's N-entry arrays end up as the reference leaves them (the target's own entry aside, quirk Q5)."""
    period = None if tag == "np" else np.ones(3)
    n = len(G["pos"])
    r2 = (2.5 * float(G["ll"])) ** 2
    with nb.KDTree(G["pos"], G["vel"], G["mass"], Period=period) as t:
        where = _where(t.order())
        dq = X["dense_q"]
        nn, d2 = np.zeros(n, dtype=np.int32), np.zeros(n)
        for j, q in enumerate(dq):
            t.SearchBallPos(int(where[q]), r2, j + 1, nn, d2)
        keep = np.ones(n, dtype=bool)
        keep[dq] = False
        assert np.array_equal(nn[keep], X["dense_ball_nn_" + tag][keep]) and np.array_equal(d2[keep], X["dense_ball_d2_" + tag][keep])
        nn[:] = 0
        d2[:] = 0
        for j, x in enumerate(G["xq"][:12]):
            t.SearchBallPos(x, r2, j + 1, nn, d2)
        assert np.array_equal(nn, X["dense_ballx_nn_" + tag]) and np.array_equal(d2, X["dense_ballx_d2_" + tag])
        nn[:] = 0
        d2[:] = 0
        for j, q in enumerate(dq):
            off, idx, dd = t.SearchCriterionTagged([int(where[q])], nb.FOF6D, G["params"], ids=True, want_d2=True)
            take = (nn[idx] > j + 1) | (nn[idx] == 0)                      # KDLeafNode.cxx:418
            nn[idx[take]] = j + 1
            d2[idx[take]] = dd[take]
        assert np.array_equal(nn, X["dense_c6_nn_" + tag])
        if tag == "np":
            assert np.array_equal(d2, X["dense_c6_d2_" + tag])
        nn2 = np.zeros(n, dtype=np.int32)
        t.SearchCriterion(int(where[dq[0]]), nb.FOF6D, G["params"], 3, nn2)
        assert np.array_equal(nn2 == 3, X["dense_c6_nn_" + tag] == 1)


def test_node_mirror_against_reference(nb, G, X):
    """GetRoot() / FindLeafNode() consumers: the host mirror of the node arrays has the reference's split dimensions and
    cut values (depth-first order) and its leaves hold the same particles"""
    with nb.KDTree(G["pos"], G["vel"], G["mass"]) as t:
        order = t.order()
        where = _where(order)
        s, e, c, b = t.nodes()
        dims, vals = [], []
        stack = [0]
        while stack:
            slot = stack.pop()
            if c[slot] < 0:
                continue
            dims.append(int(c[slot]))
            vals.append(float(b[2 * slot + 1, 2 * c[slot] + 1]))
            stack.append(2 * slot + 2)
            stack.append(2 * slot + 1)
        assert np.array_equal(dims, X["cut_dims"]) and np.array_equal(vals, X["cut_vals"])
        off, ids = X["leaf_off"], X["leaf_ids"]
        for j, q in enumerate(X["dense_q"]):
            _, a, z = t.FindLeafNode(int(where[q]))
            assert np.array_equal(np.sort(order[a:z]), ids[off[j]:off[j + 1]])
        off, ids = X["leafx_off"], X["leafx_ids"]
        for j, x in enumerate(G["xq"][:12]):
            _, a, z = t.FindLeafNode(x)
            assert np.array_equal(np.sort(order[a:z]), ids[off[j]:off[j + 1]])


@pytest.mark.parametrize("tag", ["np", "p"])
def test_golden_filtered_knn(nb, G, X, tag):
    """FindNearestCheck (tt | Coordinate) and FindNearestCriterion (tt | Particle; FOF3d, FOF6d) against the reference: same
    d2 rows bit for bit, same neighbour sets, (-1, 1e32) padding where fewer than k particles qualify.  Periodic trees
    reproduce the reference's drop-the-nearest behaviour by default; tree_form=False returns the k nearest instead."""
    period = None if tag == "np" else np.ones(3)
    kf = int(X["kf"])
    with nb.KDTree(G["pos"], G["vel"], G["mass"], Period=period) as t:
        order = t.order()
        nn, d2 = t.FindNearestCheck(X["types"], kf, ids=True)
        assert np.array_equal(by_id(order, d2), X["nnchk_d2_" + tag]) and rows_equal_as_sets(by_id(order, nn), X["nnchk_ids_" + tag])
        assert np.all(X["types"][nn] == 0)
        nn, d2 = t.FindNearestCheck(X["types"], kf, x=G["xq"], ids=True)
        assert np.array_equal(d2, X["nnchkx_d2_" + tag]) and rows_equal_as_sets(nn, X["nnchkx_ids_" + tag])
        for crit, name in ((nb.FOF3D, "c3"), (nb.FOF6D, "c6")):
            nn, d2 = t.FindNearestCriterion(crit, G["params"], kf, ids=True)
            assert np.array_equal(by_id(order, d2), X["nn%s_d2_%s" % (name, tag)])
            assert rows_equal_as_sets(by_id(order, nn), X["nn%s_ids_%s" % (name, tag)])
            nn, d2 = t.FindNearestCriterion(crit, G["params"], kf, x=X["xn"], v=X["vn"], ids=True)
            assert np.array_equal(d2, X["nn%sx_d2_%s" % (name, tag)]) and rows_equal_as_sets(nn, X["nn%sx_ids_%s" % (name, tag)])
        if tag == "p":
            # without the reference's off-by-one the rows start one neighbour earlier
            nn1, d21 = t.FindNearestCheck(X["types"], kf + 1, ids=True, tree_form=False)
            assert np.array_equal(by_id(order, d21)[:, 1:], X["nnchk_d2_p"])
        with pytest.raises(nb.NbkError) as e:
            t.FindNearestCriterion(nb.FOFVEL, G["params"], kf)
        assert e.value.code == -3


def test_golden_smoothed_velocity_moments(nb, G, X):
    """CalcSmoothVel / CalcSmoothVelDisp against the reference (fp64 atomics: only the summation order differs)"""
    k = int(G["k"])
    with nb.KDTree(G["pos"], G["vel"], X["mass2"]) as t:
        rho = t.CalcDensity(k)
        np.testing.assert_allclose(rho, X["sm_rho"], rtol=RTOL_RHO)
        sv = t.CalcSmoothVel(k, rho=X["sm_rho"])
        np.testing.assert_allclose(sv, X["sm_vel"], rtol=0, atol=1e-11 * np.abs(X["sm_vel"]).max())
        sd = t.CalcSmoothVelDisp(X["sm_vel"], k, rho=X["sm_rho"])
        np.testing.assert_allclose(sd, X["sm_disp"], rtol=0, atol=1e-11 * np.abs(X["sm_disp"]).max())
        np.testing.assert_allclose(sd, np.transpose(sd, (0, 2, 1)), rtol=0, atol=1e-12 * np.abs(sd).max())
        sk = t.CalcSmoothVelSkew(X["sm_vel"], X["sm_disp"], k, rho=X["sm_rho"])
        np.testing.assert_allclose(sk, X["sm_skew"], rtol=0, atol=1e-10 * np.abs(X["sm_skew"]).max())
        ku = t.CalcSmoothVelKurtosis(X["sm_vel"], X["sm_disp"], k, rho=X["sm_rho"])
        np.testing.assert_allclose(ku, X["sm_kurt"], rtol=0, atol=1e-10 * np.abs(X["sm_kurt"]).max())
        # densityset != 1: the density is computed first
        sv2 = t.CalcSmoothVel(k)
        np.testing.assert_allclose(sv2, X["sm_vel"], rtol=0, atol=1e-9 * np.abs(X["sm_vel"]).max())
