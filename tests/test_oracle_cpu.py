"""CPU suite (-m "not gpu"): pins the oracle port against the golden vectors produced by the reference, checks
the port against the live reference build when oracle/_ref exists, the host-side logic, and that the C-ABI
library loads and exports every symbol include/nbk.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.util import ROOT, canon, csr_rows_sorted, load_golden, load_golden_phase, rows_equal_as_sets


@pytest.fixture(scope="module")
def G():
    return load_golden()


def test_known_kernel_norms(port, G):
    # SURVEY.md 8c known answers
    norm, tab = port.kernel_table(3, 0, 1000)
    assert norm == pytest.approx(0.31830988618379069, rel=0, abs=1e-16)
    norm_e, tab_e = port.kernel_table(3, 2, 1000)
    assert norm_e == pytest.approx(0.074603879574325946, rel=0, abs=1e-16)
    assert norm == float(G["kernnorm_sph"]) and norm_e == float(G["kernnorm_epan"])
    assert tab_e[-1] == 0.0 and tab[-1] == 0.0 and tab_e[0] == norm_e


@pytest.mark.parametrize("tag", ["np", "p"])
def test_port_knn_matches_reference_golden(port, G, tag):
    period = None if tag == "np" else np.ones(3)
    k = int(G["k"])
    for which in (0, 1):
        ids, d2 = port.knn_particles(G["pos"], k, period=period, which=which)
        assert np.array_equal(d2, G["knn%d_d2_%s" % (which, tag)])          # bit-exact fp64
        assert rows_equal_as_sets(ids, G["knn%d_ids_%s" % (which, tag)])
    ids, d2 = port.knn_points(G["pos"], G["xq"], k, period=period)
    assert np.array_equal(d2, G["knnx_d2_" + tag]) and rows_equal_as_sets(ids, G["knnx_ids_" + tag])


def test_periodic_self_rules_q3(G):
    # quirk Q3: periodic FindNearestPos(tt) returns self first; FindNearest(tt) never returns self
    ids0, ids1 = G["knn0_ids_p"], G["knn1_ids_p"]
    n = len(ids0)
    assert np.array_equal(ids0[:, 0], np.arange(n)) and np.all(G["knn0_d2_p"][:, 0] == 0)
    assert not np.any(ids1 == np.arange(n)[:, None])
    assert not np.any(G["knn0_ids_np"] == np.arange(n)[:, None])


@pytest.mark.parametrize("tag", ["np", "p"])
def test_port_density_matches_reference_golden(port, G, tag):
    k = int(G["k"])
    rho, h = port.density(G["pos"], G["mass"], k)
    np.testing.assert_allclose(rho, G["rho_" + tag], rtol=1e-13)
    # quirk Q2: the period is ignored by CalcDensity
    np.testing.assert_allclose(G["rho_p"], G["rho_np"], rtol=1e-13)
    vr = port.veldensity(G["pos"], G["vel"], 8, k)
    np.testing.assert_allclose(vr, G["vrho_" + tag], rtol=1e-13)
    if tag == "np":
        rho_s, _ = port.density(G["pos"], G["mass"], k, kerntype=0)
        np.testing.assert_allclose(rho_s, G["rho_sph_np"], rtol=1e-13)


@pytest.mark.parametrize("tag", ["np", "p"])
def test_port_fof_matches_reference_golden(port, G, tag):
    period = None if tag == "np" else np.ones(3)
    ll = float(G["ll"])
    g, ng = port.fof(G["pos"], None, 0, [ll * ll], period, 5, 0)
    assert ng == G["fof_" + tag].max() and np.array_equal(canon(g), canon(G["fof_" + tag]))
    g, ng = port.fof(G["pos"], None, 0, [ll * ll], period, 5, 1)
    assert np.array_equal(canon(g), canon(G["fof_ord_" + tag]))
    assert np.array_equal(np.bincount(g)[1:], np.bincount(G["fof_ord_" + tag])[1:])   # same size ranking
    g, ng = port.fof(G["pos"], G["vel"], 4, G["params"], period, 5, 0)
    assert ng > 0 and np.array_equal(canon(g), canon(G["fof6d_" + tag]))
    g, ng = port.fof(G["pos"], G["vel"], 2, G["params"], period, 5, 0)
    assert np.array_equal(canon(g), canon(G["fof3d_" + tag]))


@pytest.mark.parametrize("tag", ["np", "p"])
def test_port_ball_matches_reference_golden(port, G, tag):
    period = None if tag == "np" else np.ones(3)
    off, idx = port.ball_points(G["pos"], G["xq"], (3 * float(G["ll"])) ** 2, period)
    assert np.array_equal(off, G["ball_off_" + tag])
    assert np.array_equal(np.concatenate(csr_rows_sorted(off, idx)), G["ball_idx_" + tag])


@pytest.mark.parametrize("tag", ["np", "p"])
def test_port_phase_knn_matches_reference_golden(port, G, tag):
    """FindNearestPhase(tt) / FindNearestPhase(x, v) (KDFindNearest.cxx:347-361,543-555): tests/golden/ref_phase.npz holds the
    reference's rows on a TPHS tree built with Aniso = -1"""
    H = load_golden_phase()
    period = None if tag == "np" else np.ones(3)
    k = int(H["k"])
    ids, d2 = port.knn_phase_particles(G["pos"], H["vel"], H["qsel"], k, period=period)
    assert np.array_equal(d2, H["phase_d2_" + tag]) and rows_equal_as_sets(ids, H["phase_ids_" + tag])
    assert not np.any(ids == H["qsel"][:, None])                     # the particle itself is never returned, periodic or not
    ids, d2 = port.knn_phase_points(G["pos"], H["vel"], G["xq"], H["vq"], k, period=period)
    assert np.array_equal(d2, H["phasex_d2_" + tag]) and rows_equal_as_sets(ids, H["phasex_ids_" + tag])
    # the 6D key is the position part plus the velocity part: never below the position distance to the same particle
    dx = G["pos"][ids] - G["xq"][:, None, :]
    if period is not None:
        dx -= np.round(dx)
    assert np.all(d2 >= (dx ** 2).sum(-1) * (1 - 1e-12))


def test_phase_port_vs_live_reference(port):
    """the phase-space port against the reference itself on a fresh seed: TPHS tree (Aniso = -1, both FindNearestPhase(tt) and
    FindNearest(tt)) and TPHYS tree (FindNearestPhase walks whatever tree it is called on), periodic and not"""
    from oracle.pyoracle import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built on this machine")
    from nbodylib_b200.synth import clustered_small
    pos, vel, mass = clustered_small(5000, seed=7)
    vel = (vel * (0.1 / vel.std())).astype(np.float32).astype(np.float64)
    rng = np.random.default_rng(3)
    qs = rng.permutation(5000)[:1500].astype(np.int32)
    xq = rng.random((200, 3))
    vq = rng.normal(size=(200, 3)) * 0.1
    for period in (None, np.ones(3)):
        pi, pd = port.knn_phase_particles(pos, vel, qs, 20, period=period)
        px, pxd = port.knn_phase_points(pos, vel, xq, vq, 20, period=period)
        for tt in (Ref.TPHS, Ref.TPHYS):
            R = Ref(pos, vel, mass, treetype=tt, period=period, aniso=-1)
            for which in ((0, 1) if tt == Ref.TPHS else (0,)):
                ids, d2 = R.knn_phase_particles(qs, 20, which=which)
                assert np.array_equal(d2, pd) and rows_equal_as_sets(ids, pi)
            ids, d2 = R.knn_phase_points(xq, vq, 20)
            assert np.array_equal(d2, pxd) and rows_equal_as_sets(ids, px)
            R.close()


def test_velocity_tree_searches_of_the_live_reference(port):
    """what the device path mirrors on TVEL trees: FindNearestVel(tt | v) returns the k nearest in velocity space, never reflected,
    whether or not the tree has a period; FindNearest(tt) equals it on a non-periodic tree and, on a PERIODIC velocity tree, drops
    the nearest neighbour (a latent bug that is documented, not reproduced)"""
    from oracle.pyoracle import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built on this machine")
    from nbodylib_b200.synth import clustered_small
    pos, vel, mass = clustered_small(6007, seed=31)
    vq = np.random.default_rng(9).normal(size=(129, 3)) * vel.std()
    qs = np.arange(0, 6007, 3, dtype=np.int32)
    oi, od = port.knn_particles(vel, 11)
    ox, oxd = port.knn_points(vel, vq, 10)
    for period in (None, np.ones(3)):
        R = Ref(pos, vel, mass, treetype=Ref.TVEL, period=period)
        ids, d2 = R.knn_particle_list(qs, 10, which=2)
        assert np.array_equal(d2, od[qs, :10]) and rows_equal_as_sets(ids, oi[qs, :10])
        ids, d2 = R.knn_vel_points(vq, 10)
        assert np.array_equal(d2, oxd) and rows_equal_as_sets(ids, ox)
        ids, d2 = R.knn_particle_list(qs, 10, which=1)
        if period is None:
            assert np.array_equal(d2, od[qs, :10])
        else:
            assert np.array_equal(d2, od[qs, 1:11])              # neighbours 2 .. k+1
        R.close()


def test_port_vs_live_reference(port):
    """Where the reference compiles (oracle/_ref), re-check the port on a fresh seed incl. tree shape facts."""
    from oracle.pyoracle import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built on this machine")
    from nbodylib_b200.synth import clustered_small
    pos, vel, mass = clustered_small(6000, seed=99)
    for period in (None, np.ones(3)):
        R = Ref(pos, vel, mass, period=period)
        assert (R.num_nodes, R.num_leaves) == (1023, 512)
        ids, d2 = R.knn_particles(24, which=0)
        pi, pd = port.knn_particles(pos, 24, period=period, which=0)
        assert np.array_equal(d2, pd) and rows_equal_as_sets(ids, pi)
        ll = 0.3 / 6000 ** (1 / 3)
        g, ng = R.fof(ll, 6, 0)
        pg, png = port.fof(pos, None, 0, [ll * ll], period, 6, 0)
        assert ng == png and np.array_equal(canon(g), canon(pg))
        R.close()


def test_full_host_density_loop_equals_library_calcdensity():
    """The full-host OpenMP kNN-density loop that bench.py times and the scale-parity tests check against (oracle/ref_driver.cxx
    ref_calc_density_omp: the caller-side omp loop over FindNearestPos of src/tests/test_kdtree.cxx:279-301 + the accumulation
    of KDCalcSmoothQuantities.cxx:260-300) gives the library's serial CalcDensity up to the summation order; the explicit-list
    kNN entry point equals the range form."""
    from oracle.pyoracle import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built on this machine")
    from nbodylib_b200.synth import clustered_small
    pos, vel, mass = clustered_small(20000, seed=314)
    R = Ref(pos, vel, mass, period=None)
    Ref.set_threads(4)
    a = R.calc_density_omp(32)
    b = R.calc_density(32)
    np.testing.assert_allclose(a, b, rtol=1e-12)
    q = np.random.default_rng(0).permutation(20000)[:500].astype(np.int32)
    ids, d2 = R.knn_particle_list(q, 32)
    ids_all, d2_all = R.knn_particles(32)
    assert np.array_equal(ids, ids_all[q]) and np.array_equal(d2, d2_all[q])
    R.close()


def test_reference_harness_compiles_unchanged_against_the_shim(built):
    """Where the reference sources are present: src/tests/test_kdtree.cxx compiles unchanged with the reference's NBody / Math
    headers and the shim's KDTree.h first on the include path (NBK_USE_REFERENCE_PARTICLE), and links against libnbk.so.
    (It needs a GPU to run: tests/test_gpu_parity.py::test_reference_harness_runs_unchanged_on_the_shim.)"""
    import os
    from oracle import pyoracle
    if not os.path.isdir("/root/reference/src/tests"):
        pytest.skip("reference sources not present")
    exe = pyoracle.build_harness()
    assert exe is not None and os.access(exe, os.X_OK)


def test_fofvel_in_the_reference_is_not_a_pair_relation():
    """Why FOFVel (FOFFunc.h:39-46) has no device implementation.  The criterion constrains velocities only,
    while the reference's walk prunes in position space (params[1]) and applies the criterion to EVERY particle of every leaf it
    opens (KDSplitNode.cxx:991-1028, KDLeafNode.cxx:591-619).  On the live reference: (a) most of the particles
    SearchCriterionTagged(FOFVel) returns lie beyond the position radius -- which ones depends on the leaf boxes, not on the
    pair; (b) the returned relation is not symmetric (j is linked from i but i not from j), so the groups FOFCriterion builds
    from it depend on the order in which its breadth-first search discovers particles; (c) the partition it returns is
    neither the components of the symmetric pair relation (velocity criterion AND position radius) nor those of the
    symmetrised walk relation.  There is no order-independent definition to be bit-exact against."""
    from oracle.pyoracle import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built on this machine")
    from nbodylib_b200.synth import clustered_small
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    n = 4000
    pos, vel, mass = clustered_small(n, seed=31)
    R = Ref(pos, vel, mass, period=None)
    ll = 0.4 / n ** (1 / 3)
    sv2 = ((vel - vel.mean(0)) ** 2).sum(1).mean() / 3
    params = np.zeros(10)
    params[1] = ll * ll
    params[2] = params[6] = params[7] = 0.05 * sv2
    off, ids = R.search_criterion_particles(np.arange(n, dtype=np.int32), 1, params)
    src = np.repeat(np.arange(n), np.diff(off))
    walk = np.zeros((n, n), dtype=bool)
    walk[src, ids] = True
    np.fill_diagonal(walk, False)
    dx2 = ((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1)
    dv2 = ((vel[:, None, :] - vel[None, :, :]) ** 2).sum(-1)
    pair = (dv2 / params[6] < 1) & (dx2 < params[1])
    np.fill_diagonal(pair, False)
    assert (walk & ~(dx2 < params[1])).sum() > 1000                 # (a) links beyond the position radius
    assert (walk & pair).sum() == pair.sum()                          #     (every true pair is found as well)
    assert (walk & ~walk.T).sum() > 100                               # (b) asymmetric links
    g, ng = R.fof_criterion(1, params, 2, 0)

    def components(adj):
        i, j = np.nonzero(adj)
        _, comp = connected_components(coo_matrix((np.ones(len(i), dtype=np.int8), (i, j)), shape=(n, n)), directed=False)
        sizes = np.bincount(comp)
        lab = np.where(sizes[comp] >= 2, comp + 1, 0)
        return lab
    assert not np.array_equal(canon(g), canon(components(pair)))      # (c)
    assert not np.array_equal(canon(g), canon(components(walk | walk.T)))
    R.close()


def test_reference_tree_shape_known_answer():
    """N=1e6, b=16 -> 131071 nodes / 65536 leaves: the closed-form shape used by the device build."""
    def shape(n, b):
        nodes = leaves = 0
        level = {n: 1}
        while level:
            nxt = {}
            for s, c in level.items():
                nodes += c
                if s <= b:
                    leaves += c
                else:
                    for ch in ((s + 1) // 2, s // 2):
                        nxt[ch] = nxt.get(ch, 0) + c
            level = nxt
        return nodes, leaves
    assert shape(1000000, 16) == (131071, 65536)
    assert shape(2500, 16) == (511, 256)
    assert shape(13, 16) == (1, 1)


def test_cabi_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "nbk.h")).read()
    declared = sorted(set(re.findall(r"\b(nbk_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 18
    lib = ctypes.CDLL(os.path.join(ROOT, "nbodylib_b200", "libnbk.so"))
    for name in declared:
        assert hasattr(lib, name), "libnbk.so does not export %s" % name
    from nbodylib_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared
    # the sharded layer's own library (include/nbk_sharded.h): loads without a GPU, exports every declared entry point, and a
    # call without a device fails with a status code instead of computing anything on the host
    hdr = open(os.path.join(ROOT, "include", "nbk_sharded.h")).read()
    declared = sorted(set(re.findall(r"\b(nbk_(?:comm|sharded)_[a-z0-9_]+)\s*\(", hdr)))
    S = _lib.load_sharded()
    for name in declared:
        assert hasattr(S, name), "libnbk_sharded.so does not export %s" % name
    assert sorted(_lib.SHARDED_EXPORTS) == declared
    import torch
    if not torch.cuda.is_available():
        comm = ctypes.c_void_p()
        assert S.nbk_comm_init_rank(1, 0, ctypes.addressof((ctypes.c_ubyte * 128)()), 0, ctypes.byref(comm)) == -2
        assert not comm.value and b"cuda" in _lib.load().nbk_last_error().lower()


def test_no_cpu_fallback_without_device(built):
    """On a machine without a GPU the product must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nbodylib_b200 import KDTree, NbkError
    with pytest.raises(NbkError) as e:
        KDTree(np.random.rand(100, 3))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under nbodylib_b200/ may import, include, link or dlopen it."""
    pkg = os.path.join(ROOT, "nbodylib_b200")
    pat = re.compile(r"(^\s*(from|import)\s+\S*oracle)|(#include\s*[\"<][^\">]*oracle)|(liboracle|libnbref|pyoracle)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".cxx")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), "%s references the oracle" % f


def test_synth_is_fp32_representable_and_deterministic():
    from nbodylib_b200.synth import clustered_box, clustered_small, uniform_box
    p, v, m = clustered_small(3000, seed=5)
    assert np.array_equal(p, p.astype(np.float32).astype(np.float64)) and p.min() >= 0 and p.max() < 1
    p2, _, _ = clustered_small(3000, seed=5)
    assert np.array_equal(p, p2)
    pos, vel, mass = clustered_box(16, seed=3, nhalo=4, min_members=32)
    assert pos.shape == (4096, 3) and float(pos.min()) >= 0 and float(pos.max()) < 1
    pu, _, _ = uniform_box(1000)
    assert pu.shape == (1000, 3)


def _checked_fof_expectation(pos, types, p6, period):
    """numpy restatement of the checked FOF entry points on a small set: link matrix of FOF3d (FOFFunc.h:30-35),
    components of the check == 0 ("basis") particles, and for every other particle the basis groups that link it."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components
    d = pos[:, None, :] - pos[None, :, :]
    if period is not None:
        d = d - np.round(d / period) * period
    link = ((d[..., 0] ** 2) / p6 + (d[..., 1] ** 2) / p6 + (d[..., 2] ** 2) / p6) < 1
    np.fill_diagonal(link, False)
    basis = types == 0
    bb = link & basis[:, None] & basis[None, :]
    _, comp = connected_components(csr_matrix(bb), directed=False)
    return link, basis, comp


@pytest.mark.parametrize("periodic", [False, True])
def test_reference_checked_fof_semantics(periodic):
    """Pins what the three check-function entry points of the reference compute (live reference, no GPU):
    FOF / FOFCriterion with ipcheckflag drop the checked particles entirely; FOFCriterionSetBasisForLinks groups the
    unchecked ones by their mutual links and hands every checked particle to the EARLIEST-DISCOVERED group linking it.
    These are the semantics nbk_fof(precheck) and nbk_fof_criterion_basis implement on the device."""
    from oracle.pyoracle import Ref, have_ref
    from nbodylib_b200.synth import clustered_small
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    n = 3000
    pos, vel, mass = clustered_small(n, seed=21)
    rng = np.random.default_rng(5)
    types = (rng.random(n) < 0.35).astype(np.int32)
    period = np.ones(3) if periodic else None
    ll = 0.35 / n ** (1 / 3)
    params = np.zeros(10)
    params[1] = params[6] = ll * ll
    link, basis, comp = _checked_fof_expectation(pos, types, params[6], period)
    R = Ref(pos, vel, mass, period=period)
    R.set_types(types)
    # (1) ipcheckflag: checked particles are not there at all
    for which, prm in ((0, np.array([ll] + [0.0] * 9)), (1, params)):
        g, ng = R.fof_checked(which, 0, prm, minnum=2, order=0)
        assert np.all(g[~basis] <= 0)
        sizes = np.bincount(comp[basis])
        want = np.where(sizes[comp] >= 2, comp + 1, 0) * basis
        assert np.array_equal(canon(np.where(basis, g, 0)), canon(want))
    # (2) SetBasisForLinks, minnum = 1 so that no group is dissolved
    g, ng = R.fof_checked(2, 0, params, minnum=1, order=0)
    assert np.array_equal(canon(np.where(basis, g, 0)), canon(np.where(basis, comp + 1, 0)))
    amb = 0
    for i in np.nonzero(~basis)[0]:
        cand = np.unique(g[link[i] & basis])
        if len(cand) == 0:
            assert g[i] == 0
        else:
            assert g[i] == cand.min()        # order = 0 numbers the groups in discovery order
            amb += len(cand) > 1
    assert amb > 0                            # the rule was actually exercised
    R.close()


# ---- extra golden vectors (single-target estimators, criterion search, dense forms, node getters) -------------------
@pytest.fixture(scope="module")
def X():
    from tests.util import load_golden_extra
    return load_golden_extra()


def test_single_target_estimators_restatement(port, G, X):
    """CalcDensityParticle / CalcVelDensityParticle / CalcSmoothLocalValue of the reference == the gather-only restatement
    the CUDA epilogue follows (KDCalcSmoothQuantities.cxx:768-921, 1704-1735), bit for bit; the period is ignored (Q2)."""
    from tests.util import gather_density, gather_veldensity
    k, kv = int(G["k"]), int(X["kv"])
    _, kern = port.kernel_table(3, 2, 1000)
    ids, d2 = G["knn0_ids_np"], G["knn0_d2_np"]
    q = X["qsel"]
    mine = np.array([gather_density(kern, d2[i], X["mass2"][ids[i]]) for i in q])
    assert np.array_equal(mine, X["dens_part_np"])
    assert np.array_equal(X["dens_part_p"], X["dens_part_np"]) and np.array_equal(X["vdens_part_p"], X["vdens_part_np"])
    vmine = np.array([gather_veldensity(kern, G["vel"][i], G["vel"][ids[i]], kv) for i in q])
    np.testing.assert_allclose(vmine, X["vdens_part_np"], rtol=1e-14)
    # position forms: coordinate search (a particle at distance 0 counts), same sums
    idx, d2x = G["knnx_ids_np"], G["knnx_d2_np"]
    pm = np.array([gather_density(kern, d2x[i], X["mass2"][idx[i]]) for i in range(len(idx))])
    assert np.array_equal(pm, X["dens_pos_np"]) and np.array_equal(X["dens_pos_p"], X["dens_pos_np"])
    pv = np.array([gather_veldensity(kern, X["vq"][i], G["vel"][idx[i]], kv) for i in range(len(idx))])
    np.testing.assert_allclose(pv, X["vdens_pos_np"], rtol=1e-14)
    # CalcSmoothLocalValue(Nsmooth, dist, weight): dist descending
    assert gather_density(kern, X["slv_dist"][::-1] ** 2, X["slv_weight"][::-1]) == pytest.approx(float(X["slv_value"]), rel=1e-15)


@pytest.mark.parametrize("tag", ["np", "p"])
def test_criterion_search_restatement(G, X, tag):
    """SearchCriterionTagged of the reference == every particle other than the target meeting the criterion over the
    reference's periodic images (KDFindNearest.cxx:660-706, KDLeafNode.cxx:445-492, KDSplitNode.cxx:1496-1529)."""
    from tests.util import crit_rows
    period = None if tag == "np" else np.ones(3)
    pos, vel, params, q = G["pos"], G["vel"], G["params"], X["qsel"]
    for crit, name in ((0, "c3"), (2, "c6")):
        rows = crit_rows(pos, vel, pos[q], vel[q], crit, params, period, exclude=q)
        off, idx = X["%s_off_%s" % (name, tag)], X["%s_idx_%s" % (name, tag)]
        assert all(np.array_equal(rows[i], idx[off[i]:off[i + 1]]) for i in range(len(q)))
        rows = crit_rows(pos, vel, X["xn"], X["vn"], crit, params, period)
        off, idx = X["%sx_off_%s" % (name, tag)], X["%sx_idx_%s" % (name, tag)]
        assert all(np.array_equal(rows[i], idx[off[i]:off[i + 1]]) for i in range(len(rows)))


@pytest.mark.parametrize("tag", ["np", "p"])
def test_dense_search_restatement(G, X, tag):
    """dense SearchBallPos / SearchCriterion (KDFindNearest.cxx:567-603): nn[ID] = imark, dist2[ID] = position d2; later
    calls overwrite earlier marks in the ball form, the criterion form only claims unmarked or later-marked particles.
    The target itself is marked only when the reference swallows a whole node around it (quirk Q5): masked out here."""
    from tests.util import ball_min_d2, crit_rows
    period = None if tag == "np" else np.ones(3)
    pos, vel, ll = G["pos"], G["vel"], float(G["ll"])
    r2 = (2.5 * ll) ** 2
    n = len(pos)
    for key, queries, targets in (("dense_ball", pos[X["dense_q"]], X["dense_q"]), ("dense_ballx", G["xq"][:12], None)):
        nn, d2 = np.zeros(n, dtype=np.int32), np.zeros(n)
        for j, x in enumerate(queries):
            b = ball_min_d2(pos, x, period)
            hit = b < r2
            nn[hit] = j + 1
            d2[hit] = b[hit]
        ref_nn, ref_d2 = X[key + "_nn_" + tag].copy(), X[key + "_d2_" + tag].copy()
        keep = np.ones(n, dtype=bool)
        if targets is not None:
            keep[targets] = False
        assert np.array_equal(nn[keep], ref_nn[keep]) and np.array_equal(d2[keep], ref_d2[keep])
    nn, d2 = np.zeros(n, dtype=np.int32), np.zeros(n)
    q = X["dense_q"]
    rows = crit_rows(pos, vel, pos[q], vel[q], 2, G["params"], period, exclude=q)
    for j, row in enumerate(rows):
        take = row[(nn[row] > j + 1) | (nn[row] == 0)]
        nn[take] = j + 1
        d = pos[q[j]][None, :] - pos[take]
        d2[take] = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    assert np.array_equal(nn, X["dense_c6_nn_" + tag])
    if tag == "np":      # periodic: the reference reports the distance to the image that matched; pinned for the marks only
        assert np.array_equal(d2, X["dense_c6_d2_" + tag])


def test_node_getters_known_answers(G, X):
    """cut values are the median particle's coordinate = the left child's upper boundary (KDTree.cxx:1012-1013); node IDs
    number the nodes depth first; FindLeafNode(tt) returns the <= bucket-sized leaf holding tt (KDFindNearest.cxx:709-722)"""
    assert np.array_equal(X["cut_ids"][:3], [0, 1, 2]) and len(X["cut_ids"]) == int(G["nodes"][0]) - int(G["nodes"][1])
    off, ids = X["leaf_off"], X["leaf_ids"]
    for j, q in enumerate(X["dense_q"]):
        leaf = ids[off[j]:off[j + 1]]
        assert q in leaf and len(leaf) <= 16


@pytest.mark.parametrize("tag", ["np", "p"])
def test_filtered_knn_restatement(G, X, tag):
    """FindNearestCheck / FindNearestCriterion of the reference (KDFindNearest.cxx:363-441) == k nearest (minimum image over
    the reference's reflections) among the particles other than the target that pass the filter; the periodic forms search
    k+1 and drop the nearest; rows are padded with (-1, 1e32)."""
    from tests.util import ball_min_d2, crit_rows
    period = None if tag == "np" else np.ones(3)
    pos, vel, kf, types = G["pos"], G["vel"], int(X["kf"]), X["types"]
    drop = 0 if period is None else 1
    for i in range(0, len(pos), 97):
        d2 = ball_min_d2(pos, pos[i], period)
        d2[i] = np.inf
        d2[types != 0] = np.inf
        want = np.sort(d2)[drop:kf + drop]
        assert np.array_equal(want, X["nnchk_d2_" + tag][i])
    q = np.arange(0, len(pos), 97)
    for crit, name in ((0, "c3"), (2, "c6")):
        rows = crit_rows(pos, vel, pos[q], vel[q], crit, G["params"], period, exclude=q)
        for j, i in enumerate(q):
            d2 = np.sort(ball_min_d2(pos, pos[i], period)[rows[j]])
            want = np.full(kf, 1e32)
            got = d2[drop:kf + drop]
            want[:len(got)] = got
            assert np.array_equal(want, X["nn%s_d2_%s" % (name, tag)][i])
            assert np.all(X["nn%s_ids_%s" % (name, tag)][i][len(got):] == -1)


def test_smoothed_velocity_moments_restatement(port, G, X):
    """CalcSmoothVel / CalcSmoothVelDisp of the reference (KDCalcSmoothQuantities.cxx:480-614) == symmetric gather + scatter
    with weights 0.5 W(r_ij, h_i) m / rho of the contributing particle; dispersion about the receiver's smoothed mean."""
    from tests.util import wsm_table
    _, kern = port.kernel_table(3, 2, 1000)
    ids, d2, k = G["knn0_ids_np"], G["knn0_d2_np"], int(G["k"])
    n, rho, m, vel = len(ids), X["sm_rho"], X["mass2"], G["vel"]
    hi = 0.5 * np.sqrt(d2[:, -1])
    W = np.array([[0.5 * wsm_table(kern, np.sqrt(d2[i, j]) / hi[i]) / hi[i] ** 3 for j in range(k)] for i in range(n)])
    sv, sd = np.zeros((n, 3)), np.zeros((n, 3, 3))
    ii = np.repeat(np.arange(n), k)
    jj = ids.ravel()
    w = W.ravel()
    np.add.at(sv, ii, (w / rho[jj] * m[jj])[:, None] * vel[jj])
    np.add.at(sv, jj, (w / rho[ii] * m[ii])[:, None] * vel[ii])
    np.testing.assert_allclose(sv, X["sm_vel"], rtol=0, atol=1e-12 * np.abs(X["sm_vel"]).max())
    a = vel[jj] - X["sm_vel"][ii]
    b = vel[ii] - X["sm_vel"][jj]
    np.add.at(sd, ii, (w / rho[jj] * m[jj])[:, None, None] * a[:, :, None] * a[:, None, :])
    np.add.at(sd, jj, (w / rho[ii] * m[ii])[:, None, None] * b[:, :, None] * b[:, None, :])
    np.testing.assert_allclose(sd, X["sm_disp"], rtol=0, atol=1e-12 * np.abs(X["sm_disp"]).max())
    # CalcSmoothVelSkew / CalcSmoothVelKurtosis (:617-765): third / fourth power over the RECEIVER's dispersion; the kurtosis
    # form subtracts 3 from every contribution
    dg = np.stack([X["sm_disp"][:, c, c] for c in range(3)], axis=1)
    sk, ku = np.zeros((n, 3)), np.zeros((n, 3))
    np.add.at(sk, ii, (w / rho[jj] * m[jj])[:, None] * a ** 3 / dg[ii] ** 1.5)
    np.add.at(sk, jj, (w / rho[ii] * m[ii])[:, None] * b ** 3 / dg[jj] ** 1.5)
    np.add.at(ku, ii, (w / rho[jj] * m[jj])[:, None] * a ** 4 / dg[ii] ** 2 - 3.0)
    np.add.at(ku, jj, (w / rho[ii] * m[ii])[:, None] * b ** 4 / dg[jj] ** 2 - 3.0)
    np.testing.assert_allclose(sk, X["sm_skew"], rtol=0, atol=1e-11 * np.abs(X["sm_skew"]).max())
    np.testing.assert_allclose(ku, X["sm_kurt"], rtol=0, atol=1e-11 * np.abs(X["sm_kurt"]).max())


def test_header_is_plain_c_and_shim_links(built, tmp_path):
    """include/nbk.h must be consumable from C (the boundary is a C ABI); the C++ shim and the demo written against the
    reference's interface must compile and link against libnbk.so (running it needs a GPU: tests/test_gpu_parity.py)."""
    import shutil
    import subprocess
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if gcc is None or gxx is None:
        pytest.skip("no host compiler")
    csrc = tmp_path / "use_nbk.c"
    csrc.write_text('#include "nbk.h"\nint main(void) { nbk_info i; nbk_particles p; (void)i; (void)p; return nbk_device_count() < 0; }\n')
    lib = os.path.join(ROOT, "nbodylib_b200")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"), str(csrc),
                           "-L" + lib, "-lnbk", "-Wl,-rpath," + lib, "-o", str(tmp_path / "use_nbk")])
    subprocess.check_call([gxx, "-O1", "-std=c++17", "-fopenmp", "-Wall", "-I" + os.path.join(lib, "shim"), os.path.join(ROOT, "examples", "shim_demo.cxx"),
                           "-L" + lib, "-lnbk", "-Wl,-rpath," + lib, "-o", str(tmp_path / "shim_demo")])
    # the sharded C ABI from plain C++ (examples/sharded_demo.cxx); without a GPU its ranks fail with a status code, not a crash
    subprocess.check_call([gxx, "-O1", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "sharded_demo.cxx"),
                           "-L" + lib, "-lnbk_sharded", "-lnbk", "-Wl,-rpath," + lib, "-o", str(tmp_path / "sharded_demo")])
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([str(tmp_path / "sharded_demo"), "1", "1000"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 1 and "nbk_comm_init_rank" in r.stderr


def test_shim_permutation_on_cpu(built, tmp_path):
    """the shim permutes the caller's 88-byte particles into tree order and back on the host (the reference does it with in-place
    swaps, KDTree.cxx:328-370,1347): byte path and move path of nbk_permute_records, ragged sizes"""
    import shutil
    import subprocess
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if gxx is None:
        pytest.skip("no host C++ compiler")
    lib = os.path.join(ROOT, "nbodylib_b200")
    exe = str(tmp_path / "permute_check")
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-fopenmp", "-Wall", "-Werror", "-I" + os.path.join(lib, "shim"),
                           os.path.join(ROOT, "tests", "cxx", "permute_check.cxx"), "-L" + lib, "-lnbk", "-Wl,-rpath," + lib, "-o", exe])
    for n in ("1", "7", "65537", "1000003"):
        out = subprocess.run([exe, n], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0 and " 0 errors" in out.stdout, out.stdout + out.stderr


def test_restatements_vs_live_reference(port):
    """The numpy restatements (oracle/restate_np.py) against the reference itself on a second, seeded input (not the golden
    one): single-target estimators, criterion search, filtered kNN.  Skipped where oracle/_ref did not travel."""
    from oracle import pyoracle
    from oracle.restate_np import ball_min_d2, crit_rows, gather_density, gather_veldensity
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref not built")
    from nbodylib_b200.synth import clustered_small
    n, k, kv = 4000, 24, 10
    pos, vel, mass = clustered_small(n, seed=5, nhalo=12)
    rng = np.random.default_rng(9)
    mass = (mass * (1.0 + rng.random(n))).astype(np.float32).astype(np.float64)
    types = (rng.random(n) < 0.4).astype(np.int32)
    ll = 0.3 / n ** (1.0 / 3)
    params = np.zeros(10)
    params[1] = params[6] = ll * ll
    params[2] = params[7] = float(((vel - vel.mean(0)) ** 2).sum(1).mean() / 3) * 0.5
    q = rng.choice(n, 60, replace=False).astype(np.int32)
    _, kern = port.kernel_table(3, 2, 1000)
    for period in (None, np.ones(3)):
        R = pyoracle.Ref(pos, vel, mass, period=period)
        ids, d2 = port.knn_particles(pos, k)                       # Calc* never wrap (Q2): the non periodic lists
        assert np.array_equal(np.array([gather_density(kern, d2[i], mass[ids[i]]) for i in q]), R.calc_density_particles(q, k))
        np.testing.assert_allclose(np.array([gather_veldensity(kern, vel[i], vel[ids[i]], kv) for i in q]), R.calc_veldensity_particles(q, kv, k), rtol=1e-14)
        for crit in (0, 2):
            off, idx = R.search_criterion_particles(q, crit, params)
            rows = crit_rows(pos, vel, pos[q], vel[q], crit, params, period, exclude=q)
            assert all(np.array_equal(rows[j], np.sort(idx[off[j]:off[j + 1]])) for j in range(len(q)))
        R.set_types(types)
        nn, dd = R.knn_filtered(8)
        drop = 0 if period is None else 1
        for i in q:
            b = ball_min_d2(pos, pos[i], period)
            b[i] = np.inf
            b[types != 0] = np.inf
            assert np.array_equal(np.sort(b)[drop:8 + drop], dd[i])
        R.close()
