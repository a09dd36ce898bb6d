"""GPU parity of the phase-space nearest-neighbour search (-m gpu): KDTree::FindNearestPhase(Int_t tt | x, v) and
FindNearest on a TPHS tree built with Aniso = -1 (reference KDFindNearest.cxx:260-262,300-301,347-361,543-555), through the
C ABI (nbk_knn_phase_particles / nbk_knn_phase_points), against the reference's own rows (tests/golden/ref_phase.npz, made
by tests/golden/make_golden_phase.py from the live reference) and the brute-force port on seeded inputs.

Bar: neighbour sets and every d2 bit-exact (fp64, the reference's operation order)."""
import numpy as np
import pytest

from tests.util import load_golden, load_golden_phase, rows_equal_as_sets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nb():
    # the prebuilt in-tree library, loaded as a consumer would (numpy + ctypes only); a missing libnbk.so raises here
    import nbodylib_b200
    assert nbodylib_b200._lib.load().nbk_device_count() > 0, "GPU tests need a CUDA device"
    return nbodylib_b200


@pytest.fixture(scope="module")
def port():
    from oracle.pyoracle import Port
    return Port()


def by_id(order, rows):
    out = np.empty_like(rows)
    out[order] = rows
    return out


@pytest.mark.parametrize("tag", ["np", "p"])
@pytest.mark.parametrize("treetype", ["TPHS", "TPHYS"])
def test_golden_phase_knn(nb, tag, treetype):
    G, H = load_golden(), load_golden_phase()
    period = None if tag == "np" else np.ones(3)
    k, qsel = int(H["k"]), H["qsel"]
    with nb.KDTree(G["pos"], H["vel"], G["mass"], TreeType=getattr(nb, treetype), Period=period, Aniso=-1) as t:
        order = t.order()
        nn, d2 = t.FindNearestPhase(k, ids=True)
        assert np.all(np.diff(d2, axis=1) >= 0)
        assert np.array_equal(by_id(order, d2)[qsel], H["phase_d2_" + tag])
        assert rows_equal_as_sets(by_id(order, nn)[qsel], H["phase_ids_" + tag])
        # tree-index outputs map through order() to the same IDs; sub-ranges return the same rows
        nn_t, d2_t = t.FindNearestPhase(k, q0=100, q1=1333)
        assert np.array_equal(order[nn_t], nn[100:1333]) and np.array_equal(d2_t, d2[100:1333])
        if treetype == "TPHS":
            nn1, d21 = t.FindNearest(k, ids=True)             # KDFindNearest.cxx:260-262: the same search
            assert np.array_equal(nn1, nn) and np.array_equal(d21, d2)
        nx, dx = t.FindNearestPhase(k, x=G["xq"], v=H["vq"], ids=True)
        assert np.array_equal(dx, H["phasex_d2_" + tag]) and rows_equal_as_sets(nx, H["phasex_ids_" + tag])


@pytest.mark.parametrize("n,period,k,flags", [(20011, None, 32, 0), (12007, 1, 24, 0), (4099, 1, 9, 1 << 4), (31, None, 8, 0), (17, 1, 5, 0)])
def test_phase_port_parity_seeded(nb, port, n, period, k, flags):
    """ragged sizes (non power of two, two leaves), fp32 and forced fp64 storage, against the brute-force port"""
    from nbodylib_b200.synth import clustered_small
    pos, vel, mass = clustered_small(n, seed=n + 1)
    vel = (vel * (0.05 / vel.std())).astype(np.float32).astype(np.float64)
    period = None if period is None else np.ones(3)
    rng = np.random.default_rng(n)
    xq = rng.random((257, 3))
    vq = rng.normal(size=(257, 3)) * 0.05
    with nb.KDTree(pos, vel, mass, TreeType=nb.TPHS, Period=period, flags=flags, Aniso=-1) as t:
        order = t.order()
        nn, d2 = t.FindNearestPhase(k, ids=True)
        oi, od = port.knn_phase_particles(pos, vel, np.arange(n, dtype=np.int32), k, period=period)
        assert np.array_equal(by_id(order, d2), od) and rows_equal_as_sets(by_id(order, nn), oi)
        nx, dx = t.FindNearestPhase(k, x=xq, v=vq, ids=True)
        oi, od = port.knn_phase_points(pos, vel, xq, vq, k, period=period)
        assert np.array_equal(dx, od) and rows_equal_as_sets(nx, oi)


def test_phase_knn_config1_scale_against_reference_library(nb):
    """BASELINE config 1's geometry (10^6 uniform-random particles, periodic unit box, bucket 16, k = 32) for the phase-space search,
    against the reference LIBRARY itself (oracle/_ref, a TPHS tree built with Aniso = -1): a 20 000-particle sample of
    FindNearestPhase(tt) and 2 000 phase-space points, bit-exact"""
    from oracle.pyoracle import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref did not travel with this checkout")
    n, k = 1000000, 32
    rng = np.random.default_rng(2024)
    pos = rng.random((n, 3)).astype(np.float32).astype(np.float64)
    vel = (rng.normal(size=(n, 3)) * 0.02).astype(np.float32).astype(np.float64)
    qs = np.sort(rng.choice(n, 20000, replace=False)).astype(np.int32)
    xq = rng.random((2000, 3))
    vq = rng.normal(size=(2000, 3)) * 0.02
    period = np.ones(3)
    R = Ref(pos, vel, None, treetype=Ref.TPHS, period=period, aniso=-1)
    ri, rd = R.knn_phase_particles(qs, k, which=0)
    rx, rxd = R.knn_phase_points(xq, vq, k)
    R.close()
    with nb.KDTree(pos, vel, None, TreeType=nb.TPHS, Period=period, Aniso=-1) as t:
        assert t.info.store_bytes == 4
        order = t.order()
        nn, d2 = t.FindNearestPhase(k, ids=True)
        assert np.array_equal(by_id(order, d2)[qs], rd) and rows_equal_as_sets(by_id(order, nn)[qs], ri)
        nx, dx = t.FindNearestPhase(k, x=xq, v=vq, ids=True)
        assert np.array_equal(dx, rxd) and rows_equal_as_sets(nx, rx)


def test_phase_knn_degenerate_inputs(nb, port):
    """coincident phase-space points are never neighbours of one another in the particle form (KDLeafNode.cxx:43-57: dist2 > 0)
    but are found by the coordinate form; fewer candidates than k pads with (-1, 1e32) like every search (KDFindNearest.cxx:16-19)"""
    rng = np.random.default_rng(5)
    pos = rng.random((600, 3)).astype(np.float32).astype(np.float64)
    vel = (rng.normal(size=(600, 3)) * 0.1).astype(np.float32).astype(np.float64)
    pos[300:] = pos[:300]
    vel[300:] = vel[:300]                                   # every point twice
    with nb.KDTree(pos, vel, None, TreeType=nb.TPHS, Aniso=-1) as t:
        order = t.order()
        nn, d2 = t.FindNearestPhase(6, ids=True)
        oi, od = port.knn_phase_particles(pos, vel, np.arange(600, dtype=np.int32), 6)
        assert np.all(d2 > 0) and np.array_equal(by_id(order, d2), od)
        nx, dx = t.FindNearestPhase(2, x=pos[:50], v=vel[:50], ids=True)
        assert np.all(dx == 0) and np.array_equal(np.sort(nx, axis=1), np.stack([np.arange(50), np.arange(300, 350)], axis=1))
    with nb.KDTree(pos[:5], vel[:5], None, TreeType=nb.TPHS, Aniso=-1) as t:
        nn, d2 = t.FindNearestPhase(8, ids=True)
        assert np.all(nn[:, 4:] == -1) and np.all(d2[:, 4:] == 1e32) and np.all(nn[:, :4] >= 0)


def test_findnearestvel_on_velocity_trees(nb, port):
    """KDTree::FindNearestVel(tt | v) (KDFindNearest.cxx:335-346,530-540): velocity-space neighbours on a TVEL tree, periodic or
    not (velocity searches are never reflected, KDSplitNode.cxx:1082-1085); refused on trees whose cut planes are positions"""
    from nbodylib_b200.synth import clustered_small
    pos, vel, mass = clustered_small(6007, seed=31)
    vq = np.random.default_rng(9).normal(size=(129, 3)) * vel.std()
    for period in (None, np.ones(3)):
        with nb.KDTree(pos, vel, mass, TreeType=nb.TVEL, Period=period) as t:
            order = t.order()
            nn, d2 = t.FindNearestVel(10, ids=True)
            oi, od = port.knn_particles(vel, 10)
            assert np.array_equal(by_id(order, d2), od) and rows_equal_as_sets(by_id(order, nn), oi)
            nx, dx = t.FindNearestVel(10, v=vq, ids=True)
            oi, od = port.knn_points(vel, vq, 10)
            assert np.array_equal(dx, od) and rows_equal_as_sets(nx, oi)
    with nb.KDTree(pos, vel, mass) as t:
        with pytest.raises(nb.NbkError) as e:
            t.FindNearestVel(10)
        assert e.value.code == -3


def test_phase_knn_refusals(nb):
    """no silent approximations: metric searches (Aniso >= 0, quirk Q4), velocity trees and trees without velocities are refused"""
    rng = np.random.default_rng(1)
    pos, vel = rng.random((500, 3)), rng.normal(size=(500, 3))
    with nb.KDTree(pos, vel, None, TreeType=nb.TPHS) as t:       # constructor default Aniso = 0
        with pytest.raises(nb.NbkError) as e:
            t.FindNearest(8)
        assert e.value.code == -3
        with pytest.raises(nb.NbkError) as e:
            t.FindNearestPos(8)
        assert e.value.code == -3
        t.FindNearestPhase(8)                                    # FindNearestPhase never looks at Aniso (KDFindNearest.cxx:347-361)
    with nb.KDTree(pos, vel, None, TreeType=nb.TVEL) as t:
        with pytest.raises(nb.NbkError) as e:
            t.FindNearestPhase(8)
        assert e.value.code == -3
    with nb.KDTree(pos) as t:
        with pytest.raises(nb.NbkError) as e:
            t.FindNearestPhase(8)
        assert e.value.code == -1


def test_cxx_shim_phase_program(nb, tmp_path):
    """examples/shim_phase_demo.cxx: FindNearest / FindNearestPhase on TPHS trees through the C++ shim against a host scan"""
    import os
    import shutil
    import subprocess
    from tests.util import ROOT
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if gxx is None:
        pytest.skip("no host C++ compiler")
    exe = str(tmp_path / "shim_phase_demo")
    lib = os.path.join(ROOT, "nbodylib_b200")
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-fopenmp", "-I" + os.path.join(lib, "shim"), os.path.join(ROOT, "examples", "shim_phase_demo.cxx"),
                           "-L" + lib, "-lnbk", "-Wl,-rpath," + lib, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "shim phase demo ok" in out.stdout, out.stdout + out.stderr
    assert out.stdout.count(": 0 mismatches against brute force") == 3          # two TPHS trees, one TVEL tree
