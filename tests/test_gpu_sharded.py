"""Multi-GPU parity (-m gpu): the slab-sharded driver with the CUDA engine over NCCL against the single-GPU tree on the same
global particle set (needs >= 2 GPUs, skipped otherwise), and its single-GPU building blocks (nbk_fof_roots, nbk_union_pairs,
the world-1 driver) on one GPU."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import canon

pytestmark = pytest.mark.gpu


def _params6d(pos, vel, n):
    ll = 0.25 / n ** (1 / 3)
    sv2 = ((vel - vel.mean(0)) ** 2).sum(1).mean() / 3.0
    params = np.zeros(10)
    params[1] = params[6] = ll * ll
    params[2] = params[7] = sv2
    return params


def _worker(rank, world, port, tmp, n, native):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from nbodylib_b200.sharded import NativeShardedTree, ShardedTree
        if native:                  # the C ABI (include/nbk_sharded.h): exchange and merge in C++ over the library's own NCCL communicator
            ShardedTree = NativeShardedTree
        from nbodylib_b200.synth import clustered_small
        pos, vel, mass = clustered_small(n, seed=5)
        edges = None
        if world == 4:                          # slab faces at the x quantiles (equal counts, unequal widths); equal widths at world 2
            xs = np.sort(pos[:, 0])
            edges = np.array([0.0] + [xs[(n * r) // world] for r in range(1, world)] + [1.0])
            slab = np.searchsorted(edges, pos[:, 0], side="right") - 1
        else:
            slab = np.minimum((pos[:, 0] * world).astype(int), world - 1)
        mine = np.nonzero(slab == rank)[0]
        dev = torch.device("cuda", rank)
        f = torch.float32                       # clustered_small is fp32-representable: the slab trees keep fp32 storage
        st = ShardedTree(torch.from_numpy(pos[mine]).to(dev, f), torch.from_numpy(vel[mine]).to(dev, f), torch.from_numpy(mass[mine]).to(dev, f),
                         period=np.ones(3), rank=rank, world=world, box=(1.0, 1.0, 1.0), knn_k=32, edges=edges)
        rho = st.CalcDensity(32)
        ll = 0.25 / n ** (1 / 3)
        g, ng = st.FOF(ll, 8, 1)
        g6, ng6 = st.FOFCriterion(2, _params6d(pos, vel, n), 8, 1)
        assert st.stats["fof_setups"] == 2      # one slab tree without, one with velocities
        np.savez(os.path.join(tmp, "r%d.npz" % rank), idx=mine, rho=rho.cpu().numpy(), g=g.cpu().numpy(), g6=g6.cpu().numpy(), ng=np.array([ng, ng6]))
        st.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("native", [False, True], ids=["torch-driver", "c-abi"])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_matches_single_gpu(built, world, native):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import nbodylib_b200 as nb
    from nbodylib_b200.synth import clustered_small
    n = 200000
    pos, vel, mass = clustered_small(n, seed=5)
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(world, 29500 + np.random.randint(0, 2000), tmp, n, native), nprocs=world, join=True)
        res = [np.load(os.path.join(tmp, "r%d.npz" % r)) for r in range(world)]
    rho = np.zeros(n)
    g, g6 = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
    for r in res:
        rho[r["idx"]] = r["rho"]
        g[r["idx"]] = r["g"]
        g6[r["idx"]] = r["g6"]
    ll = 0.25 / n ** (1 / 3)
    with nb.KDTree(pos, vel, mass, Period=np.ones(3)) as t:
        np.testing.assert_allclose(rho, t.CalcDensity(32), rtol=1e-10)
        g1, ng1 = t.FOF(ll, 8, 1)
        h1, nh1 = t.FOFCriterion(nb.FOF6D, _params6d(pos, vel, n), 8, 1)
    assert int(res[0]["ng"][0]) == ng1 and np.array_equal(canon(g), canon(g1))
    assert np.array_equal(np.bincount(g)[1:], np.bincount(g1)[1:])
    assert int(res[0]["ng"][1]) == nh1 and np.array_equal(canon(g6), canon(h1))


def _variant_worker(rank, world, port, tmp, case):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from nbodylib_b200.sharded import NativeShardedTree
        pos, vel, mass, box, periodic, k, ll, minnum, order = _variant(case)
        n = len(pos)
        slab = np.minimum((pos[:, 0] / box[0] * world).astype(int), world - 1)
        mine = np.nonzero(slab == rank)[0]
        # host arrays straight into the C ABI (on_device = 0), fp64
        st = NativeShardedTree(torch.from_numpy(pos[mine]), torch.from_numpy(vel[mine]), None if mass is None else torch.from_numpy(mass[mine]),
                               period=box if periodic else None, rank=rank, world=world, box=box, knn_k=k, device=rank)
        rho = st.CalcDensity(k)
        rho2 = st.CalcDensity(k)                               # second call on the resident trees
        g, ng = st.FOF(ll, minnum, order)
        g6, ng6 = st.FOFCriterion(2, _params6d(pos, vel, n), minnum, order)
        info = st.stats
        assert info["density_setups"] >= 1 and info["n_global"] == n and info["nranks"] == world
        assert torch.equal(rho, rho2) or torch.allclose(rho, rho2, rtol=1e-12, atol=0)
        np.savez(os.path.join(tmp, "r%d.npz" % rank), idx=mine, rho=rho.cpu().numpy(), g=g.cpu().numpy(), g6=g6.cpu().numpy(), ng=np.array([ng, ng6]))
        st.close()
        NativeShardedTree.shutdown()
    finally:
        dist.destroy_process_group()


def _variant(case):
    """(pos, vel, mass, box, periodic, k, ll, minnum, order) of the extra C-ABI cases"""
    from nbodylib_b200.synth import clustered_small
    if case == "fp64-open":
        # coordinates that fp32 cannot hold (the slab trees must keep fp64 storage), open box, unordered ids, singletons kept out
        rng = np.random.default_rng(11)
        pos, vel, mass = clustered_small(60000, seed=3)
        pos = np.clip(pos + 1e-9 * rng.standard_normal(pos.shape), 1e-12, 1 - 1e-12)
        return pos, vel, mass, np.ones(3), False, 20, 0.3 / 60000 ** (1 / 3), 2, 0
    if case == "box-2x1x1":
        # a non-cubic periodic box, unit masses (mass = NULL), large groups only
        rng = np.random.default_rng(12)
        a, va, _ = clustered_small(50000, seed=4)
        b, vb, _ = clustered_small(50000, seed=5)
        pos = np.concatenate([a, b + np.array([1.0, 0, 0])])
        vel = np.concatenate([va, vb])
        return pos, vel, None, np.array([2.0, 1.0, 1.0]), True, 48, 0.25 / 50000 ** (1 / 3), 30, 1
    raise KeyError(case)


@pytest.mark.parametrize("world,case", [(3, "fp64-open"), (2, "box-2x1x1"), (4, "box-2x1x1")])
def test_sharded_c_abi_variants(built, world, case):
    """The C ABI on inputs the headline case does not touch: fp64 coordinates that are not fp32-representable, host pointers, an
    open box, a non-cubic periodic box, unit masses, an odd number of ranks, both id orders."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import nbodylib_b200 as nb
    pos, vel, mass, box, periodic, k, ll, minnum, order = _variant(case)
    n = len(pos)
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_variant_worker, args=(world, 29500 + np.random.randint(0, 2000), tmp, case), nprocs=world, join=True)
        res = [np.load(os.path.join(tmp, "r%d.npz" % r)) for r in range(world)]
    rho = np.zeros(n)
    g, g6 = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
    for r in res:
        rho[r["idx"]] = r["rho"]
        g[r["idx"]] = r["g"]
        g6[r["idx"]] = r["g6"]
    with nb.KDTree(pos, vel, mass, Period=box if periodic else None) as t:
        np.testing.assert_allclose(rho, t.CalcDensity(k), rtol=1e-10)
        g1, ng1 = t.FOF(ll, minnum, order)
        h1, nh1 = t.FOFCriterion(nb.FOF6D, _params6d(pos, vel, n), minnum, order)
    assert int(res[0]["ng"][0]) == ng1 and np.array_equal(canon(g), canon(g1)) and g.max() == ng1
    assert int(res[0]["ng"][1]) == nh1 and np.array_equal(canon(g6), canon(h1))
    if order:
        sizes = np.bincount(g)[1:]
        assert np.all(np.diff(sizes) <= 0) and np.array_equal(sizes, np.bincount(g1)[1:])


def test_sharded_demo_from_plain_cxx(built, tmp_path):
    """examples/sharded_demo.cxx: the C ABI from a C++ program without Python in the ranks -- one forked process per GPU, the
    communicator id through a file, system NCCL -- against a single tree in the same program.  One GPU: a single rank."""
    import shutil
    import subprocess
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if gxx is None:
        pytest.skip("no host compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "nbodylib_b200")
    exe = str(tmp_path / "sharded_demo")
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "sharded_demo.cxx"),
                           "-L" + lib, "-lnbk_sharded", "-lnbk", "-Wl,-rpath," + lib, "-o", exe])
    nranks = min(4, torch.cuda.device_count())
    out = subprocess.run([exe, str(nranks), "400000"], capture_output=True, text=True, timeout=600)
    print(out.stdout[-1500:], out.stderr[-1500:])
    assert out.returncode == 0 and "SHARDED DEMO OK" in out.stdout


def test_world1_driver_and_building_blocks(built):
    """One GPU: the sharded driver with a single rank (no process group needed) equals the plain tree, which exercises
    nbk_fof_roots; nbk_union_pairs is checked on a random edge list against SciPy."""
    import nbodylib_b200 as nb
    from nbodylib_b200.sharded import CudaEngine, ShardedTree
    from nbodylib_b200.synth import clustered_small
    n = 60000
    pos, vel, mass = clustered_small(n, seed=9)
    dev = torch.device("cuda", 0)
    st = ShardedTree(torch.from_numpy(pos).to(dev, torch.float32), torch.from_numpy(vel).to(dev, torch.float32), torch.from_numpy(mass).to(dev, torch.float32),
                     period=np.ones(3), rank=0, world=1, box=(1.0, 1.0, 1.0), knn_k=24)
    ll = 0.25 / n ** (1 / 3)
    rho = st.CalcDensity(24).cpu().numpy()
    g, ng = st.FOF(ll, 6, 1)
    g6, ng6 = st.FOFCriterion(2, _params6d(pos, vel, n), 6, 0)
    st.close()
    with nb.KDTree(pos, vel, mass, Period=np.ones(3)) as t:
        np.testing.assert_allclose(rho, t.CalcDensity(24), rtol=1e-10)
        g1, ng1 = t.FOF(ll, 6, 1)
        h1, nh1 = t.FOFCriterion(nb.FOF6D, _params6d(pos, vel, n), 6, 0)
        roots = t.FOFRoots(ll)
    assert ng == ng1 and np.array_equal(canon(g.cpu().numpy()), canon(g1)) and np.array_equal(np.bincount(g.cpu().numpy())[1:], np.bincount(g1)[1:])
    assert ng6 == nh1 and np.array_equal(canon(g6.cpu().numpy()), canon(h1))
    # the C ABI with a single rank (no NCCL communicator is created)
    from nbodylib_b200.sharded import NativeShardedTree
    nt = NativeShardedTree(torch.from_numpy(pos).to(dev, torch.float32), torch.from_numpy(vel).to(dev, torch.float32), torch.from_numpy(mass).to(dev, torch.float32),
                           period=np.ones(3), rank=0, world=1, box=(1.0, 1.0, 1.0), knn_k=24)
    np.testing.assert_allclose(nt.CalcDensity(24).cpu().numpy(), rho, rtol=1e-12)
    gn, ngn = nt.FOF(ll, 6, 1)
    g6n, ng6n = nt.FOFCriterion(2, _params6d(pos, vel, n), 6, 0)
    assert nt.stats["fof_setups"] == 2 and nt.stats["n_global"] == n
    nt.close()
    assert ngn == ng1 and np.array_equal(canon(gn.cpu().numpy()), canon(g1)) and np.array_equal(np.bincount(gn.cpu().numpy())[1:], np.bincount(g1)[1:])
    assert ng6n == nh1 and np.array_equal(canon(g6n.cpu().numpy()), canon(h1))
    # roots: every member of a group shares one representative, which is a member itself
    assert np.array_equal(roots[roots], roots)
    sel = g1 > 0
    assert len(np.unique(roots[sel])) == ng1
    # connected components of a random graph
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    rng = np.random.default_rng(1)
    nn, ne = 50000, 40000
    a, b = rng.integers(0, nn, ne).astype(np.int32), rng.integers(0, nn, ne).astype(np.int32)
    root = CudaEngine(0).union_pairs(nn, torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)).cpu().numpy()
    ncomp, comp = connected_components(coo_matrix((np.ones(ne, dtype=np.int8), (a, b)), shape=(nn, nn)), directed=False)
    first = np.full(ncomp, nn, dtype=np.int64)
    np.minimum.at(first, comp, np.arange(nn))
    assert np.array_equal(root, first[comp])
