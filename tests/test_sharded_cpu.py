"""World-size-2 gloo tests (CPU) of the slab-sharded driver's HOST logic: halo exchange, return of the scatter
term to the owners, halo widening, cross-slab FOF merge and global numbering.  The per-rank compute engine is
replaced by the brute-force oracle (test infrastructure), so no GPU is needed; the same driver runs on the GPU
box with the CUDA engine (tests/test_gpu_sharded.py)."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import canon


class PortEngine:
    """oracle-backed stand-in for nbodylib_b200.sharded.CudaEngine"""

    def __init__(self):
        from oracle.pyoracle import Port
        self.P = Port()
        self.norm, self.table = self.P.kernel_table(3, 2, 1000)

    def build(self, pos, vel, mass, period):
        return {"pos": pos.double().numpy().copy(), "vel": None if vel is None else vel.double().numpy().copy(), "period": period}

    def build_with_halo(self, pos, mass, gpos, gmass):
        return {"pos": np.concatenate([pos.double().numpy(), gpos.double().numpy()]), "mass": np.concatenate([mass.double().numpy(), gmass.double().numpy()]),
                "n_main": len(pos)}

    def density(self, tree, k, rho, hsm):
        pos, mass, n1 = tree["pos"], tree["mass"], tree["n_main"]
        ids, d2 = self.P.knn_particles(pos, k)
        h = 0.5 * np.sqrt(d2[:, -1])
        r = np.sqrt(d2) / h[:, None]
        i = (r * 0.5 * 999).astype(np.int64)
        delta = 2.0 / 999
        t = self.table
        w = np.where(i < 999, t[np.minimum(i, 998)] + (t[np.minimum(i + 1, 999)] - t[np.minimum(i, 998)]) * (r - delta * i) / delta, t[999])
        W = 0.5 * w / h[:, None] ** 3
        out = np.zeros(len(pos))
        q = np.arange(n1)                                      # queries: the main particles only
        out[q] += (W[q] * mass[ids[q]]).sum(1)
        np.add.at(out, ids[q].ravel(), (W[q] * mass[q][:, None]).ravel())
        rho.copy_(torch.from_numpy(out))
        hh = np.zeros(len(pos))
        hh[q] = h[q]
        hsm.copy_(torch.from_numpy(hh))
        return None

    def fof_roots(self, tree, fdist, criterion, params, out):
        if criterion < 0:
            g, ng = self.P.fof(tree["pos"], None, 0, [fdist * fdist], tree["period"], 1, 0)
        else:
            g, ng = self.P.fof(tree["pos"], tree["vel"], 2 if criterion == 0 else 4, params, tree["period"], 1, 0)
        # representative = the component's member with the largest index (any fixed member will do)
        rep = np.zeros(ng + 1, dtype=np.int64)
        rep[g] = np.arange(len(g))
        out.copy_(torch.from_numpy(rep[g].astype(np.int32)))
        return None

    def union_pairs(self, nnodes, a, b):
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components
        a, b = a.numpy(), b.numpy()
        m = coo_matrix((np.ones(len(a), dtype=np.int8), (a, b)), shape=(nnodes, nnodes))
        _, comp = connected_components(m, directed=False)
        first = np.full(comp.max() + 1 if nnodes else 0, nnodes, dtype=np.int64)
        np.minimum.at(first, comp, np.arange(nnodes))
        return torch.from_numpy(first[comp].astype(np.int32))


def _params6d(pos, vel, n):
    ll = 0.3 / n ** (1 / 3)
    sv2 = ((vel - vel.mean(0)) ** 2).sum(1).mean() / 3.0
    params = np.zeros(10)
    params[1] = params[6] = ll * ll
    params[2] = params[7] = sv2
    return params


def _worker(rank, world, port, tmp, halo, case):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nbodylib_b200.sharded import ShardedTree
        pos, vel, mass = _case(case)
        n = len(pos)
        edges = None
        if case == "quantile":                 # slab faces at the x quantiles: equal particle counts, unequal widths
            xs = np.sort(pos[:, 0])
            edges = np.array([0.0] + [xs[(n * r) // world] for r in range(1, world)] + [1.0])
            slab = np.searchsorted(edges, pos[:, 0], side="right") - 1
        else:
            slab = np.minimum((pos[:, 0] * world).astype(int), world - 1)
        mine = np.nonzero(slab == rank)[0]
        dt = torch.float32 if case == "fp32" else torch.float64
        st = ShardedTree(torch.from_numpy(pos[mine]).to(dt), torch.from_numpy(vel[mine]).to(dt), torch.from_numpy(mass[mine]).to(dt), period=np.ones(3),
                         rank=rank, world=world, box=(1.0, 1.0, 1.0), halo=halo, knn_k=16, engine=PortEngine(), edges=edges)
        rho = st.CalcDensity(16)
        ll = 0.3 / n ** (1 / 3)
        g0, ng0 = st.FOF(ll, 5, 0)
        g1, ng1 = st.FOF(ll, 5, 1)
        assert st.stats["fof_setups"] == 1                      # the slab tree stayed resident between the two calls
        g6, ng6 = st.FOFCriterion(2, _params6d(pos, vel, n), 5, 1)
        np.savez(os.path.join(tmp, "r%d.npz" % rank), idx=mine, rho=rho.numpy(), g0=g0.numpy(), g1=g1.numpy(), g6=g6.numpy(), ng=np.array([ng0, ng1, ng6]),
                 h=np.array([st.h_knn]), ghosts=np.array([st.stats["ghosts_knn"], st.stats["ghosts_fof"]]))
        st.close()
    finally:
        dist.destroy_process_group()


def _case(case):
    from nbodylib_b200.synth import clustered_small
    if case == "uneven":
        # a dense slab next to a nearly empty one (the per-rank default halo widths differ by a factor of several): the halo
        # width has to be ONE number for the group, or the dense side's k-balls reach past the ghosts it received
        rng = np.random.default_rng(5)
        a = rng.random((4000, 3)) * np.array([0.1, 1, 1])
        b = rng.random((300, 3)) * np.array([0.1, 1, 1]) + np.array([0.2, 0, 0])
        c = rng.random((40, 3)) * np.array([0.5, 1, 1]) + np.array([0.5, 0, 0])
        pos = np.concatenate([a, b, c]).astype(np.float32).astype(np.float64)
        vel = (0.05 * rng.standard_normal(pos.shape)).astype(np.float32).astype(np.float64)
        mass = (1.0 + rng.random(len(pos))).astype(np.float32).astype(np.float64)
        return pos, vel, mass
    return clustered_small(5000, seed=77)


@pytest.mark.parametrize("world,halo,case", [(2, None, "clustered"), (2, 0.01, "clustered"), (3, None, "clustered"), (4, None, "fp32"), (2, None, "uneven"),
                                             (3, None, "quantile")])
def test_sharded_host_logic_matches_single_domain(port, world, halo, case):
    pos, vel, mass = _case(case)
    n = len(pos)
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(world, 29500 + np.random.randint(0, 2000), tmp, halo, case), nprocs=world, join=True)
        res = [np.load(os.path.join(tmp, "r%d.npz" % r)) for r in range(world)]
    rho = np.zeros(n)
    g0, g1, g6 = (np.zeros(n, dtype=np.int64) for _ in range(3))
    for r in res:
        rho[r["idx"]] = r["rho"]
        g0[r["idx"]] = r["g0"]
        g1[r["idx"]] = r["g1"]
        g6[r["idx"]] = r["g6"]
    ref_rho, _ = port.density(pos, mass, 16)
    np.testing.assert_allclose(rho, ref_rho, rtol=1e-12)
    assert len({float(r["h"][0]) for r in res}) == 1              # one halo width for the whole group
    if halo is not None:                       # a deliberately narrow halo must have been widened
        assert all(float(r["h"][0]) > halo for r in res)
    ll = 0.3 / n ** (1 / 3)
    ref_g, ref_ng = port.fof(pos, None, 0, [ll * ll], np.ones(3), 5, 1)
    assert int(res[0]["ng"][0]) == int(res[0]["ng"][1]) == ref_ng
    assert np.array_equal(canon(g0), canon(ref_g)) and np.array_equal(canon(g1), canon(ref_g))
    assert g0.max() == ref_ng and g1.max() == ref_ng
    sizes = np.bincount(g1)[1:]
    assert np.all(np.diff(sizes) <= 0) and np.array_equal(np.sort(sizes), np.sort(np.bincount(ref_g)[1:]))
    ref6, ref_ng6 = port.fof(pos, vel, 4, _params6d(pos, vel, n), np.ones(3), 5, 1)
    assert int(res[0]["ng"][2]) == ref_ng6 and np.array_equal(canon(g6), canon(ref6))
    assert all(int(r["ghosts"][0]) > 0 and int(r["ghosts"][1]) >= 0 for r in res)


def _halo_too_wide_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nbodylib_b200.sharded import ShardedTree
        pos, vel, mass = _case("clustered")
        slab = np.minimum((pos[:, 0] * world).astype(int), world - 1)
        mine = np.nonzero(slab == rank)[0]
        st = ShardedTree(torch.from_numpy(pos[mine]), None, torch.from_numpy(mass[mine]), period=np.ones(3), rank=rank, world=world,
                         box=(1.0, 1.0, 1.0), engine=PortEngine())
        with pytest.raises(ValueError):
            st.FOF(0.3, 5, 0)                  # linking length above the slab width (0.25): links two slabs away would be lost
    finally:
        dist.destroy_process_group()


def test_halo_wider_than_a_slab_is_refused():
    mp.spawn(_halo_too_wide_worker, args=(4, 29500 + np.random.randint(0, 2000)), nprocs=4, join=True)
