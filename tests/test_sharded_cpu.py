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
        return {"pos": pos.numpy().copy(), "mass": None if mass is None else mass.numpy().copy(), "period": period}

    def density(self, tree, k, active, rho, hsm):
        pos, mass = tree["pos"], tree["mass"]
        ids, d2 = self.P.knn_particles(pos, k)
        act = active.numpy().astype(bool)
        h = 0.5 * np.sqrt(d2[:, -1])
        r = np.sqrt(d2) / h[:, None]
        i = (r * 0.5 * 999).astype(np.int64)
        delta = 2.0 / 999
        t = self.table
        w = np.where(i < 999, t[np.minimum(i, 998)] + (t[np.minimum(i + 1, 999)] - t[np.minimum(i, 998)]) * (r - delta * i) / delta, t[999])
        W = 0.5 * w / h[:, None] ** 3
        out = np.zeros(len(pos))
        q = np.nonzero(act)[0]
        out[q] += (W[q] * mass[ids[q]]).sum(1)
        np.add.at(out, ids[q].ravel(), (W[q] * mass[q][:, None]).ravel())
        rho.copy_(torch.from_numpy(out))
        hh = np.zeros(len(pos))
        hh[q] = h[q]
        hsm.copy_(torch.from_numpy(hh))
        return None

    def fof_labels(self, tree, ll, out):
        g, ng = self.P.fof(tree["pos"], None, 0, [ll * ll], tree["period"], 1, 0)
        out.copy_(torch.from_numpy(g))
        return ng


def _worker(rank, world, port, tmp, halo):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nbodylib_b200.sharded import ShardedTree
        from nbodylib_b200.synth import clustered_small
        pos, vel, mass = clustered_small(5000, seed=77)
        slab = np.minimum((pos[:, 0] * world).astype(int), world - 1)
        mine = np.nonzero(slab == rank)[0]
        st = ShardedTree(torch.from_numpy(pos[mine]), torch.from_numpy(vel[mine]), torch.from_numpy(mass[mine]), period=np.ones(3),
                         rank=rank, world=world, box=(1.0, 1.0, 1.0), slab_local=False, halo=halo, knn_k=16, engine=PortEngine())
        rho = st.CalcDensity(16)
        ll = 0.3 / 5000 ** (1 / 3)
        g0, ng0 = st.FOF(ll, 5, 0)
        g1, ng1 = st.FOF(ll, 5, 1)
        np.savez(os.path.join(tmp, "r%d.npz" % rank), idx=mine, rho=rho.numpy(), g0=g0.numpy(), g1=g1.numpy(), ng=np.array([ng0, ng1]),
                 h=np.array([st.h_knn]), ghosts=np.array([st.stats["ghosts_knn"], st.stats["ghosts_fof"]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,halo", [(2, None), (2, 0.01), (3, None), (4, None)])
def test_sharded_host_logic_matches_single_domain(port, world, halo):
    from nbodylib_b200.synth import clustered_small
    pos, vel, mass = clustered_small(5000, seed=77)
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(world, 29500 + np.random.randint(0, 2000), tmp, halo), nprocs=world, join=True)
        res = [np.load(os.path.join(tmp, "r%d.npz" % r)) for r in range(world)]
    rho = np.zeros(len(pos))
    g0 = np.zeros(len(pos), dtype=np.int64)
    g1 = np.zeros(len(pos), dtype=np.int64)
    for r in res:
        rho[r["idx"]] = r["rho"]
        g0[r["idx"]] = r["g0"]
        g1[r["idx"]] = r["g1"]
    ref_rho, _ = port.density(pos, mass, 16)
    np.testing.assert_allclose(rho, ref_rho, rtol=1e-12)
    if halo is not None:                       # a deliberately narrow halo must have been widened
        assert all(float(r["h"][0]) > halo for r in res)
    ll = 0.3 / 5000 ** (1 / 3)
    ref_g, ref_ng = port.fof(pos, None, 0, [ll * ll], np.ones(3), 5, 1)
    assert int(res[0]["ng"][0]) == int(res[0]["ng"][1]) == ref_ng
    assert np.array_equal(canon(g0), canon(ref_g)) and np.array_equal(canon(g1), canon(ref_g))
    assert g0.max() == ref_ng and g1.max() == ref_ng
    sizes = np.bincount(g1)[1:]
    assert np.all(np.diff(sizes) <= 0) and np.array_equal(np.sort(sizes), np.sort(np.bincount(ref_g)[1:]))
    assert all(int(r["ghosts"][0]) > 0 and int(r["ghosts"][1]) >= 0 for r in res)
