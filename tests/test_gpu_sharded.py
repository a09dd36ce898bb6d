"""Multi-GPU parity (-m gpu, needs >= 2 GPUs, skipped otherwise): the slab-sharded driver with the CUDA engine over
NCCL against the single-GPU tree on the same global particle set."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import canon

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from nbodylib_b200.sharded import ShardedTree
        from nbodylib_b200.synth import clustered_small
        pos, vel, mass = clustered_small(n, seed=5)
        slab = np.minimum((pos[:, 0] * world).astype(int), world - 1)
        mine = np.nonzero(slab == rank)[0]
        dev = torch.device("cuda", rank)
        st = ShardedTree(torch.from_numpy(pos[mine]).to(dev), torch.from_numpy(vel[mine]).to(dev), torch.from_numpy(mass[mine]).to(dev),
                         period=np.ones(3), rank=rank, world=world, box=(1.0, 1.0, 1.0), slab_local=False, knn_k=32)
        rho = st.CalcDensity(32)
        ll = 0.25 / n ** (1 / 3)
        g, ng = st.FOF(ll, 8, 1)
        np.savez(os.path.join(tmp, "r%d.npz" % rank), idx=mine, rho=rho.cpu().numpy(), g=g.cpu().numpy(), ng=np.array([ng]))
        st.close()
    finally:
        dist.destroy_process_group()


def test_sharded_matches_single_gpu(built):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import nbodylib_b200 as nb
    from nbodylib_b200.synth import clustered_small
    n, world = 200000, 2
    pos, vel, mass = clustered_small(n, seed=5)
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(world, 29500 + np.random.randint(0, 2000), tmp, n), nprocs=world, join=True)
        res = [np.load(os.path.join(tmp, "r%d.npz" % r)) for r in range(world)]
    rho = np.zeros(n)
    g = np.zeros(n, dtype=np.int64)
    for r in res:
        rho[r["idx"]] = r["rho"]
        g[r["idx"]] = r["g"]
    ll = 0.25 / n ** (1 / 3)
    with nb.KDTree(pos, vel, mass, Period=np.ones(3)) as t:
        np.testing.assert_allclose(rho, t.CalcDensity(32), rtol=1e-10)
        g1, ng1 = t.FOF(ll, 8, 1)
    assert int(res[0]["ng"][0]) == ng1 and np.array_equal(canon(g), canon(g1))
    assert np.array_equal(np.bincount(g)[1:], np.bincount(g1)[1:])
