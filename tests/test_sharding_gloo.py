"""world_size-2 gloo test of the multi-GPU host logic (sharding + result collection) on CPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, pkg


def test_shard_ranges_partition_the_batch():
    sh = pkg("sharding")
    for n in (1, 7, 1024, 65536, 1000):
        for w in (1, 2, 3, 4, 8):
            r = [sh.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_range(10, 2, 2)


class FakeLocal:
    """Test double standing in for BatchedVSMPC: output row = f(pack column), so that the collected
    result can be checked against a single-process evaluation."""

    def __init__(self, n, params, trajectories, device=0, solver=0):
        self.B = n
        self.pack = None

    def configure_pack(self, pack, jp, phase0=None):
        assert pack.shape[1] == self.B and jp.shape[1] == self.B
        return True

    def update_pack(self, pack):
        assert pack.shape == (359, self.B)
        self.pack = pack
        return True

    def solveMPC(self):
        return True

    def get_output(self):
        out = np.stack([self.pack[k % 359] * (k + 1) for k in range(54)], axis=1)
        return out, (self.pack[0] > 0).astype(np.int32)


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = pkg("sharding")
    rng = np.random.default_rng(5)
    pack = rng.normal(size=(359, n_total))
    jp = rng.normal(size=(8, n_total))
    m = sh.ShardedVSMPC(n_total, None, None, rank, world, factory=FakeLocal)
    m.configure_pack(pack, jp, np.arange(n_total) % 20)
    m.update_pack(pack)
    m.solveMPC()
    out, status = m.get_output_all()
    ref = FakeLocal(n_total, None, None)
    ref.update_pack(pack)
    ro, rs = ref.get_output()
    ok = out.shape == ro.shape and np.array_equal(out, ro) and np.array_equal(status, rs) and (m.hi - m.lo) in (n_total // world, n_total // world + 1)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [64, 37, 1])     # 1: rank 0 owns no instance and only joins the gather
def test_sharded_gather_world_size_2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + n_total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
