"""Multi-GPU checks (need >= 2 visible GPUs; skipped otherwise): handles on different devices in one process, and
the sharded path (one process per GPU, NCCL gather of the output rows only) against the single-GPU batch."""
import os
import sys

import numpy as np
import pytest

from helpers import ROOT, load_trajectories, pkg
from oracle_driver import oracle_trajectories_to_product

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _inputs(B):
    syn, P = pkg("synthetic"), pkg("pack")
    nom = syn.make_states(B, perturbed=False)
    per = syn.make_states(B, seed=77, perturbed=True, near_bound_fraction=0.2)
    jp = np.ascontiguousarray(nom["joint_pos"][:, P.DEFAULT_JOINT_SELECTOR].T)
    return P.build_pack(nom), jp, P.build_pack(per)


def test_handles_on_two_devices_in_one_process():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    bat = pkg("batched")
    B = 64
    nom, jp, per = _inputs(B)
    outs = []
    for solver in (0, 1, 2):
        res = []
        for dev in (0, 1):
            mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(load_trajectories()), device=dev, solver=solver)
            mpc.configure_pack(nom, jp)
            mpc.update_pack(per)
            mpc.solveMPC()
            out, st = mpc.get_output()
            assert (st == 0).all()
            res.append(out)
            mpc.close()
        assert np.array_equal(res[0], res[1])
        outs.append(res[0])
    assert np.abs(outs[0] - outs[2]).max() / np.abs(outs[0]).max() < 1e-9


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sh = pkg("sharding")
    nom, jp, per = _inputs(B)
    m = sh.ShardedVSMPC(B, None, oracle_trajectories_to_product(load_trajectories()), rank, world, device=rank)
    m.configure_pack(nom, jp)
    m.update_pack(per)
    m.solveMPC()
    out, status = m.get_output_all(device=torch.device("cuda", rank))
    if rank == 0:
        q.put((out, status))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_two_gpus_match_single_gpu():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    bat = pkg("batched")
    B = 101                      # uneven shards
    nom, jp, per = _inputs(B)
    mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(load_trajectories()))
    mpc.configure_pack(nom, jp)
    mpc.update_pack(per)
    mpc.solveMPC()
    ref_out, ref_status = mpc.get_output()
    mpc.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, status = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
    assert np.array_equal(out, ref_out) and np.array_equal(status, ref_status)
