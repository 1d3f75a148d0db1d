"""Multi-GPU layout: instances are independent, so a batch of B instances is split into contiguous
ranges, one per rank / GPU (one process per GPU, torch.distributed).  There is NO collective in the
solve; a collective is used only to collect the per-instance output rows (NCCL all_gather on GPUs,
gloo in the CPU tests).  SURVEY.md §8(e)."""
from __future__ import annotations

import numpy as np

from ._lib import OUT_DOUBLES


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous instance range [lo, hi) of `rank`: [g*B/G, (g+1)*B/G)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    lo = (rank * n_total) // world
    hi = ((rank + 1) * n_total) // world
    return lo, hi


def gather_rows(local_rows, n_total: int, group=None):
    """All-gather per-instance rows (torch tensor [n_local, C], any device) into [n_total, C] on every
    rank.  Uneven shards are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    C = local_rows.shape[1]
    pad = torch.zeros((nmax, C), dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


class _DeviceArray:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface v2)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def output_tensors(mpc):
    """Torch views (no copy) of the handle's device-resident output rows [B, 54] (float64) and status [B] (int32),
    from ``vsmpc_get_output_device``; valid until the next solve on the handle's stream."""
    import torch
    rows_ptr, status_ptr = mpc.output_device_ptrs()
    dev = torch.device("cuda", mpc.device)
    rows = torch.as_tensor(_DeviceArray(rows_ptr, (mpc.B, OUT_DOUBLES), "<f8"), device=dev)
    status = torch.as_tensor(_DeviceArray(status_ptr, (mpc.B,), "<i4"), device=dev)
    return rows, status


def gather_output_device(mpc, n_total: int, group=None):
    """Collect every rank's output rows with NCCL straight from the device-resident rows of the handle (no host round
    trip): returns ([n_total, 54] float64, [n_total] int32) device tensors on every rank.  The collective runs on torch's
    current stream, which must be the stream the handle solves on (``mpc.set_stream``)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    rows, status = output_tensors(mpc) if mpc is not None else (None, None)
    even = all(hi - lo == nmax for lo, hi in sizes)
    dev = rows.device if rows is not None else torch.device("cuda", torch.cuda.current_device())
    if even:
        all_rows = torch.empty((n_total, OUT_DOUBLES), dtype=torch.float64, device=dev)
        all_status = torch.empty((n_total,), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_rows, rows, group=group)
        dist.all_gather_into_tensor(all_status, status, group=group)
        return all_rows, all_status
    pad_r = torch.zeros((nmax, OUT_DOUBLES), dtype=torch.float64, device=dev)
    pad_s = torch.zeros((nmax,), dtype=torch.int32, device=dev)
    if rows is not None:
        pad_r[: rows.shape[0]] = rows
        pad_s[: status.shape[0]] = status
    buf_r = torch.empty((world * nmax, OUT_DOUBLES), dtype=torch.float64, device=dev)
    buf_s = torch.empty((world * nmax,), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(buf_r, pad_r, group=group)
    dist.all_gather_into_tensor(buf_s, pad_s, group=group)
    keep = torch.cat([torch.arange(r * nmax, r * nmax + (hi - lo), device=dev) for r, (lo, hi) in enumerate(sizes)])
    return buf_r[keep], buf_s[keep]


class ShardedVSMPC:
    """The reference surface over a batch sharded across ranks.  Every rank passes the FULL SoA pack
    (PACK_DOUBLES, B) — or only needs its own columns to be valid — and owns the instances of its range."""

    def __init__(self, n_total: int, params, trajectories, rank: int, world: int, device: int = 0, solver: int = 0,
                 factory=None):
        self.n_total, self.rank, self.world = int(n_total), int(rank), int(world)
        self.lo, self.hi = shard_range(n_total, rank, world)
        if factory is None:
            from .batched import BatchedVSMPC as factory
        # a batch smaller than the world leaves some ranks without instances: they hold no handle (vsmpc_create rejects
        # an empty batch) and only take part in the collection of the output rows
        self.local = factory(self.hi - self.lo, params, trajectories, device=device, solver=solver) \
            if self.hi > self.lo else None

    def _cols(self, a):
        return np.ascontiguousarray(np.asarray(a)[:, self.lo:self.hi])

    def configure_pack(self, pack, joint_pos_sel, phase0=None):
        if self.local is None:
            return True
        ph = None if phase0 is None else np.ascontiguousarray(np.asarray(phase0)[self.lo:self.hi])
        return self.local.configure_pack(self._cols(pack), self._cols(joint_pos_sel), ph)

    def update_pack(self, pack):
        return True if self.local is None else self.local.update_pack(self._cols(pack))

    def solveMPC(self):
        return True if self.local is None else self.local.solveMPC()

    def get_output_local(self):
        if self.local is None:
            return np.zeros((0, OUT_DOUBLES)), np.zeros(0, dtype=np.int32)
        return self.local.get_output()

    def get_output_all(self, group=None, device=None):
        """Collect the output rows of every rank (the only communication of the whole path)."""
        import torch
        import torch.distributed as dist
        if dist.get_backend(group) == "nccl":
            # GPUs: NCCL all-gather straight from the device-resident rows (vsmpc_get_output_device), one D2H at the end
            if self.local is not None:
                self.local.wait()
            rows, status = gather_output_device(self.local, self.n_total, group)
            return rows.cpu().numpy(), status.cpu().numpy()
        out, status = self.get_output_local()
        dev = device if device is not None else "cpu"
        rows = torch.from_numpy(np.concatenate([out, status[:, None].astype(np.float64)], axis=1)).to(dev)
        full = gather_rows(rows, self.n_total, group).cpu().numpy()
        return full[:, :-1], full[:, -1].astype(np.int32)
