"""Data-parallel plumbing: one process per GPU, NCCL over NVLink owned by libpgmvae.so.

New work relative to the reference (single process, single device: run.py:27-31).  The
library holds the NCCL communicator (``pgmvae_comm``) so that gradient / EMA-statistic /
count reductions are enqueued on the library's own stream; torch.distributed (gloo) is used
only as the rendezvous side channel that carries the 128-byte ncclUniqueId and for host-side
barriers.  ``GlooComm`` offers the same interface on host arrays for CPU tests of the
sharding logic.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Tuple

import numpy as np

from pgmvae import _ffi


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous share [lo, hi) of n items owned by `rank` (same rule as VqVAE.fit)."""
    return n * rank // world, n * (rank + 1) // world


class Comm:
    """NCCL communicator inside libpgmvae.so."""

    def __init__(self, ctx: _ffi.Context, rank: int, nranks: int, unique_id: bytes):
        self.ctx, self.rank, self.nranks = ctx, int(rank), int(nranks)
        nccl = _ffi.find_nccl()
        if nccl and "PGMVAE_NCCL_LIB" not in os.environ:
            os.environ["PGMVAE_NCCL_LIB"] = nccl
        self.h = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), 128) if unique_id is not None else None
        _ffi.check(_ffi.lib().pgmvae_comm_create(ctx.h, self.rank, self.nranks, buf, C.byref(self.h)))

    @staticmethod
    def unique_id() -> bytes:
        nccl = _ffi.find_nccl()
        if nccl and "PGMVAE_NCCL_LIB" not in os.environ:
            os.environ["PGMVAE_NCCL_LIB"] = nccl
        buf = C.create_string_buffer(128)
        _ffi.check(_ffi.lib().pgmvae_comm_unique_id(buf))
        return buf.raw

    def _allreduce(self, a: np.ndarray, fn) -> np.ndarray:
        d = _ffi.DeviceArray.from_numpy(self.ctx, a)
        _ffi.check(fn(self.h, d.ptr, a.size, None))
        return d.numpy()

    def allreduce_u64(self, a):
        return self._allreduce(np.ascontiguousarray(a, np.uint64), _ffi.lib().pgmvae_comm_allreduce_u64)

    def allreduce_f64(self, a):
        return self._allreduce(np.ascontiguousarray(a, np.float64), _ffi.lib().pgmvae_comm_allreduce_f64)

    def allreduce_f32(self, a):
        return self._allreduce(np.ascontiguousarray(a, np.float32), _ffi.lib().pgmvae_comm_allreduce_f32)

    def close(self):
        if self.h:
            _ffi.lib().pgmvae_comm_destroy(self.h)
            self.h = C.c_void_p()


class GlooComm:
    """Same reductions on host arrays through torch.distributed (gloo); CPU tests only."""

    def __init__(self):
        import torch.distributed as dist
        self._dist = dist
        self.rank, self.nranks = dist.get_rank(), dist.get_world_size()
        self.h = None

    def _allreduce(self, a, dtype):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a).astype(dtype))
        self._dist.all_reduce(t)
        return t.numpy()

    def allreduce_u64(self, a):
        return self._allreduce(a, np.int64).astype(np.uint64)

    def allreduce_f64(self, a):
        return self._allreduce(a, np.float64)

    def allreduce_f32(self, a):
        return self._allreduce(a, np.float32)


def init_from_env(ctx: _ffi.Context = None):
    """(comm, rank, world) from torchrun's RANK / WORLD_SIZE / MASTER_* environment.
    world == 1 returns (None, 0, 1) and never touches torch."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world <= 1:
        return None, 0, 1
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    ctx = ctx or _ffi.get_context(int(os.environ.get("LOCAL_RANK", rank)))
    # The library's persistent kernels fill the SMs they are given with one statically scheduled CTA each; PGMVAE_COMM_SMS=n
    # holds NCCL to n CTAs and leaves that many SMs out of the compute grids, so that the overlapped exchange has SMs of
    # its own.  Measured on cfg3 (round 2): 117.8 vs 117.5 ms per step at 2 GPUs, 123.1 vs 122.0 ms at 8 with n = 8 vs 0 --
    # the 6 % of compute given up is not won back, so the default is 0 (NCCL and the GEMMs share the machine).
    comm_sms = int(os.environ.get("PGMVAE_COMM_SMS", "0"))
    if comm_sms > 0:
        os.environ.setdefault("NCCL_MAX_CTAS", str(comm_sms))
        os.environ.setdefault("NCCL_MIN_CTAS", str(min(comm_sms, 4)))
    box = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    comm = Comm(ctx, rank, world, box[0])
    ctx.reserve_sms(max(comm_sms, 0))
    return comm, rank, world
