"""Multi-GPU plumbing: one process per GPU, cells sharded by contiguous row blocks, `torch.distributed` (NCCL over
NVLink) for the only cross-rank state of the RUV-III fit: replicate-group sums + counts, batch sums, the
NA/Inf flag and the n x n Gram partial.  Everything else is rank-local.

Two transports for those sums:
  * `init_builtin_nccl(ctx)`: the library's own NCCL communicator (dlopen'd libnccl, scm_ctx_comm_init_nccl); torch.distributed
    is only used once, to hand rank 0's 128-byte unique id to the other ranks.  This is what bench.py runs.
  * `torch_allreduce_hook`: the host-supplied sum all-reduce (`scm_ctx_set_allreduce`) with a raw buffer pointer, wrapped
    zero-copy as a torch tensor.  The same hook works on host memory with the `gloo` backend, which is how the CPU tests
    exercise the host-side logic.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_rows(m: int, world: int, rank: int):
    """Contiguous block of ceil(m / world) cells per rank (the last ranks may be short or empty)."""
    per = -(-m // world)
    lo = min(m, rank * per)
    return lo, min(m, lo + per)


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier for a raw device pointer."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def torch_allreduce_hook(device: int | None):
    """Returns fn(ptr, count, dtype, stream) that sums the buffer in place over the default process group.
    device = CUDA ordinal for NCCL, or None for host buffers (gloo)."""
    import torch
    import torch.distributed as dist

    def fn(ptr, count, dtype, stream):
        del stream  # the context was created on torch's current stream, so ordering is already right
        np_dt = np.float64 if dtype == 0 else np.int64
        if device is None:
            buf = (C.c_char * (int(count) * 8)).from_address(int(ptr))
            t = torch.from_numpy(np.frombuffer(buf, dtype=np_dt, count=int(count)))
        else:
            t = torch.as_tensor(_RawCuda(ptr, count, "<f8" if dtype == 0 else "<i8"), device=torch.device("cuda", device))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    return fn


def init_builtin_nccl(ctx, rank: int | None = None, world: int | None = None):
    """Create the library's own NCCL communicator for `ctx` over the default torch.distributed process group."""
    import torch.distributed as dist
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world

    def bcast(b):
        box = [b]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    ctx.init_nccl(rank, world, bcast)
    return ctx
