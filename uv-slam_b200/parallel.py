"""Multi-GPU plumbing: one process per GPU, `torch.distributed` for rendezvous and the exchange.

Two ways the path shards:
  * window-parallel - independent windows are split over the ranks, no data-path collective;
  * factor-parallel - ONE window's landmarks are split over the ranks (landmark k -> rank k % N, IMU factors
    and the prior on rank 0); every LM iteration the ranks sum their partial reduced camera systems
    (ONE all-reduce of d^2 + 3d doubles + the 16-double accumulators per window) and, for the candidate cost, the
    accumulators once more.  The C library holds its own NCCL communicator (`uvs_comm_init_nccl`, libnccl bound at run
    time) or takes the reduction as a callback (`uvs_comm_init`), implemented here with
    `torch.distributed.all_reduce` on the library's own CUDA stream (the same callback runs over gloo on host buffers in
    the CPU tests).
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_windows(windows, rank: int, world: int):
    """contiguous, balanced split of a list of windows (window-parallel mode)"""
    n = len(windows)
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    return windows[lo:hi]


def landmark_owner(index, world: int):
    """rank that owns global landmark `index` in the factor-parallel mode (mirrors the device code)"""
    return np.asarray(index) % world


def init_factor_parallel(solver, dist, rank: int, world: int, how: str = "nccl"):
    """Puts `solver` into the factor-parallel mode.  how = "nccl": the library's own NCCL communicator - rank 0 creates the
    ncclUniqueId (uvs_comm_unique_id), torch.distributed only carries its 128 bytes to the other ranks, every rank calls
    uvs_comm_init_nccl; how = "callback": the all-reduce as a callback into torch.distributed (uvs_comm_init)."""
    if how == "callback":
        solver.comm_init(rank, world, make_allreduce(dist, "cuda"))
        return
    box = [solver.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    solver.comm_init_nccl(box[0], rank, world)


def make_allreduce(dist, device: str = "cuda"):
    """reduce_fn(ptr, count, stream) -> 0 for Solver.comm_init: sums `count` doubles at `ptr` over all ranks,
    in place, ordered on `stream` (a CUDA stream handle; ignored for host buffers)."""
    import torch

    if device == "cuda":
        class _DevPtr:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}

        def reduce_fn(ptr, count, stream):
            t = torch.as_tensor(_DevPtr(ptr, count), device="cuda")
            with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return 0
    else:
        def reduce_fn(ptr, count, stream):
            buf = (C.c_double * count).from_address(ptr)
            t = torch.frombuffer(buf, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return 0
    return reduce_fn
