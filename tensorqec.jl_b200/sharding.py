"""Shot sharding across GPUs: the replacement of the reference's job farm.

Reference: src/multiprocessing.jl:5-12, 41-52 (`SimpleMultiprocessing.multiprocess_run`: a job queue over
`Distributed` RemoteChannels, whole jobs shipped to worker processes).  Shots are independent and the compiled
schedule is read-only, so here every rank (one process per GPU) holds a replica of the plan and decodes a contiguous
range of the global shot index; the Philox counters are global shot indices, so the sampled errors do not depend on
the number of ranks.  The only exchange is one all-reduce(sum) of four int64 counters at the end.
"""
from __future__ import annotations

from typing import Callable, Iterable, List

import numpy as np


def shard_range(n_shots: int, rank: int, world: int):
    """Contiguous range [lo, hi) of rank `rank`: the first n_shots % world ranks take one extra shot."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside 0..{world - 1}")
    base, rem = divmod(int(n_shots), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def allreduce_counts(counts, device=None) -> np.ndarray:
    """Sum the int64 counters over all ranks of the default torch.distributed group (NCCL on GPUs, gloo on CPU)."""
    dist, _, world = _dist()
    counts = np.asarray(counts, dtype=np.int64)
    if world == 1:
        return counts.copy()
    import torch
    t = torch.from_numpy(counts.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda(device if device is not None else torch.cuda.current_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def library_comm(device: int):
    """This rank's communicator INSIDE libtqec_cuda.so (`tqec_comm_init`): rank 0 makes the NCCL unique id, the default
    torch.distributed group only carries its 128 bytes to the other ranks (a Julia host would use Distributed.jl for
    that).  -> `_cabi.Comm`, or None in a single-process run."""
    dist, rank, world = _dist()
    if world == 1:
        return None
    import torch
    from . import _cabi
    import os
    import sys
    box = [_cabi.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    # ncclCommInitRank prints a "NCCL version ..." banner on the C-level stdout of rank 0; callers (bench.py) promise a
    # clean stdout, so the banner is sent to stderr
    sys.stdout.flush()
    saved = os.dup(1)
    try:
        os.dup2(2, 1)
        comm = _cabi.Comm(world, rank, box[0], device)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    return comm


def sharded_mc_run(mc, n_shots: int, seed: int = 0, chunk: int = 0, comm=None):
    """Run `mc` (a threshold.MonteCarlo bound to this rank's GPU) on this rank's share of `n_shots` and all-reduce
    the counters: inside the fused pipeline when `comm` (from `library_comm`) is given, else through torch.distributed.
    -> (global counts[4], local device ms, (lo, hi))."""
    _, rank, world = _dist()
    lo, hi = shard_range(n_shots, rank, world)
    counts, ms = mc.run(hi - lo, seed, shot_offset=lo, chunk=chunk, comm=comm)
    if comm is not None:
        return counts, ms, (lo, hi)
    return allreduce_counts(counts, mc.device), ms, (lo, hi)


def multiprocess_run(func: Callable, inputs: Iterable) -> List:
    """multiprocessing.jl:41-52 kept for API parity: applies `func` to every input and returns the results in input
    order.  With torch.distributed initialised the inputs are dealt round-robin to the ranks and gathered; without
    it (the reference's single-worker case, test/multiprocessing.jl:3-6) it runs in-process."""
    inputs = list(inputs)
    dist, rank, world = _dist()
    if world == 1:
        return [func(x) for x in inputs]
    mine = {i: func(x) for i, x in enumerate(inputs) if i % world == rank}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    merged = {}
    for g in gathered:
        merged.update(g)
    return [merged[i] for i in range(len(inputs))]
