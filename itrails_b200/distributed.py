"""Multi-GPU plumbing: one process per GPU, blocks sharded across ranks.

Independent MAF blocks are independent chains (optimizer.py:112-113, 260-261,
374-376), so the path shards with no data-path collective.  The only exchange is
the sum of per-rank log-likelihood partials (one FP64 per parameter set), done with
``torch.distributed.all_reduce`` (NCCL over NVLink on GPUs, gloo in CPU tests).
Viterbi / posterior outputs stay sharded.
"""
from __future__ import annotations

import numpy as np


def lpt_partition(lengths, world_size):
    """Greedy longest-processing-time assignment of blocks to ranks by column count.
    Returns a list (per rank) of block-index arrays, each in ascending block order.
    Deterministic: ties go to the lowest rank."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    parts = [[] for _ in range(world_size)]
    for b in order:
        r = int(np.argmin(load))
        parts[r].append(int(b))
        load[r] += lengths[b]
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def is_active():
    # A process group can only exist if the caller has imported torch.distributed already;
    # never import torch from here (1.7 s of start-up for every single-GPU CLI run).
    import sys
    dist = sys.modules.get("torch.distributed")
    if dist is None:
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank_world():
    if not is_active():
        return 0, 1
    import torch.distributed as dist
    return dist.get_rank(), dist.get_world_size()


def allreduce_sum(values, device=None):
    """Sum a small float64 vector over all ranks (identity when not distributed)."""
    values = np.atleast_1d(np.asarray(values, dtype=np.float64))
    if not is_active():
        return values
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    dev = torch.device("cuda", device if device is not None else torch.cuda.current_device()) \
        if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(values.copy()).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allreduce_max(value):
    if not is_active():
        return float(value)
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_blocks(V_lst):
    """This rank's blocks under the LPT partition: (local V_lst, global block ids)."""
    rank, world = rank_world()
    if world == 1:
        return list(V_lst), np.arange(len(V_lst), dtype=np.int64)
    ids = lpt_partition([len(v) for v in V_lst], world)[rank]
    return [V_lst[i] for i in ids], ids
