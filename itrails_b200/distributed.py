"""Multi-GPU plumbing: blocks sharded across GPUs, results returned in input order.

Independent MAF blocks are independent chains (optimizer.py:112-113, 260-261,
374-376), so the path shards with no data-path collective.  Blocks are assigned to
*parts* by greedy longest-processing-time; a part is one GPU: ``world_size`` processes
(torchrun, one GPU each) times the GPUs each process drives itself (``ngpu``).  The only
exchange on the compute path is the sum of per-part log-likelihood partials (one FP64
per parameter set: ``torch.distributed.all_reduce`` — NCCL over NVLink on GPUs, gloo in
CPU tests).  Decoded paths / posteriors are gathered AFTER the computation so that every
wrapper returns one result per ``V_lst`` entry in input order, as the reference does
(workflow_viterbi.py:688-743, workflow_posterior.py:693-716).
"""
from __future__ import annotations

import numpy as np


def lpt_partition(lengths, n_parts):
    """Greedy longest-processing-time assignment of blocks to parts by column count.
    Returns a list (per part) of block-index arrays, each in ascending block order.
    Deterministic: ties go to the lowest part.  Parts may be empty when there are
    fewer blocks than parts."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(n_parts, dtype=np.int64)
    parts = [[] for _ in range(n_parts)]
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


def _device():
    import torch
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def allreduce_sum(values, device=None):
    """Sum a small float64 vector over all ranks (identity when not distributed)."""
    values = np.atleast_1d(np.asarray(values, dtype=np.float64))
    if not is_active():
        return values
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", device) if (device is not None and dist.get_backend() == "nccl") else _device()
    t = torch.from_numpy(values.copy()).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allreduce_max(value):
    if not is_active():
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if is_active():
        import torch.distributed as dist
        dist.barrier()


def shard_blocks(V_lst, n_local=1, local_index=0):
    """The blocks of part ``rank * n_local + local_index`` under the LPT partition over
    ``world * n_local`` parts: (local V_lst, global block ids ascending).  May be empty."""
    rank, world = rank_world()
    if world * n_local == 1:
        return list(V_lst), np.arange(len(V_lst), dtype=np.int64)
    ids = lpt_partition([len(v) for v in V_lst], world * n_local)[rank * n_local + local_index]
    return [V_lst[i] for i in ids], ids


def allgather_ragged(local, counts, chunk_bytes=1 << 28):
    """Every rank contributes ``local`` (an array whose first axis has ``counts[rank]``
    entries; trailing shape and dtype equal on all ranks) and receives the list of all
    ranks' arrays.  Done as one broadcast per rank in pieces of at most ``chunk_bytes``,
    so no padding to the largest contribution and no pickling.  Identity when not
    distributed."""
    rank, world = rank_world()
    local = np.ascontiguousarray(local)
    if world == 1:
        return [local]
    import torch
    import torch.distributed as dist
    dev = _device()
    tail = local.shape[1:]
    row_bytes = max(1, int(np.prod(tail, dtype=np.int64)) * local.dtype.itemsize)
    rows_per = max(1, chunk_bytes // row_bytes)
    out = []
    for r in range(world):
        n = int(counts[r])
        buf = local if r == rank else np.empty((n,) + tail, dtype=local.dtype)
        if r == rank and buf.shape[0] != n:
            raise ValueError(f"rank {rank} holds {buf.shape[0]} rows, the partition says {n}")
        for i in range(0, n, rows_per):
            piece = buf[i:i + rows_per]
            # (uint16 has no NCCL type: ship raw bytes)
            t = torch.from_numpy(piece.view(np.uint8).reshape(-1) if r == rank
                                 else np.empty(piece.nbytes, dtype=np.uint8)).to(dev)
            dist.broadcast(t, src=r)
            if r != rank:
                piece.view(np.uint8).reshape(-1)[:] = t.cpu().numpy()
        out.append(buf)
    return out


def allgather_object(obj):
    if not is_active():
        return [obj]
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
