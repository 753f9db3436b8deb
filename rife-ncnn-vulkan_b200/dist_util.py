"""Multi-process plumbing (one process per GPU): contiguous sharding of the frame-pair stream and the one collective
the path has -- rank 0 reads the model directory and broadcasts the packed model.
Backend-agnostic (NCCL on the GPU box, gloo in the CPU tests)."""
import numpy as np


def shard_pairs(n_pairs, world, rank):
    """Contiguous chunk [lo, hi) of the pair list for `rank`; chunks differ by at most one pair and cover the list.
    Chunk boundaries duplicate one source frame (pair i uses frames i and i+1), as the reference's shared queue would."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def broadcast_blob(blob, dist, device=None, src=0):
    """Broadcast a uint8 numpy blob held by rank `src`; every rank returns it as a numpy array.
    `dist` = torch.distributed (initialised); `device` = torch device for the collective (cuda for NCCL, cpu for gloo)."""
    import torch
    rank = dist.get_rank()
    dev = device if device is not None else torch.device("cpu")
    n = torch.tensor([0 if blob is None else int(blob.size)], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    if rank == src:
        t = torch.from_numpy(np.ascontiguousarray(blob, dtype=np.uint8)).to(dev)
    else:
        t = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src)
    return t.cpu().numpy()


def max_over_ranks(value, dist, device=None):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
