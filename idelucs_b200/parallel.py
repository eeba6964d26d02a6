"""Multi-GPU sharding of the hot path: one process per GPU (torch.distributed, NCCL over
NVLink on the B200 box; gloo in the CPU tests).  Sequences are independent, so they are
partitioned over the ranks with no data-path collective; the only exchanges are
  * the StandardScaler statistics (per-rank (count, mean, M2) partials, all-gathered and merged
    in rank order by idl_scaler_finalize — identical result on every rank),
  * the gradient all-reduce of the data-parallel MLP replicas (torch DDP, NCCL),
  * an all-gather of the cluster assignments / latents at inference.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_ranges(lengths, world_size):
    """Contiguous sequence ranges [lo, hi) per rank, balanced by total bases (featurisation
    cost is ~ proportional to length + a constant per profile).  Deterministic, host-only."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = lengths.size
    cost = np.cumsum(lengths + 4096)          # + per-sequence output cost proxy
    total = int(cost[-1]) if n else 0
    bounds = [0]
    for r in range(1, world_size):
        target = total * r // world_size
        bounds.append(int(np.searchsorted(cost, target, side="left")) if n else 0)
    bounds.append(n)
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def all_gather_rows(x, counts, group=None):
    """All-gather row blocks of different length (rank r contributes counts[r] rows) -> the
    concatenation in rank order, on every rank."""
    world = dist.get_world_size(group)
    m = int(max(counts))
    pad = torch.zeros((m,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[: x.shape[0]] = x
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: int(c)] for o, c in zip(out, counts)], dim=0)


def all_gather_ragged(tensors, group=None):
    """All-gather several row-aligned tensors (same number of rows on a rank, different across ranks): exchanges the row
    counts once, then one padded all-gather per tensor.  Returns the rank-ordered concatenations (inference tail of
    idelucs/models.py:164-171: predictions, probabilities and latent vectors of every sequence on every rank)."""
    world = dist.get_world_size(group)
    dev = tensors[0].device
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([tensors[0].shape[0]], dtype=torch.int64, device=dev), group=group)
    counts = [int(c.item()) for c in counts]
    return [all_gather_rows(t, counts, group) for t in tensors]
