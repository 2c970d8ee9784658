"""Data-parallel plumbing: the minibatch shards by columns, W/A/B are replicated, and the only exchange
is one all-reduce(sum) per step of the packed (k x (k+d)) partial sums [Ht^T Ht | Ht^T Xt]
(SURVEY.md §8e).  Works with any torch.distributed backend (nccl on the B200s, gloo in the CPU tests)."""
from __future__ import annotations

import os


def shard_range(n: int, world: int, rank: int):
    """Contiguous block [lo, hi) of the n minibatch columns owned by `rank` (sizes differ by at most 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank %r/%r" % (world, rank))
    base, rem = divmod(int(n), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def default_group(process_group=None):
    """(group, world, rank) for the reference-facing classes: an explicit group, else the WORLD group when
    torch.distributed is initialised with more than one rank, else single process."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None, 1, 0
    group = process_group if process_group is not None else dist.group.WORLD
    world = dist.get_world_size(group)
    if world <= 1:
        return None, 1, 0
    return group, world, dist.get_rank(group)


def pack_partial(HtH, HtX):
    """[HtH | HtX] -> one (k x (k+d)) buffer (what gets all-reduced)."""
    import torch
    return torch.cat([HtH, HtX], dim=1).contiguous()


def allreduce_packed(P, group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(P, group=group)
    return P


def init_from_env(backend="nccl"):
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, world, local_rank)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local
