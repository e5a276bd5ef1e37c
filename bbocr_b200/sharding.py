"""Data-parallel page sharding: pages are independent units, so ranks take disjoint page sets and no
collective touches the data path.  `torch.distributed` is used only for the barrier / max-over-ranks timing and for
gathering per-rank result counts."""
from __future__ import annotations


def shard_round_robin(n_items: int, rank: int, world: int):
    """Static round-robin for homogeneous configs (2, 3, 4): item i -> rank i % world."""
    return list(range(rank, n_items, world))


def shard_by_cost(costs, rank: int, world: int):
    """Greedy longest-first bin packing for mixed page sizes (config 5): every rank computes the same assignment
    deterministically from the cost list (pixel counts), so no communication is needed."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += costs[i]
        if r == rank:
            mine.append(i)
    return sorted(mine)


def gather_counts(local_count: int):
    """All-gather of one integer per rank (result accounting only).  Works on gloo and nccl."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [local_count]
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([local_count], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x.item()) for x in out]
