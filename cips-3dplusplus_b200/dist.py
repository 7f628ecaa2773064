"""Multi-GPU plumbing: the path shards by (latent, pose) image with no data-path collective
(gen_images.py:57-91 does the same in the reference).  One process per GPU; NCCL is used only to
all-gather rendered maps and to all-reduce latent/camera gradients in batched inversion."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous split of n_items across ranks; earlier ranks take the remainder."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_maps(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gather per-rank image shards (dim 0) back into the full (n_items, ...) tensor."""
    ws = dist.get_world_size(group)
    if ws == 1:
        return local
    counts = [shard_range(n_items, r, ws) for r in range(ws)]
    mx = max(e - s for s, e in counts)
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = local.new_empty((ws * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * mx: r * mx + (e - s)] for r, (s, e) in enumerate(counts)], 0)


def allreduce_grads(tensors, group=None):
    """One flattened SUM all-reduce for a list of (small) gradient tensors, in place."""
    if dist.get_world_size(group) == 1 or not tensors:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[o:o + n].view_as(t))
        o += n
    return tensors
