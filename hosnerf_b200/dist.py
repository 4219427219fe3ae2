"""Multi-GPU plumbing for the per-ray path: rays are independent, so the hot path shards with no
data-path collective.  One process per GPU (torchrun), torch.distributed for the two
places the reference communicates:

  * eval: every rank renders the rays  i = rank (mod world)  of a padded frame and the results are
    re-interleaved after an all_gather - S1/src/data/sampler.py:44-46 (DDPSequnetialSampler),
    S1/src/data/interface.py:152-156 (padding), S1/src/model/interface.py:30-39 (alter_gather_cat).
  * training (next round): one flat all-reduce of the gradient buffer per step.

Works with any backend (NCCL on the B200 box, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def pad_to_world(n_rays: int, world_size: int) -> int:
    """Number of rays after the reference's padding to a multiple of the world size."""
    return (n_rays + world_size - 1) // world_size * world_size


def shard_indices(n_rays: int, rank: int, world_size: int, mode: str = "strided") -> torch.Tensor:
    """Ray indices rendered by ``rank``.  'strided' = reference (ray i -> rank i mod W, padded by
    repeating the last ray); 'contiguous' = equal blocks (used for whole-frame inference)."""
    total = pad_to_world(n_rays, world_size)
    if mode == "strided":
        idx = torch.arange(rank, total, world_size)
    elif mode == "contiguous":
        per = total // world_size
        idx = torch.arange(rank * per, (rank + 1) * per)
    else:
        raise ValueError(mode)
    return idx.clamp(max=n_rays - 1)


def shard_batch(batch: dict, rank: int, world_size: int, mode: str = "strided") -> dict:
    n = batch["rays_o"].shape[0]
    idx = shard_indices(n, rank, world_size, mode).to(batch["rays_o"].device)
    out = {}
    for k, v in batch.items():
        out[k] = v[idx] if isinstance(v, torch.Tensor) and v.dim() >= 1 and v.shape[0] == n else v
    return out


def gather_rays(local: torch.Tensor, n_rays: int, mode: str = "strided") -> torch.Tensor:
    """Inverse of shard_batch for a per-ray result [n_local, C]: all_gather + re-interleave
    (the permute(1,0,2).flatten(0,1) of alter_gather_cat) and drop the padding."""
    rank, w = world()
    if w == 1:
        return local[:n_rays]
    parts = [torch.empty_like(local) for _ in range(w)]
    dist.all_gather(parts, local.contiguous())
    stacked = torch.stack(parts, 0)                       # [W, n_local, C]
    if mode == "strided":
        full = stacked.permute(1, 0, *range(2, stacked.dim())).flatten(0, 1)
    else:
        full = stacked.flatten(0, 1)
    return full[:n_rays]


def allreduce_flat_(tensors, average: bool = True):
    """One collective for a list of gradient tensors (flat bucket), in place."""
    rank, w = world()
    if w == 1 or not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat)
    if average:
        flat /= w
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
