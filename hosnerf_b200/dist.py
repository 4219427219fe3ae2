"""Multi-GPU plumbing for the per-ray path: rays are independent, so the hot path shards with no
data-path collective.  One process per GPU (torchrun), torch.distributed for the two
places the reference communicates:

  * eval: every rank renders the rays  i = rank (mod world)  of a padded frame and the results are
    re-interleaved after an all_gather - S1/src/data/sampler.py:44-46 (DDPSequnetialSampler),
    S1/src/data/interface.py:152-156 (padding), S1/src/model/interface.py:30-39 (alter_gather_cat).
  * training: the gradients live in ONE flat fp32 buffer (``FlatGrads``; every ``param.grad`` is a view into it) that is
    all-reduced once per step - S1/run.py:141-156 (DDP) - bucketed per MLP so the collective of the level whose backward
    has finished overlaps the backward of the next one.

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


class FlatGrads:
    """All gradients of ``module`` in one flat fp32 buffer, ``param.grad`` being views into it, grouped in buckets
    (``bucket_of(name) -> int``; default: one bucket).  ``reduce_bucket(b)`` starts an asynchronous all-reduce of one bucket
    (on the communication stream, ordered after the kernels already enqueued on the current stream); ``finish()`` waits for
    all of them and averages.  With world size 1 both are no-ops."""

    def __init__(self, module, bucket_of=None):
        named = [(k, p) for k, p in module.named_parameters() if p.requires_grad]
        bucket_of = bucket_of or (lambda name: 0)
        self.order = sorted(range(len(named)), key=lambda i: (bucket_of(named[i][0]), i))
        total = sum(p.numel() for _, p in named)
        dev = named[0][1].device
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.views, self.ranges = {}, {}
        off = 0
        for i in self.order:
            name, p = named[i]
            n = p.numel()
            v = self.flat[off:off + n].view_as(p)
            p.grad = v
            self.views[name] = v
            b = bucket_of(name)
            lo, hi = self.ranges.get(b, (off, off))
            self.ranges[b] = (min(lo, off), off + n)
            off += n
        self.works, self.reduced = [], set()
        self.nbytes = total * 4
        # True: reduce_bucket() calls made from inside backward are ignored and finish() reduces everything - for a backward
        # that is replayed from a CUDA graph (train.GraphedStep), where the collective has to stay outside the capture
        self.defer = False

    def zero_(self):
        self.flat.zero_()
        self.reduced = set()

    def add_(self, name, g):
        self.views[name].add_(g)

    def reduce_bucket(self, b, _from_finish: bool = False):
        _, w = world()
        if w == 1 or b not in self.ranges or b in self.reduced or (self.defer and not _from_finish):
            return
        self.reduced.add(b)
        lo, hi = self.ranges[b]
        self.works.append(dist.all_reduce(self.flat[lo:hi], async_op=True))

    def reduce_all(self):
        for b in sorted(self.ranges):
            self.reduce_bucket(b, _from_finish=True)

    def finish(self, average: bool = True):
        _, w = world()
        self.reduce_all()                      # buckets whose backward produced nothing still take part in the collective
        for wk in self.works:
            wk.wait()
        self.works = []
        self.reduced = set()                   # ready for the next step even when zero_() is replayed from a CUDA graph
        if w > 1 and average:
            self.flat.div_(w)
