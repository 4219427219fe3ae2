"""world_size-2 gloo tests (CPU) of the ray sharding / gather plumbing used by bench.py --gpus N
and whole-frame inference."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hosnerf_b200 import dist as hd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rays, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        batch = {"rays_o": torch.randn(n_rays, 3, generator=g), "rays_d": torch.randn(n_rays, 3, generator=g),
                 "radii": torch.rand(n_rays, 1, generator=g), "times": torch.tensor(0.0)}
        local = hd.shard_batch(batch, rank, world, mode)
        assert local["rays_o"].shape[0] == hd.pad_to_world(n_rays, world) // world
        assert local["times"].dim() == 0
        rendered = local["rays_o"] * 2.0 + local["rays_d"]          # stand-in for a per-ray result
        full = hd.gather_rays(rendered, n_rays, mode)
        ok = torch.equal(full, batch["rays_o"] * 2.0 + batch["rays_d"])
        grads = [torch.full((5,), float(rank + 1)), torch.full((2, 3), float(10 * (rank + 1)))]
        hd.allreduce_flat_(grads, average=True)
        ok = ok and torch.allclose(grads[0], torch.full((5,), 1.5)) and torch.allclose(grads[1], torch.full((2, 3), 15.0))
        # FlatGrads: param.grad are views of one flat buffer, buckets reduced asynchronously, averaged in finish()
        net = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Linear(4, 2))
        fg = hd.FlatGrads(net, bucket_of=lambda name: int(name.split(".")[0]))
        assert fg.flat.numel() == sum(p.numel() for p in net.parameters()) and sorted(fg.ranges) == [0, 1]
        for name, p in net.named_parameters():
            assert p.grad.data_ptr() == fg.views[name].data_ptr()
            fg.add_(name, torch.full_like(p, float(rank + 1)))
        fg.reduce_bucket(1)                    # one bucket early (overlap), the other picked up by finish()
        fg.finish()
        ok = ok and all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in net.parameters())
        fg.zero_()
        ok = ok and float(fg.flat.abs().max()) == 0.0 and not fg.reduced
        # deferred mode (a backward replayed from a CUDA graph): reduce_bucket() calls from inside the backward are ignored,
        # finish() reduces everything, and the object is ready for the next step without zero_() having run on the host
        fg.defer = True
        for step in range(2):
            fg.flat.zero_()                    # what the replayed graph does (no host-side bookkeeping)
            for name, p in net.named_parameters():
                fg.add_(name, torch.full_like(p, float(rank + 1 + step)))
            fg.reduce_bucket(0)
            ok = ok and not fg.works and not fg.reduced
            fg.finish()
            ok = ok and all(torch.allclose(p.grad, torch.full_like(p, 1.5 + step)) for p in net.parameters()) and not fg.reduced
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rays,mode", [(10, "strided"), (11, "strided"), (11, "contiguous")])
def test_shard_gather_world2(n_rays, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rays, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_eval_tail_gather_and_psnr_single_process():
    """LitMipNeRF360.alter_gather_cat / psnr_each (S1 interface.py:28-51) on one process: images are cut out of the
    concatenated per-step outputs in order; PSNR of a known error."""
    from hosnerf_b200 import LitMipNeRF360
    lit = LitMipNeRF360.__new__(LitMipNeRF360)
    sizes = [(2, 3), (1, 4)]
    flat = torch.arange(10 * 3, dtype=torch.float32).reshape(10, 3) / 30.0
    outs = [{"rgb": flat[:4]}, {"rgb": flat[4:]}]
    imgs = LitMipNeRF360.alter_gather_cat(lit, outs, "rgb", sizes)
    assert [tuple(i.shape) for i in imgs] == [(2, 3, 3), (1, 4, 3)]
    assert torch.equal(torch.cat([i.reshape(-1, 3) for i in imgs]), flat)
    ps = LitMipNeRF360.psnr_each(lit, [torch.full((2, 2, 3), 0.5)], [torch.full((2, 2, 3), 0.6)])
    assert abs(float(ps[0]) - 20.0) < 1e-4


def test_single_process_passthrough():
    x = torch.arange(12.0).view(4, 3)
    assert torch.equal(hd.gather_rays(x, 4), x)
    assert hd.shard_indices(5, 1, 2).tolist() == [1, 3, 4]          # padded: index 5 -> clamped to the last ray
