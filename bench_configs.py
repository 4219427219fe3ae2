"""Secondary bench.py configurations (BASELINE.json configs[2..4]); `python bench.py --config C3|C4|C5 [--gpus N ...]`.

  C3  stage-2 human-object branch, 6144 rays x 128 samples, forward (``Network.forward``), 1 GPU per rank (weak scaling).
  C4  complete HOSNeRF training chunk: 8192 rays in total, sharded contiguously over the ranks (strong scaling);
      background 64 + 64 proposal + 64 NeRF samples (NeRFMLP 1024 wide, the checkpoint default) + 128 human samples,
      forward + backward with the mean(rgb) surrogate objective, ONE flat fp32 gradient buffer over both modules
      (74.2 M parameters, 297 MB) all-reduced over NCCL, bucketed (human | NeRF MLP | proposal MLPs) and asynchronous.
  C5  free-view inference of one 1920 x 1080 frame (2.07 M rays): contiguous ray shards over the ranks (strong scaling),
      rgb shards exchanged with one all_gather per frame; frames/s.

Every function returns the JSON line (dict) rank 0 prints; timing = CUDA events, max over ranks.
"""
from __future__ import annotations

import math
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def _ctx():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local, dev


def _max_over_ranks(vals, dev, world):
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def _event_ms(fn, K, W, world, flush=None):
    for _ in range(W):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        if flush is not None:
            flush.zero_()
        ev[i][0].record()
        fn()
        ev[i][1].record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev)


FLOP_NR, FLOP_CNL = 2 * 101120, 2 * 524800          # non-rigid / canonical MLP FLOP per sample


# ----------------------------------------------------------------------------------------------------------------- C3
def run_c3(args, peaks, clock_sampler):
    from hosnerf_b200 import Network, _lib, default_cfg, synth
    world, rank, local, dev = _ctx()
    n, S = 6144, 128
    K, W = args.steps, max(args.warmup, 3)
    net = Network(default_cfg(), stage2=True, precision="fp16")
    synth.fill_params_(net, 0)
    synth.boost_human_density_(net)
    net = net.to(dev)
    host = synth.make_human_batch(n, ray_seed=2 + rank)
    hb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        with torch.no_grad():
            return net(**hb, cycle_outputs=False)["rgb"]
    sampler = clock_sampler(local)
    sampler.start()
    _lib.LAUNCHES = 0
    ms = _event_ms(step, K, W, world, flush)
    launches = _lib.LAUNCHES - 0
    # end to end: pinned host rays in, rgb out, every step
    pin = {k: host[k].contiguous().pin_memory() for k in ("rays", "near", "far")}

    def e2e():
        kw = dict(hb)
        kw.update({k: v.to(dev, non_blocking=True) for k, v in pin.items()})
        with torch.no_grad():
            return net(**kw, cycle_outputs=False)["rgb"].cpu()
    for _ in range(3):
        e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        out = e2e()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms, e2e_ms = _max_over_ranks([ms, e2e_ms], dev, world)
    if rank != 0:
        return None
    hbm, tf_burst, tf_sus, src = peaks()
    flops = n * S * (FLOP_NR + FLOP_CNL)
    line = {"metric": "ray_samples_per_s", "value": n * S * K * world / (ms * 1e-3), "unit": "ray-samples/s",
            "rays_per_s": n * K * world / (ms * 1e-3), "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "C3: stage-2 human-object Network.forward, 6144 rays x 128 samples (LBS warp + non-rigid MLP + canonical "
                                   "MLP + S2 composite), random SMPL pose",
                       "rays_per_step_per_gpu": n, "samples_per_ray": S, "l2": "256 MiB flush between timed steps",
                       "note": "cycle side path off (render loop); the reference evaluates it in eval too - the reference arm includes it"},
            "e2e": {"value": n * S * K * world / (e2e_ms * 1e-3), "unit": "ray-samples/s", "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": sum(v.numel() * 4 for v in pin.values()), "d2h_bytes_per_step": out.numel() * 4,
                    "api": "Network.forward(rays from pinned host memory ...)['rgb'].cpu()"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "whole step vs the MLP FLOPs (non-rigid 6x128 + canonical 8x256, 2 fused launches)",
                         "achieved": flops * K / (ms * 1e-3) / 1e12, "peak": tf_burst, "unit": "TFLOP/s",
                         "frac": flops * K / (ms * 1e-3) / 1e12 / tf_burst, "peak_source": f"{src} bf16_tflops (burst)", "traffic": None},
            "clocks": sampler.summary()}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = c3_cpu(512)
    return line


def c3_cpu(n):
    """The reference algorithm of the human branch on the host CPU: the oracle port (oracle/human_ref.py, pinned to the
    reference's fixtures) on a bounded sample of the C3 rays."""
    from hosnerf_b200 import Network, default_cfg, synth
    from oracle import human_ref as HR
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    net = Network(default_cfg(), stage2=True)
    synth.fill_params_(net, 0)
    synth.boost_human_density_(net)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    b = synth.make_human_batch(n)
    ts = []
    with torch.no_grad():
        for i in range(3):
            t0 = time.perf_counter()
            HR.network_forward(sd, b, stage2=True)
            ts.append(time.perf_counter() - t0)
    sec = min(ts[1:])
    return {"value": n * 128 / sec, "unit": "ray-samples/s", "cores": threads, "kind": "port",
            "sample": f"{n} of 6144 rays x 128 samples, best of 2 after 1 warm-up, torch CPU fp32 oracle port (incl. the cycle side path)"}


# ----------------------------------------------------------------------------------------------------------------- C4
def run_c4(args, peaks, clock_sampler):
    from hosnerf_b200 import MipNeRF360, Network, _lib, default_cfg, synth, train_hosnerf_chunk
    from hosnerf_b200.dist import FlatGrads
    world, rank, local, dev = _ctx()
    n_total = 8192
    n = n_total // world
    K, W = max(3, min(args.steps, 20)), max(args.warmup, 3)
    bkg = MipNeRF360("/nonexistent", num_prop_samples=64, num_nerf_samples=64, opaque_background=True, stage3=True)
    synth.fill_params_(bkg, 0)
    bkg = bkg.to(dev)
    human = Network(default_cfg())
    synth.fill_params_(human, 0)
    synth.boost_human_density_(human)
    human = human.to(dev)
    human.static_shapes = True          # static-shape training forward: the whole chunk is captured in ONE CUDA graph below
    hb_all = synth.make_human_batch(n_total)
    sl = slice(rank * n, (rank + 1) * n)
    hb = dict(hb_all)
    hb["rays"], hb["near"], hb["far"] = hb_all["rays"][:, sl].contiguous(), hb_all["near"][sl].contiguous(), hb_all["far"][sl].contiguous()
    hb["is_train"] = True
    hb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in hb.items()}
    for k in ("time", "iter_val"):        # host scalars (the state index / Hann window they select are frozen into the graph)
        if isinstance(hb.get(k), torch.Tensor):
            hb[k] = float(hb[k].reshape(-1)[0])
    Mw = synth.random_rigid()
    ro, rd = hb_all["rays"][0][sl], hb_all["rays"][1][sl]
    ro_w = (Mw[:3, :3] @ ro.T).T + Mw[:3, 3]
    rd_w = (Mw[:3, :3] @ rd.T).T
    bb = {"rays_o": ro_w, "rays_d": rd_w, "viewdirs": rd_w / rd_w.norm(dim=-1, keepdim=True), "radii": torch.full((n, 1), 1e-3)}
    bb = {k: v.to(dev).contiguous() for k, v in bb.items()}
    bb["times"] = torch.tensor(0.0)       # stays on the host
    Mw = Mw.to(dev)

    class Both(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model, self.human = bkg, human          # the reference's attribute names / checkpoint prefixes (S3 LitMipNeRF360)
    both = Both()

    def bucket(name):           # human first (its backward finishes first), then the NeRF MLP, then the proposal MLPs
        if name.startswith("human."):
            return 0
        return 1 if name.startswith(f"model.mlps.{bkg.num_levels - 1}.") else 2
    sink = FlatGrads(both, bucket_of=bucket)
    opt = torch.optim.Adam(both.parameters(), lr=1e-5, fused=True)
    exposed = []
    from hosnerf_b200.train import GraphedStep

    def fwd_bwd():
        sink.zero_()
        out = train_hosnerf_chunk(bkg, human, bb, hb, Mw, randomized=False, dense=True)
        loss = out["rgb"].mean()
        loss.backward()
        return loss.detach()
    graphed = GraphedStep(fwd_bwd, warmup=2)

    # zero + forward + objective + backward: one graph replay; then the flat gradient buffer is all-reduced (3 buckets) and
    # the fused Adam step runs - both enqueued eagerly
    def step(fb):
        loss = fb()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sink.reduce_all()
        sink.finish()
        e1.record()
        exposed.append((e0, e1))
        opt.step()
        return loss

    def measure(fb, k):
        for _ in range(W):
            step(fb)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        exposed.clear()
        l0 = _lib.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(k):
            loss = step(fb)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        return e0.elapsed_time(e1) / k, sum(a.elapsed_time(b) for a, b in exposed) / k, wall / k, _lib.LAUNCHES - l0, loss
    sampler = clock_sampler(local)
    sampler.start()
    eager_ms, _, _, _, _ = measure(fwd_bwd, max(3, K // 2))       # the same step enqueued launch by launch (host-bound at this size)
    graph_error = None
    try:
        ms, exp_ms, wall, launches, loss = measure(graphed, K)
    except Exception as e:            # capture refused on this box: report the eager step and say so
        graph_error = f"{type(e).__name__}: {e}"[:300]
        ms, exp_ms, wall, launches, loss = measure(fwd_bwd, K)
    ms, wall = ms * K, wall * K
    ar = 0.0
    if world > 1:
        for _ in range(2):
            dist.all_reduce(sink.flat)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            dist.all_reduce(sink.flat)
        b.record()
        torch.cuda.synchronize()
        ar = a.elapsed_time(b) / 5
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms, exp_ms, ar, wall, eager_ms = _max_over_ranks([ms, exp_ms, ar, wall, eager_ms], dev, world)
    if rank != 0:
        return None
    hbm, tf_burst, tf_sus, src = peaks()
    # forward MLP FLOPs per ray: 2 x 64 proposal samples + 64 NeRF-1024 samples + 128 human samples; the backward
    # adds dgrad + wgrad for the NeRF and human MLPs (the proposal MLPs get no gradient from this objective)
    f_prop, f_nerf, f_h = 2 * 64 * 2 * 342272, 64 * 2 * 8803072, 128 * (FLOP_NR + FLOP_CNL)
    flops = n_total * (f_prop + 3 * (f_nerf + f_h))
    return {"metric": "rays_per_s", "value": n_total * K / (ms * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "wall_ms_per_step": wall / K, "eager_ms_per_step": eager_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic", "loss": float(loss.detach()),
            "config": {"workload": "C4: complete HOSNeRF training chunk, 8192 rays x (64 + 64 proposal, 64 NeRF-1024w, 128 human) samples, "
                                   "forward + backward (mean(rgb) surrogate) + flat gradient all-reduce + Adam; zero / forward / "
                                   "backward replayed from ONE CUDA graph per step (train.GraphedStep; static-shape forward: "
                                   "Network.static_shapes, train_hosnerf_chunk(dense=True)), eager_ms_per_step = the same step "
                                   "enqueued launch by launch",
                       "rays_per_step_total": n_total, "rays_per_gpu": n,
                       "note": "fp16 operands / fp32 accumulate (BASELINE says bf16: same tensor-core rate, the kernels are kind::f16)"},
            "gpu_launches": launches, "graph_error": graph_error,
            "collective": {"op": "ncclAllReduce(sum) over one flat fp32 buffer, 3 buckets", "bytes": sink.nbytes, "exposed_ms_per_step": exp_ms,
                           "standalone_ms": ar, "bus_gbs": (2 * (world - 1) / world) * sink.nbytes / (ar * 1e-3) / 1e9 if ar else None},
            "roofline": {"bound": "tensor", "kernel": "whole step vs algorithmic MLP FLOPs (fwd + dgrad + wgrad)", "achieved": flops * K / (ms * 1e-3) / 1e12 / world,
                         "peak": tf_burst, "unit": "TFLOP/s per GPU", "frac": flops * K / (ms * 1e-3) / 1e12 / world / tf_burst,
                         "peak_source": f"{src} bf16_tflops (burst)", "traffic": None},
            "clocks": sampler.summary()}


# ----------------------------------------------------------------------------------------------------------------- C5
def run_c5(args, peaks, clock_sampler):
    from hosnerf_b200 import MipNeRF360, Network, _lib, camera, default_cfg, ops, synth
    world, rank, local, dev = _ctx()
    H, W_img = (int(x) for x in os.environ.get("HOSNERF_C5_RES", "1080x1920").split("x"))
    CHUNK = 65536
    K = max(1, min(args.steps, 3))
    bkg = MipNeRF360("/nonexistent", num_prop_samples=64, num_nerf_samples=64, opaque_background=True, stage3=True, precision="fp16")
    synth.fill_params_(bkg, 0)
    bkg = bkg.to(dev)
    human = Network(default_cfg(), stage2=False, precision="fp16")
    synth.fill_params_(human, 0)
    synth.boost_human_density_(human)
    human = human.to(dev)
    hb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.make_human_batch(8).items()}
    M = synth.random_rigid()
    f = 1500.0 * H / 1080.0
    Kmat = np.array([[f, 0.0, W_img / 2.0], [0.0, f, H / 2.0], [0.0, 0.0, 1.0]])
    R, T = np.eye(3), np.array([0.0, 0.0, 3.0])
    sk = synth.make_skeleton(0)
    bmin = sk["cnl_bbox_min_xyz"].double().numpy()
    bmax = bmin + 2.0 / sk["cnl_bbox_scale_xyz"].double().numpy()
    n = H * W_img
    per = (n + world - 1) // world
    lo, hi = rank * per, min(n, (rank + 1) * per)
    Mr, Mt = M[:3, :3].to(dev), M[:3, 3].to(dev)
    radius = float(np.linalg.norm(M[:3, 0].numpy()) / f * 2 / np.sqrt(12))

    def render_frame():
        o_h, d_h = camera.get_rays_from_KRT(H, W_img, Kmat, R, T)        # all rays on the device (1.6 ms), this rank keeps [lo, hi)
        o_h, d_h = o_h.view(-1, 3)[lo:hi], d_h.view(-1, 3)[lo:hi]
        o_w, d_w = o_h @ Mr.T + Mt, d_h @ Mr.T
        viewdirs = d_w / d_w.norm(dim=-1, keepdim=True)
        m_tot = hi - lo
        rgb = torch.zeros(per, 3, device=dev)
        n_hit = 0
        for c0 in range(0, m_tot, CHUNK):
            c1 = min(m_tot, c0 + CHUNK)
            m = c1 - c0
            bb = {"rays_o": o_w[c0:c1].contiguous(), "rays_d": d_w[c0:c1].contiguous(), "viewdirs": viewdirs[c0:c1].contiguous(),
                  "radii": torch.full((m, 1), radius, device=dev), "times": hb["time"]}
            _, hist = bkg(bb, 1.0, False, False, 0.1, 1e6)
            h = hist[-1]
            oc, dc = o_h[c0:c1].contiguous(), d_h[c0:c1].clone()
            near, far, hit = camera.rays_intersect_3d_bbox(np.stack([bmin, bmax]), oc, dc)
            S_h = human.cfg.N_samples
            h_rgb, h_den = torch.zeros(m, S_h, 3, device=dev), torch.zeros(m, S_h, device=dev)
            h_msk, h_pts = torch.zeros(m, S_h, device=dev), torch.zeros(m, S_h, 3, device=dev)
            k = int(hit.sum())
            n_hit += k
            if k > 0:
                kw = dict(hb)
                kw.update(rays=torch.stack([oc[hit], dc[hit]], 0), near=near[:, None], far=far[:, None])
                out = human(**kw, cycle_outputs=False)
                h_rgb[hit], h_den[hit] = out["human_rgb"].reshape(k, S_h, 3), out["human_density"].reshape(k, S_h)
                h_msk[hit], h_pts[hit] = out["pts_mask"].reshape(k, S_h), out["newsmpl_pts"].reshape(k, S_h, 3)
            rgb[c0:c1], _, _ = ops.composite_s3(h["rgb"].contiguous(), h["density"].contiguous(), h["tdist"].contiguous(), h_rgb, h_den,
                                                h_msk, h_pts, M, bb["rays_o"], bb["rays_d"], want_human_w=False)
        if world > 1:       # the one exchange of the path: every rank ends up with the whole frame (S1 interface.py:30-39)
            parts = [torch.empty_like(rgb) for _ in range(world)]
            dist.all_gather(parts, rgb)
            frame = torch.cat(parts, 0)[:n]
        else:
            frame = rgb[:n]
        return frame, n_hit
    sampler = clock_sampler(local)
    sampler.start()
    with torch.no_grad():
        Hs = H
        H = max(2, min(H, (CHUNK * world) // W_img))        # warm-up on a strip (weight packing, lazy inits, NCCL channels)
        n, per = H * W_img, (H * W_img + world - 1) // world
        lo, hi = rank * per, min(n, (rank + 1) * per)
        render_frame()
        H = Hs
        n, per = H * W_img, (H * W_img + world - 1) // world
        lo, hi = rank * per, min(n, (rank + 1) * per)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = _lib.LAUNCHES
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        t0 = time.perf_counter()
        for i in range(K):
            ev[i][0].record()
            frame, n_hit = render_frame()
            ev[i][1].record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
    launches = _lib.LAUNCHES - l0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    hits = torch.tensor([float(n_hit)], device=dev)
    if world > 1:
        dist.all_reduce(hits)
    ms, wall = _max_over_ranks([ms, wall], dev, world)
    if rank != 0:
        return None
    return {"metric": "frames_per_s", "value": K / (ms * 1e-3), "unit": "frames/s", "rays_per_s": n * K / (ms * 1e-3), "n_gpus": world, "steps": K,
            "warmup": 1, "ms_per_step": ms / K, "wall_ms_per_step": wall / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": f"C5: free-view frame {H}x{W_img} = {n} rays, background 64 + 64 + 64 samples (NeRFMLP 1024 wide) on every ray, "
                                   "human branch (128 samples) on the rays that hit the canonical box, stage-3 depth-merge composite; forward only",
                       "rays_hit_human_box": int(hits[0]), "sharding": "contiguous ray ranges per rank, one all_gather of rgb per frame",
                       "chunk": CHUNK},
            "collective": {"op": "ncclAllGather of [rays/world, 3] fp32 per frame", "bytes_per_rank": per * 12},
            "gpu_launches": launches, "frame_finite": bool(torch.isfinite(frame).all()), "clocks": sampler.summary()}
