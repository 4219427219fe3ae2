"""Event-timed LBS warp / forward warp at the C3 point count (6144 x 128 points, 24 bones, 32^3 volume)."""
import sys, torch
sys.path.insert(0, '.')
from hosnerf_b200 import ops, synth, Network, default_cfg
dev = "cuda:0"
hb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.make_human_batch(6144).items()}
hn = Network(default_cfg(), stage2=True, precision="fp16"); synth.fill_params_(hn, 0); hn = hn.to(dev)
with torch.no_grad():
    out = hn(**hb, cycle_outputs=False)
    fr = hn._cache["frame"]
    print({k: (tuple(v.shape) if isinstance(v, torch.Tensor) else type(v).__name__) for k, v in fr.items()})
    n, S = 6144, hn.cfg.N_samples
    t_lin = torch.linspace(0, 1, S, device=dev)
    rays_o, rays_d = hb["rays"][0].float().contiguous(), hb["rays"][1].float().contiguous()
    z, pts = ops.human_samples(rays_o, rays_d, hb["near"].reshape(-1).float().contiguous(), hb["far"].reshape(-1).float().contiguous(), t_lin, None)
    flat = pts.view(-1, 3)
    Rb, Tb, vol = fr["Rb"][0].contiguous(), fr["Tb"][0].contiguous(), fr["vol"].contiguous()
    def timeit(fn, name, k=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / k * 1e3:.1f} us for {flat.shape[0]} points x {Rb.shape[0]} bones")
    timeit(lambda: ops.lbs_warp(flat, Rb, Tb, vol, fr["bbox_min"], fr["bbox_scale"]), "lbs_warp")
    x, m = ops.lbs_warp(flat, Rb, Tb, vol, fr["bbox_min"], fr["bbox_scale"])
    print("mask>0 fraction", float((m > 0).float().mean()), "mask>0.005", float((m > 0.005).float().mean()))
    timeit(lambda: ops.lbs_forward(x, fr["Rf"][0].contiguous(), fr["Tf"][0].contiguous(), vol, fr["bbox_min"], fr["bbox_scale"]), "lbs_forward")
