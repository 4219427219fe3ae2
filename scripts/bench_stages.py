"""Achieved HBM GB/s (algorithmic bytes / CUDA-event time) of the bandwidth-bound stages at the bench sizes, and
step times of the secondary configurations (C3: human-object branch, 6144 rays x 128 samples; S2 composite).
Writes a markdown table to stdout; profiles/r1_stage_rooflines.md is its output on a B200."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hosnerf_b200 import MipNeRF360, Network, default_cfg, ops, synth

dev = "cuda:0"
peaks = {}
pp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peaks = json.load(open(pp))
HBM = peaks.get("hbm_gbs", 6650.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


rows = []
N, S = 4096, 128
g = torch.Generator().manual_seed(0)
b = {k: v.to(dev) for k, v in synth.make_bkg_batch(N, seed=1).items()}
net = MipNeRF360("/nonexistent", num_levels=2, num_prop_samples=S, num_nerf_samples=S, nerf_netwidth=256,
                 opaque_background=True, precision="fp16")
synth.fill_params_(net, 0)
net = net.to(dev)
with torch.no_grad():
    _, hist = net(b, 1.0, False, False, 0.1, 1e6)
sd0, w0 = hist[0]["sdist"].contiguous(), hist[0]["weights"].contiguous()
u_base, mj = net._u_base(S, False, dev)
ms = timeit(lambda: ops.resample_level(sd0, w0, True, 0.0025 + 0.5 / S, 1.0, 0.0, u_base, None, mj, 0.0, 1.0, 10.0, 1e-6))
byt = N * ((S + 1) + S + 2 * (S + 1)) * 4
rows.append(("resample_level_kernel (level 1: 128 -> 128, dilate + CDF + invert)", f"{N} rays", ms, byt))
dens = hist[1]["density"].contiguous()
rgb = hist[1]["rgb"].contiguous()
td = (1.0 / (hist[1]["sdist"] / 1e6 + (1.0 - hist[1]["sdist"]) / 0.1)).contiguous()
ms = timeit(lambda: ops.composite_mip360(dens, td, b["rays_d"], rgb, True, 1.0))
byt = N * S * 20 + N * S * 4 + N * 12
rows.append(("composite_mip360_kernel (weights + rgb)", f"{N} x {S}", ms, byt))

# ---- human branch (C3)
n_h = 6144
hb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.make_human_batch(n_h).items()}
P = n_h * 128
pts = torch.randn(P, 3, generator=g).to(dev) * 0.5
R = torch.eye(3).repeat(26, 1, 1).to(dev).contiguous()
T = (torch.randn(26, 3, generator=g) * 0.1).to(dev)
vol = torch.rand(27, 32, 32, 32, generator=g).to(dev)
ms = timeit(lambda: ops.lbs_warp(pts, R, T, vol, [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]))
rows.append(("lbs_warp_kernel (26 bones x 8 corners per point)", f"{P} points", ms, P * 28))
gathers = P * 26 * 8
lbs_note = f"{gathers / ms / 1e6:.1f} G gathers/s from the L2-resident 3.4 MiB volume"
raw = torch.randn(n_h, 128, 4, generator=g).to(dev)
mask = torch.rand(n_h, 128, 1, generator=g).to(dev)
z = torch.sort(torch.rand(n_h, 128, generator=g) * 2 + 2, dim=-1).values.to(dev)
dirs = torch.randn(n_h, 3, generator=g).to(dev)
ms = timeit(lambda: ops.composite_nerf(raw, mask, z, dirs, [255.0, 255.0, 255.0], True))
rows.append(("composite_nerf_kernel (S2 _raw2outputs)", f"{n_h} x 128", ms, n_h * 128 * 24 + n_h * 128 * 4 + n_h * 20))

print("| kernel | size | ms | algorithmic bytes | achieved GB/s | fraction of measured HBM peak (%.0f GB/s) |" % HBM)
print("|---|---|---|---|---|---|")
for name, size, ms, byt in rows:
    gbs = byt / ms / 1e6
    print(f"| `{name}` | {size} | {ms:.4f} | {byt / 1e6:.1f} MB | {gbs:.0f} | {100 * gbs / HBM:.1f} % |")
print()
print("LBS:", lbs_note)

# ---- whole human-branch step (C3), both precision modes
for prec in ("fp16", "fp32"):
    hn = Network(default_cfg(), stage2=False, precision=prec)
    synth.fill_params_(hn, 0)
    synth.boost_human_density_(hn)
    hn = hn.to(dev)
    with torch.no_grad():
        ms = timeit(lambda: hn(**hb, cycle_outputs=False), reps=5)
    print(f"C3 human branch (Network.forward, {n_h} rays x 128 samples, {prec}): {ms:.3f} ms/step = {n_h * 128 / ms / 1e3:.1f} M ray-samples/s")

# ---- C4-style stage-3 chunk, forward only: background (128 proposal + 64 NeRF samples, 256 wide) + human branch (128 samples)
# on the same 8192 rays + depth-merge composite (render_hosnerf_chunk)
from hosnerf_b200 import render_hosnerf_chunk
n4 = 8192
hb4 = synth.make_human_batch(n4)
Mw = synth.random_rigid()
ro, rd = hb4["rays"][0], hb4["rays"][1]
ro_w = (Mw[:3, :3] @ ro.T).T + Mw[:3, 3]
rd_w = (Mw[:3, :3] @ rd.T).T
bb4 = {"rays_o": ro_w, "rays_d": rd_w, "viewdirs": rd_w / rd_w.norm(dim=-1, keepdim=True), "radii": torch.full((n4, 1), 1e-3),
       "times": torch.tensor(0.0)}
bb4 = {k: v.to(dev) for k, v in bb4.items()}
hb4 = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in hb4.items()}
bkg4 = MipNeRF360("/nonexistent", num_levels=2, num_prop_samples=128, num_nerf_samples=64, nerf_netwidth=256, opaque_background=True,
                  stage3=True, precision="fp16")
synth.fill_params_(bkg4, 0)
bkg4 = bkg4.to(dev)
hn4 = Network(default_cfg(), stage2=False, precision="fp16")
synth.fill_params_(hn4, 0)
synth.boost_human_density_(hn4)
hn4 = hn4.to(dev)
ms = timeit(lambda: render_hosnerf_chunk(bkg4, hn4, bb4, hb4, Mw), reps=5)
print(f"C4-style stage-3 chunk, forward only ({n4} rays x (128 prop + 64 NeRF + 128 human) samples, fp16): {ms:.3f} ms/chunk = "
      f"{n4 / ms / 1e3:.2f} M rays/s")
