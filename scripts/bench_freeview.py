"""C5 (BASELINE.json configs[4]): free-view inference of one H x W frame - rays generated on the device from (K, R, T),
background branch (3 levels 64/64/64, NeRFMLP 1024 wide) on every ray, human-object branch (128 samples) on the rays that
hit the canonical bounding box, stage-3 depth-merge composite; forward only, fp16 mode.  Prints seconds per frame and
frames/s for ONE GPU (frames or ray shards are independent across GPUs).  A side measurement, not the bench.py workload.

    python scripts/bench_freeview.py [H W [rays_per_chunk]]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from hosnerf_b200 import MipNeRF360, Network, camera, default_cfg, ops, synth

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1080
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
CHUNK = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
dev = "cuda:0"

bkg = MipNeRF360("/nonexistent", num_prop_samples=64, num_nerf_samples=64, opaque_background=True, stage3=True, precision="fp16")
synth.fill_params_(bkg, 0)
bkg = bkg.to(dev)
human = Network(default_cfg(), stage2=False, precision="fp16")
synth.fill_params_(human, 0)
synth.boost_human_density_(human)
human = human.to(dev)
hb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.make_human_batch(8).items()}
M = synth.random_rigid()                                  # newsmpl -> scale-world similarity
Minv = torch.linalg.inv(M.double())

# camera in the new-SMPL frame looking at the canonical box from z = -3 (the C3 ray family), pinhole f = 1500 px at 1080p
f = 1500.0 * H / 1080.0
K = np.array([[f, 0.0, W / 2.0], [0.0, f, H / 2.0], [0.0, 0.0, 1.0]])
R = np.eye(3)
T = np.array([0.0, 0.0, 3.0])
sk = synth.make_skeleton(0)
bmin = sk["cnl_bbox_min_xyz"].double().numpy()
bmax = bmin + 2.0 / sk["cnl_bbox_scale_xyz"].double().numpy()


def render_frame():
    o_h, d_h = camera.get_rays_from_KRT(H, W, K, R, T)                       # human-branch rays, new-SMPL frame
    o_h, d_h = o_h.view(-1, 3), d_h.view(-1, 3)
    Mr, Mt = M[:3, :3].to(dev), M[:3, 3].to(dev)
    o_w = o_h @ Mr.T + Mt                                                     # the same rays in the scale-world frame
    d_w = d_h @ Mr.T
    n = o_w.shape[0]
    viewdirs = d_w / d_w.norm(dim=-1, keepdim=True)
    radii = torch.full((n, 1), float(np.linalg.norm(Mr.cpu().numpy()[:, 0]) / f * 2 / np.sqrt(12)), device=dev)
    rgb = torch.empty(n, 3, device=dev)
    n_hit = 0
    for c0 in range(0, n, CHUNK):
        c1 = min(n, c0 + CHUNK)
        m = c1 - c0
        bb = {"rays_o": o_w[c0:c1].contiguous(), "rays_d": d_w[c0:c1].contiguous(), "viewdirs": viewdirs[c0:c1].contiguous(),
              "radii": radii[c0:c1], "times": hb["time"]}
        _, hist = bkg(bb, 1.0, False, False, 0.1, 1e6)
        h = hist[-1]
        oc, dc = o_h[c0:c1].contiguous(), d_h[c0:c1].clone()
        near, far, hit = camera.rays_intersect_3d_bbox(np.stack([bmin, bmax]), oc, dc)
        S_h = human.cfg.N_samples
        h_rgb = torch.zeros(m, S_h, 3, device=dev)
        h_den = torch.zeros(m, S_h, device=dev)
        h_msk = torch.zeros(m, S_h, device=dev)
        h_pts = torch.zeros(m, S_h, 3, device=dev)
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
    return rgb, n_hit


with torch.no_grad():
    Hs, Ws = H, W
    H, W = max(2, CHUNK // W), W                          # warm-up: one chunk-sized strip (weight packing, lazy inits)
    render_frame()
    H, W = Hs, Ws
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rgb, n_hit = render_frame()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
n = H * W
print(f"free view {H}x{W}: {n} rays ({n_hit} hit the human box), bkg 64+64+64 samples (NeRFMLP 1024 wide) + human 128: "
      f"{dt:.3f} s/frame = {1 / dt:.3f} frames/s on 1 GPU, {n / dt / 1e6:.2f} M rays/s; rgb finite: {bool(torch.isfinite(rgb).all())}, "
      f"mean {rgb.mean().item():.4f}")
