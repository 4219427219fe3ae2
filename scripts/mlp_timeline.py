"""Per-unit timeline of CTA 0 of the tcgen05 MLP kernels (debug hook hos_mlp_debug_timeline).
usage: python scripts/mlp_timeline.py [variant: 1 single-CTA | 2 cluster pair] [prop|nerf]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hosnerf_b200 import MipNeRF360, synth, _lib, ops

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 2
which = sys.argv[2] if len(sys.argv) > 2 else "nerf"
dev = "cuda:0"
ops.set_mlp_variant(variant)
net = MipNeRF360("/nonexistent", num_levels=2, num_prop_samples=128, num_nerf_samples=128, nerf_netwidth=256,
                 opaque_background=True, precision="fp16")
synth.fill_params_(net, 0)
net = net.to(dev)
b = {k: v.to(dev) for k, v in synth.make_bkg_batch(4096, seed=1).items()}
lib = _lib.load()
with torch.no_grad():
    for _ in range(3):
        net(b, 1.0, False, False, 0.1, 1e6)
    buf = torch.zeros(1024, dtype=torch.int64, device=dev)
    if which == "nerf":
        ops.MLP_TIMELINE = buf                           # every handle stamps its launches; the last one (NeRF MLP) remains
        net(b, 1.0, False, False, 0.1, 1e6)
    else:
        mlp = net.mlps[0]
        _, hist = net(b, 1.0, False, False, 0.1, 1e6)
        ops.MLP_TIMELINE = buf
        sd = hist[0]["sdist"]
        td = (1.0 / (sd / 1e6 + (1.0 - sd) / 0.1)).contiguous()
        mlp.eval_samples(td, b["rays_o"], b["rays_d"], b["radii"].reshape(-1).contiguous(), b["viewdirs"], 0.0, "fp16")
    torch.cuda.synchronize()
    ops.MLP_TIMELINE = None
if variant == 1:
    t = buf.cpu()[:256].view(64, 4)
    t0 = int(t[0, 0])
    print("layer-iter: mma_start  mma_issued  epi_start  epi_end | mma_issue_span  mma->epi_start  epi_span  epi_end->next_mma")
    for i in range(24):
        a, bb, c, d = [int(x) - t0 for x in t[i]]
        nxt = int(t[i + 1, 0]) - t0
        print(f"{i:3d}: {a:8d} {bb:8d} {c:8d} {d:8d} | {bb - a:6d} {c - bb:6d} {d - c:6d} {nxt - d:6d}")
else:
    t = buf.cpu()[:768].view(64, 12)
    t0 = int(t[0, 0])
    nl = 4 if which == "prop" else len(next(iter(net.mlps[-1]._cache["f16"].values()))[1].layers)
    print("unit(tile,layer): mma_enter issued | epi_enter acc_ready epi_done | feat_start feat_end | "
          "issue_span wait | epi_wait epi_span | feat_span | mma_enter->next")
    for i in range(63):
        r = [int(x) for x in t[i]]
        if r[0] == 0:
            break
        rel = [x - t0 if x else 0 for x in r[:8]]
        nxt = int(t[i + 1, 0])
        print(f"{i:3d} (t{i // nl} l{i % nl}): {rel[0]:8d} {rel[1]:8d} | {rel[3]:8d} {rel[4]:8d} {rel[5]:8d} | "
              f"{rel[6]:8d} {rel[7]:8d} | {r[1] - r[0]:6d} {r[8]:6d} | {r[4] - r[3]:6d} {r[5] - r[4]:6d} | "
              f"{(r[7] - r[6]) if r[6] else 0:6d} | {(nxt - r[0]) if nxt else 0:6d}")
if variant == 2:
    for base, name in ((768, "unit 11"), (800, "unit 12")):
        c = [int(x) for x in buf.cpu()[base:base + 12]]
        if c[0]:
            print(name, "epilogue warp 4, per 64-column chunk: (ld ready -> processed, -> signalled, -> next ld ready)",
                  [(c[i * 3 + 1] - c[i * 3], c[i * 3 + 2] - c[i * 3 + 1], (c[i * 3 + 3] - c[i * 3 + 2]) if i < 3 else 0) for i in range(4)])

if variant == 2:
    tt = buf.cpu()
    for uu in (11, 12):
        base = 840 + (uu - 11) * 16
        c = [int(x) for x in tt[base:base + 15]]
        ref = int(tt[uu * 12 + 0])
        if c[0]:
            print(f"unit {uu} MMA thread groups (rel. to mma_enter): [wait_start, wait_end, issued]",
                  [(c[i * 3] - ref, c[i * 3 + 1] - ref, c[i * 3 + 2] - ref) for i in range(5) if c[i * 3]],
                  "| epilogue acc_ready/epi_done of previous unit:", int(tt[(uu - 1) * 12 + 4]) - ref, int(tt[(uu - 1) * 12 + 5]) - ref)

if variant == 2:
    for base, name in ((900, "unit 9 (l0)"), (920, "unit 14 (skip)")):
        c = [int(x) for x in buf.cpu()[base:base + 17]]
        if c[0]:
            print(name, "feature warp 12: per chunk (slot acquired -> chunk done = compute), then (chunk done -> next slot acquired = blocked):",
                  [(c[2 * i + 1] - c[2 * i], (c[2 * i + 2] - c[2 * i + 1]) if i < 7 else 0) for i in range(8)])
