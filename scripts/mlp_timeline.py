"""Per-layer timeline of CTA 0 of the tcgen05 MLP kernel (debug hook hos_mlp_debug_timeline)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hosnerf_b200 import MipNeRF360, synth, _lib

dev = "cuda:0"
net = MipNeRF360("/nonexistent", num_levels=2, num_prop_samples=128, num_nerf_samples=128, nerf_netwidth=256,
                 opaque_background=True, precision="fp16")
synth.fill_params_(net, 0)
net = net.to(dev)
b = {k: v.to(dev) for k, v in synth.make_bkg_batch(4096, seed=1).items()}
with torch.no_grad():
    for _ in range(3):
        net(b, 1.0, False, False, 0.1, 1e6)
    buf = torch.zeros(256, dtype=torch.int64, device=dev)
    lib = _lib.load()
    lib.hos_mlp_debug_timeline(buf.data_ptr())
    # run only the NeRF level's MLP last so its stamps remain: do a full forward, stamps of the last launch (nerf) stay
    net(b, 1.0, False, False, 0.1, 1e6)
    torch.cuda.synchronize()
    lib.hos_mlp_debug_timeline(None)
t = buf.cpu().view(64, 4)
t0 = int(t[0, 0])
print("layer-iter: mma_start  mma_issued  epi_start  epi_end   (cycles, relative) | mma_issue_span  mma->epi_start  epi_span  epi_end->next_mma")
for i in range(24):
    a, bb, c, d = [int(x) - t0 for x in t[i]]
    nxt = int(t[i + 1, 0]) - t0
    print(f"{i:3d}: {a:8d} {bb:8d} {c:8d} {d:8d} | {bb - a:6d} {c - bb:6d} {d - c:6d} {nxt - d:6d}")
