"""Per-unit clock64 stamps of cluster 0 for the NeRF MLP launch of a C2 step (mlp_pair_kernel debug timeline)."""
import sys, torch
sys.path.insert(0, '.')
from hosnerf_b200 import MipNeRF360, synth, _lib
import hosnerf_b200.ops as ops
from bench import MODEL_KW
dev = "cuda:0"
net = MipNeRF360("/nonexistent", **MODEL_KW, precision="fp16")
synth.fill_params_(net, 0); net = net.to(dev)
b = {k: v.to(dev) for k, v in synth.make_bkg_batch(4096, seed=1).items()}
tl = torch.zeros(2048, dtype=torch.int64, device=dev)
with torch.no_grad():
    net(b, 1.0, False, False, 0.1, 1e6)
    ops.MLP_TIMELINE = tl
    net(b, 1.0, False, False, 0.1, 1e6)      # both MLP launches stamp; the NeRF MLP (second) overwrites the proposal MLP's stamps
    torch.cuda.synchronize()
t = tl.cpu().view(-1)[:64 * 12].view(64, 12)
t0 = int(t[0, 0])
print("unit: mma_start mma_end(issue) issue-wait | epi_start epi_tfull epi_end | mma_wait_sum | feat_start feat_end")
for u in range(33):
    r = t[u]
    f = lambda x: int(x) - t0 if int(x) else -1
    print(f"{u:3d}: {f(r[0]):7d} {f(r[1]):7d} {int(r[1]) - int(r[0]) - int(r[8]):6d} | {f(r[3]):7d} {f(r[4]):7d} {f(r[5]):7d} | {int(r[8]):6d} | {f(r[6]):7d} {f(r[7]):7d}")
