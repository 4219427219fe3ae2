"""Per-unit clock64 stamps of cluster 0 for the non-rigid MLP launch of a C3 step (mlp_pair_kernel debug timeline)."""
import sys, torch
sys.path.insert(0, '.')
from hosnerf_b200 import Network, default_cfg, synth, _lib
dev = "cuda:0"
hb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.make_human_batch(6144).items()}
hn = Network(default_cfg(), stage2=True, precision="fp16")
synth.fill_params_(hn, 0); synth.boost_human_density_(hn); hn = hn.to(dev)
lib = _lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "nr"
var = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with torch.no_grad():
    hn(**hb, cycle_outputs=False)
    m = hn._cache[which]
    tl = torch.zeros(2048, dtype=torch.int64, device=dev)
    lib.hos_mlp_set_variant(m._h, var)
    lib.hos_mlp_debug_timeline(m._h, tl.data_ptr()); m._opt_key = (var, tl.data_ptr())
    import hosnerf_b200.ops as ops
    ops.MLP_VARIANT, ops.MLP_TIMELINE = var, tl
    hn(**hb, cycle_outputs=False)
    torch.cuda.synchronize()
t = tl.cpu().view(-1)[:64 * 12].view(64, 12)
t0 = int(t[0, 0])
print("unit: mma_start mma_end(issue) | epi_start epi_tfull epi_end | mma_wait_sum   (cycles, relative)")
for u in range(40):
    r = t[u]
    f = lambda x: int(x) - t0 if int(x) else -1
    print(f"{u:3d}: {f(r[0]):7d} {f(r[1]):7d} | {f(r[3]):7d} {f(r[4]):7d} {f(r[5]):7d} | {int(r[8]):6d}")
