import sys, torch
sys.path.insert(0, '.')
from hosnerf_b200 import Network, default_cfg, synth, _lib
import hosnerf_b200.human as H
dev = "cuda:0"
hb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.make_human_batch(6144).items()}
hn = Network(default_cfg(), stage2=True, precision="fp16")
synth.fill_params_(hn, 0); synth.boost_human_density_(hn); hn = hn.to(dev)
lib = _lib.load()
def run(tag):
    with torch.no_grad():
        for _ in range(3):
            out = hn(**hb, cycle_outputs=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            out = hn(**hb, cycle_outputs=False)
        e1.record(); torch.cuda.synchronize()
    print(tag, e0.elapsed_time(e1) / 20, "ms/step")
    return out["rgb"].clone()
a = run("default (duo where possible)")
print({k: type(v).__name__ for k, v in hn._cache.items()})
for key in ("nr", "cnl"):
    m = hn._cache.get(key)
    if m is not None and hasattr(m, "_h"):
        lib.hos_mlp_set_variant(m._h, 3); m._opt_key = (3, None)
b = run("variant 3 (one tile pair in flight)")
print("max |rgb diff|", float((a - b).abs().max()))
from torch.profiler import profile, ProfilerActivity
for var in (0, 3):
    for key in ("nr", "cnl16"):
        m = hn._cache.get(key)
        lib.hos_mlp_set_variant(m._h, var); m._opt_key = (var, None)
    with torch.no_grad():
        hn(**hb, cycle_outputs=False)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as p:
            hn(**hb, cycle_outputs=False)
            torch.cuda.synchronize()
    print("variant", var, [(e.name[:28], round(e.cuda_time, 1)) for e in p.events() if "hos::" in e.name])
