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
            hn(**hb, cycle_outputs=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            hn(**hb, cycle_outputs=False)
        e1.record(); torch.cuda.synchronize()
    print(tag, e0.elapsed_time(e1) / 20, "ms/step")
run("pair kernels, fused fourier")
H.FUSE_FOURIER = False
run("pair kernels, materialised PE")
nr = hn._cache["nr"]
lib.hos_mlp_set_variant(nr._h, 1); nr._opt_key = (1, None)
import hosnerf_b200.ops as ops
run("NR on single-CTA kernel, materialised PE")
