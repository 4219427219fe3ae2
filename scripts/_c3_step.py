import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hosnerf_b200 import Network, default_cfg, synth
dev = "cuda:0"
hb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.make_human_batch(6144).items()}
hn = Network(default_cfg(), stage2=False, precision="fp16")
synth.fill_params_(hn, 0); synth.boost_human_density_(hn); hn = hn.to(dev)
with torch.no_grad():
    for _ in range(3):
        hn(**hb, cycle_outputs=False)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(5):
        hn(**hb, cycle_outputs=False)
    torch.cuda.synchronize()
    print("ms/step wall", (time.perf_counter() - t0) / 5 * 1e3)
