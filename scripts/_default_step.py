import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hosnerf_b200 import MipNeRF360, synth
dev = "cuda:0"
b = {k: v.to(dev) for k, v in synth.make_bkg_batch(4096, seed=1).items()}
net = MipNeRF360("/nonexistent", opaque_background=True, precision="fp16")
synth.fill_params_(net, 0); net = net.to(dev)
with torch.no_grad():
    for _ in range(6):
        net(b, 1.0, False, False, 0.1, 1e6)
torch.cuda.synchronize()
