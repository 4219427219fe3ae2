"""Kernel table of C2 inference steps (4096 rays x (128 + 128) samples, fp16 fused path), torch profiler / CUPTI."""
import sys, torch
sys.path.insert(0, '.')
from hosnerf_b200 import MipNeRF360, synth
from bench import MODEL_KW
dev = "cuda:0"
net = MipNeRF360("/nonexistent", **MODEL_KW, precision="fp16")
synth.fill_params_(net, 0); net = net.to(dev)
b = {k: v.to(dev) for k, v in synth.make_bkg_batch(4096, seed=1).items()}
from torch.profiler import profile, ProfilerActivity
with torch.no_grad():
    for _ in range(3): net(b, 1.0, False, False, 0.1, 1e6)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        for _ in range(10): net(b, 1.0, False, False, 0.1, 1e6)
        torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
print("resample launches (us):", [round(e.device_time, 1) for e in p.events() if "resample_level" in e.name][:8])
print("composite launches (us):", [round(e.device_time, 1) for e in p.events() if "composite_mip360" in e.name][:8])
