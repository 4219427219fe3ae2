"""Kernel table of one C5 background chunk (65536 rays, 64 + 64 proposal + 64 NeRF-1024w samples), torch profiler."""
import sys, torch
sys.path.insert(0, '.')
from hosnerf_b200 import MipNeRF360, synth
dev = "cuda:0"
bkg = MipNeRF360("/nonexistent", num_prop_samples=64, num_nerf_samples=64, opaque_background=True, stage3=True, precision="fp16")
synth.fill_params_(bkg, 0); bkg = bkg.to(dev)
bb = {k: v.to(dev) for k, v in synth.make_bkg_batch(65536, s3_times=True).items()}
with torch.no_grad():
    for _ in range(2):
        bkg(bb, 1.0, False, False, 0.1, 1e6)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        bkg(bb, 1.0, False, False, 0.1, 1e6)
    e1.record(); torch.cuda.synchronize()
    print("ms per chunk", e0.elapsed_time(e1) / 3)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        bkg(bb, 1.0, False, False, 0.1, 1e6)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
