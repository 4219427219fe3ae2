import sys, torch
sys.path.insert(0, '.')
from hosnerf_b200 import Network, default_cfg, synth
dev = "cuda:0"
hb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.make_human_batch(6144).items()}
hn = Network(default_cfg(), stage2=True, precision="fp16")
synth.fill_params_(hn, 0); synth.boost_human_density_(hn); hn = hn.to(dev)
with torch.no_grad():
    for _ in range(3):
        hn(**hb, cycle_outputs=False)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        for _ in range(5):
            hn(**hb, cycle_outputs=False)
        torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=50))
