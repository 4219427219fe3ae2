"""C4 training chunk on ONE GPU with N rays (default 1024 = the per-rank share at 8 GPUs): enqueue vs total time and the kernel table."""
import sys, time, torch
sys.path.insert(0, '.')
from hosnerf_b200 import MipNeRF360, Network, default_cfg, synth, train_hosnerf_chunk
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
bkg = MipNeRF360("/nonexistent", num_prop_samples=64, num_nerf_samples=64, opaque_background=True, stage3=True)
synth.fill_params_(bkg, 0); bkg = bkg.to(dev)
human = Network(default_cfg()); synth.fill_params_(human, 0); synth.boost_human_density_(human); human = human.to(dev)
hb = synth.make_human_batch(n); hb["is_train"] = True
Mw = synth.random_rigid()
ro, rd = hb["rays"][0], hb["rays"][1]
ro_w = (Mw[:3, :3] @ ro.T).T + Mw[:3, 3]; rd_w = (Mw[:3, :3] @ rd.T).T
bb = {"rays_o": ro_w, "rays_d": rd_w, "viewdirs": rd_w / rd_w.norm(dim=-1, keepdim=True), "radii": torch.full((n, 1), 1e-3), "times": torch.tensor(0.0)}
bb = {k: v.to(dev).contiguous() for k, v in bb.items()}
hb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in hb.items()}
params = list(bkg.parameters()) + list(human.parameters())
opt = torch.optim.Adam(params, lr=1e-5, fused=True)
def step():
    opt.zero_grad(set_to_none=False)
    out = train_hosnerf_chunk(bkg, human, bb, hb, Mw, randomized=False)
    out["rgb"].mean().backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"n={n}: enqueue {1e3*(t1-t0)/5:.2f} ms/step, total {1e3*(t2-t0)/5:.2f} ms/step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
    for _ in range(2): step()
    torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=46))
print(p.key_averages().table(sort_by="self_cpu_time_total", row_limit=12, max_name_column_width=46))
