"""C4 training chunk at N rays per GPU: eager vs whole-step CUDA graph (train.GraphedStep) - gradients agree, ms per step."""
import os, sys, time, torch
sys.path.insert(0, '.')
from hosnerf_b200 import MipNeRF360, Network, default_cfg, synth, train_hosnerf_chunk
from hosnerf_b200.dist import FlatGrads
from hosnerf_b200.train import GraphedStep
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
bkg = MipNeRF360("/nonexistent", num_prop_samples=64, num_nerf_samples=64, opaque_background=True, stage3=True)
synth.fill_params_(bkg, 0); bkg = bkg.to(dev)
human = Network(default_cfg()); synth.fill_params_(human, 0); synth.boost_human_density_(human); human = human.to(dev)
human.static_shapes = True
hb = synth.make_human_batch(n); hb["is_train"] = True
Mw = synth.random_rigid()
ro, rd = hb["rays"][0], hb["rays"][1]
ro_w = (Mw[:3, :3] @ ro.T).T + Mw[:3, 3]; rd_w = (Mw[:3, :3] @ rd.T).T
bb = {"rays_o": ro_w, "rays_d": rd_w, "viewdirs": rd_w / rd_w.norm(dim=-1, keepdim=True), "radii": torch.full((n, 1), 1e-3)}
bb = {k: v.to(dev).contiguous() for k, v in bb.items()}
bb["times"] = torch.tensor(0.0)                      # host scalar: the state index is resolved without a device read
hb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in hb.items()}
for k in ("time", "iter_val"):                      # host scalars (frozen into the graph)
    if isinstance(hb.get(k), torch.Tensor):
        hb[k] = float(hb[k].reshape(-1)[0])
hb["rand"] = torch.rand(n, human.cfg.N_samples, device=dev)
Mw = Mw.to(dev)

class Both(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.model, self.human = bkg, human
both = Both()
sink = FlatGrads(both)
opt = torch.optim.Adam(both.parameters(), lr=0.0, fused=True)       # lr 0: the parameters stay put, so eager and replay see the same step

def fwd_bwd():
    sink.zero_()
    out = train_hosnerf_chunk(bkg, human, bb, hb, Mw, randomized=False, dense=True)
    which = os.environ.get("LOSS", "all")
    if which == "bkg":
        loss = out["ray_history"][-1]["rgb"].mean() + out["ray_history"][-1]["density"].mean()
    elif which == "human":
        loss = out["net_output"]["human_rgb"].mean() + out["net_output"]["human_density"].mean()
    elif which == "cycle":
        loss = (out["net_output"]["deform_pts_final"] - out["net_output"]["observe_pts"]).pow(2).mean()
    else:
        loss = out["rgb"].mean()
    loss.backward()
    return loss.detach()

def timed(f, k=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): f()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, 1e3 * (t1 - t0) / k

def eager():
    l = fwd_bwd(); sink.finish(); opt.step(); return l
ms, enq = timed(eager)
print(f"n={n} eager  : {ms:.2f} ms/step (enqueue {enq:.2f})")
l_e = float(eager()); g_e = sink.flat.clone()
import os
if os.environ.get('ANOMALY'): torch.autograd.set_detect_anomaly(True)
gs = GraphedStep(fwd_bwd, warmup=0)
def graphed():
    l = gs(); sink.finish(); opt.step(); return l
l_g = float(graphed()); g_g = sink.flat.clone()
print("loss eager/graph", l_e, l_g, "grad rel diff", float((g_g - g_e).norm() / g_e.norm()), "max abs", float((g_g - g_e).abs().max()),
      "launches/step", gs.launches_per_step)
ms, enq = timed(graphed)
print(f"n={n} graphed: {ms:.2f} ms/step (enqueue {enq:.2f})")
l_g2 = float(graphed()); print("replay loss", l_g2, "grad rel diff", float((sink.flat - g_e).norm() / g_e.norm()))
if os.environ.get("PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        for _ in range(3): graphed()
        torch.cuda.synchronize()
    print(p.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
