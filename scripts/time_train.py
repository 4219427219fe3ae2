"""Kernel table of one stage-1 training step on the C2 shape (torch profiler) + CPU-side wall time."""
import sys, time, torch
sys.path.insert(0, '.')
from hosnerf_b200 import LitMipNeRF360, synth
from hosnerf_b200.dist import FlatGrads
from bench import MODEL_KW, N_RAYS
dev = torch.device("cuda", 0)
lit = LitMipNeRF360("/nonexistent", **MODEL_KW)
synth.fill_params_(lit.model, 0)
lit = lit.to(dev)
lit._train_frac = 0.5
batch = {k: v.to(dev) for k, v in synth.make_bkg_batch(N_RAYS, seed=100).items()}
batch["target"] = torch.rand(N_RAYS, 3, device=dev)
sink = FlatGrads(lit.model, bucket_of=lambda name: int(name.split(".")[1]))
lit.model._grad_sink = sink
opt = torch.optim.Adam(lit.parameters(), lr=1e-4, fused=True)
def step():
    sink.zero_()
    loss = lit.training_objective(batch, randomized=True)["loss"]
    loss.backward()
    sink.finish()
    opt.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3 * (t1 - t0) / 5:.2f} ms/step, total {1e3 * (t2 - t0) / 5:.2f} ms/step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as p:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=50))
