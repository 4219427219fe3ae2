"""Step time of the C2 background render in each precision mode + per-kernel table (torch profiler)."""
import sys, torch
sys.path.insert(0, '.')
from hosnerf_b200 import LitMipNeRF360, synth
from bench import MODEL_KW, N_RAYS, S_PROP, S_NERF, FLOP_PROP, FLOP_NERF
dev = torch.device("cuda", 0)
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp16x3"]
for prec in modes:
    lit = LitMipNeRF360("/nonexistent", precision=prec, **MODEL_KW)
    synth.fill_params_(lit.model, 0)
    lit = lit.to(dev)
    b = {k: v.to(dev) for k, v in synth.make_bkg_batch(N_RAYS, seed=1).items()}
    for _ in range(3):
        lit.render_rays(b, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for _ in range(K):
        lit.render_rays(b, 0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    fl = N_RAYS * (S_PROP * FLOP_PROP + S_NERF * FLOP_NERF)
    print(f"{prec}: {ms:.3f} ms/step, {fl / ms / 1e9:.1f} TFLOP/s effective, {N_RAYS * (S_PROP + S_NERF) / ms / 1e3:.1f} M ray-samples/s")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        for _ in range(3):
            lit.render_rays(b, 0)
        torch.cuda.synchronize()
    print(p.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
