"""Reference-default configuration (Backpack.gin: 64 + 64 proposal samples, 32 NeRF samples, NeRFMLP 1024 wide):
ms per 4096-ray step in fp16 mode (fused proposal kernels + wide-layer GEMMs) and in fp32 mode (SIMT parity kernels),
and the TFLOP/s of the wide-layer GEMM launches.  Not the bench.py workload - a side measurement for DESIGN.md."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hosnerf_b200 import MipNeRF360, synth, ops

dev = "cuda:0"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = {k: v.to(dev) for k, v in synth.make_bkg_batch(N, seed=1).items()}
for prec in ("fp16", "fp32"):
    net = MipNeRF360("/nonexistent", opaque_background=True, precision=prec)
    synth.fill_params_(net, 0)
    net = net.to(dev)
    with torch.no_grad():
        for _ in range(3):
            net(b, 1.0, False, False, 0.1, 1e6)
        torch.cuda.synchronize()
        K = 10 if prec == "fp16" else 3
        ops.PROFILE = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            net(b, 1.0, False, False, 0.1, 1e6)
        e1.record()
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
    ms = e0.elapsed_time(e1) / K
    line = f"{prec}: {ms:.3f} ms per {N}-ray step ({N * 160 / ms / 1e3:.1f} M ray-samples/s)"
    g = [(t[0], t[1], t[2].elapsed_time(t[3])) for t in prof if isinstance(t[0], tuple)]
    if g:
        flop = sum(2.0 * rows * k[1] * k[2] for k, rows, _ in g)
        tms = sum(x[2] for x in g)
        line += f"; wide-layer GEMMs: {len(g) // K} launches/step, {tms / K:.3f} ms/step, {flop / tms / 1e9:.0f} TFLOP/s"
    print(line)
