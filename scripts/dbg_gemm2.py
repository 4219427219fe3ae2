import sys, torch
sys.path.insert(0, '.')
from hosnerf_b200 import ops
rows, k0, n, split = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
g = torch.Generator(device="cuda").manual_seed(5)
a0 = torch.randn(rows, k0, device="cuda", generator=g)
w0 = torch.randn(n, k0, device="cuda", generator=g) / 27
ref = (a0.double() @ w0.double().t())
torch.cuda.synchronize()
try:
    if split:
        y = ops.gemm_tma(ops.split16(a0), ops.split16(w0), n, out_lo=True, out32=True)[2]
    else:
        y = ops.gemm_tma(a0.half(), w0.half(), n, out32=True)[2]
    torch.cuda.synchronize()
    print(sys.argv[1:], "rel", float((y.double() - ref).abs().max() / ref.abs().max()))
except Exception as e:
    print(sys.argv[1:], "FAILED", repr(e)[:80])
