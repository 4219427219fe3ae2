"""One layer GEMM per mode at the C2 row count, for ncu captures: python scripts/ncu_gemm.py [fwd|dgrad|split|wgrad]"""
import sys, torch
sys.path.insert(0, '.')
from hosnerf_b200 import ops
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
rows, k, n = 4096 * 128, 256, 256
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(rows, k, device="cuda", generator=g).half()
w = (torch.randn(n, k, device="cuda", generator=g) / 16).half()
b = torch.randn(n, device="cuda", generator=g)
if mode in ("fwd1k", "dgrad1k", "wgrad1k", "colsum"):
    r2 = 65536
    a1k = torch.randn(r2, 1024, device="cuda", generator=g).half()
    w1k = (torch.randn(1024, 1024, device="cuda", generator=g) / 32).half()
    b1k = torch.randn(1024, device="cuda", generator=g)
    o1k = torch.zeros(1024, 1024, device="cuda")
    gw = torch.randn(rows, 1, device="cuda", generator=g)
    o1 = torch.zeros(1, n, device="cuda")
for _ in range(2):
    if mode == "fwd1k":
        ops.gemm_tma(a1k, w1k, 1024, bias=b1k, relu=True)
    elif mode == "dgrad1k":
        ops.gemm_tma(a1k, w1k, 1024, mode=1, mask=a1k)
    elif mode == "wgrad1k":
        ops.wgrad_tma(a1k, a1k, o1k)
    elif mode == "colsum":
        ops.colsum_f16(a, o1, g=gw)
    elif mode == "fwd":
        ops.gemm_tma(a, w, n, bias=b, relu=True)
    elif mode == "dgrad":
        ops.gemm_tma(a, w, n, mode=1, mask=a)
    elif mode == "split":
        ops.gemm_tma((a, a), (w, w), n, bias=b, relu=True, out_lo=True)
    elif mode == "wgrad":
        out = torch.zeros(n, k, device="cuda")
        cs = torch.zeros(n, device="cuda")
        ops.wgrad_tma(a, a, out, colsum=cs)
torch.cuda.synchronize()
