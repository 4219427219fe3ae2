"""Event-timed layer GEMMs at the C2 row count (524288 rows, 256 x 256): us per launch, achieved GB/s and TFLOP/s."""
import sys, torch
sys.path.insert(0, '.')
from hosnerf_b200 import ops
rows, k, n = 4096 * 128, 256, 256
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(rows, k, device="cuda", generator=g).half()
a2 = torch.randn(rows, k, device="cuda", generator=g).half()
h = torch.randn(rows, k, device="cuda", generator=g).half()
w = (torch.randn(n, k, device="cuda", generator=g) / 16).half()
b = torch.randn(n, device="cuda", generator=g)
feat = torch.randn(rows, 504, device="cuda", generator=g).half()
w504 = (torch.randn(n, 504, device="cuda", generator=g) / 16).half()
hw = torch.randn(1, n, device="cuda", generator=g)
out = torch.zeros(n, k, device="cuda")
out504 = torch.zeros(n, 504, device="cuda")
cs = torch.zeros(n, device="cuda")
y = torch.empty(rows, n, device="cuda", dtype=torch.float16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
cases = {
    "fwd 256->256": (lambda: ops.gemm_tma(a, w, n, bias=b, relu=True, y16=y), 2 * rows * k * 2, 2 * rows * k * n),
    "fwd nobias": (lambda: ops.gemm_tma(a, w, n, relu=True, y16=y), 2 * rows * k * 2, 2 * rows * k * n),
    "fwd plain": (lambda: ops.gemm_tma(a, w, n, y16=y), 2 * rows * k * 2, 2 * rows * k * n),
    "fwd 504->256": (lambda: ops.gemm_tma(feat, w504, n, bias=b, relu=True, y16=y), rows * (504 + 256) * 2, 2 * rows * 504 * n),
    "dgrad+mask": (lambda: ops.gemm_tma(a, w, n, mode=1, mask=h, y16=y), 3 * rows * k * 2, 2 * rows * k * n),
    "dgrad nomask": (lambda: ops.gemm_tma(a, w, n, mode=1, y16=y), 2 * rows * k * 2, 2 * rows * k * n),
    "fwd no-out": (lambda: ops.gemm_tma(a, w, n, bias=b, relu=True, out16=False, head=(hw, None, 0, 0.0)), rows * k * 2, 2 * rows * k * n),
    "split fwd": (lambda: ops.gemm_tma((a, a2), (w, w), n, bias=b, relu=True, out_lo=True), 4 * rows * k * 2, 6 * rows * k * n),
    "wgrad 256x256": (lambda: ops.wgrad_tma(a, h, out, colsum=cs), 2 * rows * k * 2, 2 * rows * k * n),
    "wgrad 256x504": (lambda: ops.wgrad_tma(a, feat, out504), rows * (256 + 504) * 2, 2 * rows * 504 * n),
}
r2 = 65536
a1k = torch.randn(r2, 1024, device="cuda", generator=g).half()
h1k = torch.randn(r2, 1024, device="cuda", generator=g).half()
w1k = (torch.randn(1024, 1024, device="cuda", generator=g) / 32).half()
b1k = torch.randn(1024, device="cuda", generator=g)
y1k = torch.empty(r2, 1024, device="cuda", dtype=torch.float16)
o1k = torch.zeros(1024, 1024, device="cuda")
cases["fwd 1024 (65k rows)"] = (lambda: ops.gemm_tma(a1k, w1k, 1024, bias=b1k, relu=True, y16=y1k), 2 * r2 * 1024 * 2, 2 * r2 * 1024 * 1024)
cases["dgrad 1024 (65k rows)"] = (lambda: ops.gemm_tma(a1k, w1k, 1024, mode=1, mask=h1k, y16=y1k), 3 * r2 * 1024 * 2, 2 * r2 * 1024 * 1024)
cases["wgrad 1024 (65k rows)"] = (lambda: ops.wgrad_tma(a1k, h1k, o1k), 2 * r2 * 1024 * 2, 2 * r2 * 1024 * 1024)
gw = torch.randn(rows, 1, device="cuda", generator=g)
gw3 = torch.randn(rows, 3, device="cuda", generator=g)
o1 = torch.zeros(1, n, device="cuda"); o3 = torch.zeros(3, n, device="cuda")
cases["colsum head 1x256"] = (lambda: ops.colsum_f16(a, o1, g=gw), rows * k * 2, 2 * rows * k)
cases["colsum head 3x256"] = (lambda: ops.colsum_f16(a, o3, g=gw3), rows * k * 2, 6 * rows * k)
cases["colsum bias 256"] = (lambda: ops.colsum_f16(a, cs), rows * k * 2, rows * k)
for name, (fn, nbytes, flops) in cases.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = sorted(ts)[len(ts) // 2]
    print(f"{name:16s} {us:8.1f} us  {nbytes / us / 1e3:7.0f} GB/s  {flops / us / 1e6:7.0f} TFLOP/s")
