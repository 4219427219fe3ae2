"""1024-wide layer GEMM (eval path, tiled fp16): pair kernel vs quad kernel, event-timed at the C5 chunk size."""
import sys, torch
sys.path.insert(0, '.')
from hosnerf_b200 import ops
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536 * 64
n, k = 1024, 1024
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(rows, k, device="cuda", generator=g)
xt = ops.pack_rows_f16(x); del x
W = torch.randn(n, k, device="cuda", generator=g) / 32
b = torch.randn(n, device="cuda", generator=g)
ys = {}
for cs in (2, 4, 0):
    lin = ops.TiledLinear(n, k, 0); lin.set_cluster(cs); lin.set_weight(W, b)
    for _ in range(2): y, _h = lin.forward(xt, rows)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): y, _h = lin.forward(xt, rows)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"cluster {cs}: {ms:.3f} ms  {2 * rows * n * k / ms / 1e9:.0f} TFLOP/s")
    ys[cs] = y
print("pair == quad bitwise:", bool(torch.equal(ys[2], ys[4])))
