"""Row-major fp16 TMA layer GEMMs (csrc/gemm_tc.cu) against torch on the same device: forward (NT), data gradient (NN, W read
as an MN-major operand), weight gradient (TN, both operands MN-major) and the split-precision (hi + lo) forward.
The torch fp32 / fp64 matmul is the reference here because the kernels are floating point (nn.Linear of S1 model.py:212-259
and its autograd backward)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from hosnerf_b200 import ops
    return ops


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("rows,k0,k1,n,relu", [(300, 256, 0, 256, True), (1000, 504, 0, 256, True), (4096 + 77, 256, 504, 256, True),
                                               (513, 256, 0, 128, False), (2048, 128, 40, 128, True), (700, 256, 0, 1024, True)])
def test_gemm_forward_fp16(rows, k0, k1, n, relu):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(rows + n)
    a0 = torch.randn(rows, k0, device="cuda", generator=g).half()
    w0 = (torch.randn(n, k0, device="cuda", generator=g) / (k0 + k1) ** 0.5).half()
    a1 = torch.randn(rows, k1, device="cuda", generator=g).half() if k1 else None
    w1 = (torch.randn(n, k1, device="cuda", generator=g) / (k0 + k1) ** 0.5).half() if k1 else None
    b = torch.randn(n, device="cuda", generator=g)
    y16, _, y32 = ops.gemm_tma(a0, w0, n, a1=a1, w1=w1, bias=b, relu=relu, out32=True)
    ref = a0.float() @ w0.float().t() + b
    if k1:
        ref = ref + a1.float() @ w1.float().t()
    if relu:
        ref = ref.relu()
    assert _rel(y32, ref) < 2e-5
    assert _rel(y16.float(), ref) < 2e-3


def test_gemm_forward_split_precision():
    """x = hi + lo for both operands: three tensor-core passes reproduce the fp32 product to ~1e-6."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    rows, k0, k1, n = 1500, 504, 256, 256
    a0, a1 = torch.randn(rows, k0, device="cuda", generator=g), torch.randn(rows, k1, device="cuda", generator=g).relu()
    w0, w1 = (torch.randn(n, k0, device="cuda", generator=g) / 27), (torch.randn(n, k1, device="cuda", generator=g) / 27)
    b = torch.randn(n, device="cuda", generator=g)
    yh, yl, y32 = ops.gemm_tma(ops.split16(a0), ops.split16(w0), n, a1=ops.split16(a1), w1=ops.split16(w1), bias=b, relu=True,
                               out_lo=True, out32=True)
    ref = (a0.double() @ w0.double().t() + a1.double() @ w1.double().t() + b.double()).relu()
    assert _rel(y32, ref) < 3e-6, _rel(y32, ref)
    assert _rel(yh.float() + yl.float(), ref) < 3e-6
    # the single-plane product of the same operands is three orders of magnitude coarser
    y1 = ops.gemm_tma(a0.half(), w0.half(), n, a1=a1.half(), w1=w1.half(), bias=b, relu=True, out16=False, out32=True)[2]
    assert _rel(y1, ref) > 1e-4


@pytest.mark.parametrize("rows,kred,n", [(1000, 256, 256), (4096 + 5, 128, 256), (640, 256, 64), (900, 256, 504), (300, 1024, 1024)])
def test_gemm_dgrad(rows, kred, n):
    """dA = (dZ W) .* (H > 0): W [kred, n] consumed through MN-major descriptors, no transposed copy."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(rows + n)
    dz = torch.randn(rows, kred, device="cuda", generator=g).half()
    w = (torch.randn(kred, n, device="cuda", generator=g) / kred ** 0.5).half()
    h = torch.randn(rows, (n + 7) // 8 * 8, device="cuda", generator=g).half()
    y, _, y32 = ops.gemm_tma(dz, w, n, mask=h, mode=1, out32=True)
    ref = (dz.float() @ w.float()) * (h[:, :n] > 0)
    assert _rel(y32[:, :n], ref) < 2e-5
    assert _rel(y[:, :n].float(), ref) < 2e-3
    y2 = ops.gemm_tma(dz, w, n, mode=1, out16=False, out32=True)[2]
    assert _rel(y2[:, :n], dz.float() @ w.float()) < 2e-5


@pytest.mark.parametrize("rows,m,nq,tr", [(1000, 256, 256, False), (128 * 300 + 17, 256, 504, False), (5000, 128, 256, True),
                                          (3000, 128, 128, False), (2000, 1024, 768, False), (777, 256, 64, False),
                                          (128 * 40 + 5, 1024, 1528, False), (4100, 640, 1024, True), (300, 520, 520, False)])
def test_wgrad(rows, m, nq, tr):
    """dW += P^T Q over all rows (split over clusters, fp32 atomics)."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(rows + nq)
    p = torch.randn(rows, m, device="cuda", generator=g).half()
    q = torch.randn(rows, nq, device="cuda", generator=g).half()
    base = torch.randn((nq, m) if tr else (m, nq), device="cuda", generator=g)
    out = base.clone()
    cs = torch.ones(m, device="cuda")
    ops.wgrad_tma(p, q, out, transpose_out=tr, colsum=cs)
    assert _rel(cs, 1.0 + p.double().sum(0)) < 1e-5, _rel(cs, 1.0 + p.double().sum(0))
    ref = p.double().t() @ q.double()
    ref = base.double() + (ref.t() if tr else ref)
    assert _rel(out, ref) < 1e-5, _rel(out, ref)


def test_colsum_and_head_dgrad():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(3)
    rows, n = 5003, 256
    x = torch.randn(rows, n, device="cuda", generator=g).half()
    out = torch.zeros(n, device="cuda")
    ops.colsum_f16(x, out)
    assert _rel(out, x.double().sum(0)) < 1e-5
    gw = torch.randn(rows, 3, device="cuda", generator=g)
    out3 = torch.zeros(3, n, device="cuda")
    ops.colsum_f16(x, out3, g=gw)
    assert _rel(out3, gw.double().t() @ x.double()) < 1e-5
    W = torch.randn(3, n, device="cuda", generator=g)
    add = torch.randn(rows, n, device="cuda", generator=g).half()
    y = ops.head_dgrad(gw, W, n, add=add, mask=x)
    ref = (gw @ W + add.float()) * (x > 0)
    assert _rel(y.float(), ref) < 2e-3


def test_empty_and_argument_errors():
    """Empty batches are no-ops; malformed descriptors come back as status codes with a message (no exceptions across the ABI)."""
    import ctypes
    from hosnerf_b200 import _lib
    ops = _ops()
    lib = _lib.load()
    a = torch.zeros(0, 256, device="cuda", dtype=torch.float16)
    w = torch.zeros(256, 256, device="cuda", dtype=torch.float16)
    y, _, _ = ops.gemm_tma(a, w, 256)
    assert y.shape == (0, 256)
    out = torch.zeros(256, 256, device="cuda")
    ops.wgrad_tma(a, a, out)
    assert float(out.abs().max()) == 0.0
    d = _lib.GemmTmaDesc()
    stream = torch.cuda.current_stream().cuda_stream
    assert lib.hos_gemm_tma(ctypes.byref(d), stream) != 0 and b"hos_gemm_tma" in lib.hos_last_error()       # nothing set
    x = torch.randn(256, 100, device="cuda").half()                  # 200-byte pitch: not a multiple of 16
    with pytest.raises(RuntimeError, match="16-byte"):
        ops.gemm_tma(x, w[:, :100].contiguous(), 256)
    g = torch.randn(64, 5, device="cuda")
    with pytest.raises(AssertionError):
        ops.head_dgrad(g, torch.randn(4, 256, device="cuda"), 256)   # head wider than its weight rows
    assert lib.hos_wgrad_tma(None, 256, 256, None, 256, 256, 10, None, 256, 0, None, stream) != 0
    assert lib.hos_lbs_warp_backward(None, None, None, None, None, None, 10, 26, 32, None, None, None, None, None, stream) != 0
