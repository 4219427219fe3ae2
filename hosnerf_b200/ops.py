"""Thin torch <-> C-ABI glue: every function here checks its tensors (CUDA, fp32,
contiguous), allocates the outputs with torch, and calls one entry point of
libhosnerf_b200.so on the current CUDA stream.  No arithmetic happens in Python and
there is no CPU path: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_F32 = torch.float32
PROFILE = None        # bench.py sets this to a list: (n_layers, rows, start_event, stop_event) per MLP launch


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    """cudaStream_t of torch's current stream.  The public accessor costs ~17 us per call (a training step makes ~200);
    the raw accessor torch itself uses for its extensions is two orders of magnitude cheaper."""
    if _raw_stream is not None and _cur_device is not None:
        return _raw_stream(_cur_device())
    return torch.cuda.current_stream().cuda_stream


def _chk(t, name, dtype=_F32):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"hosnerf_b200: `{name}` must be a CUDA tensor (this package has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"hosnerf_b200: `{name}` must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"hosnerf_b200: `{name}` must be contiguous")
    return t


def _p(t):
    return None if t is None else t.data_ptr()


def _host3(v, n=3):
    vals = [float(x) for x in (v.detach().cpu().reshape(-1).tolist() if isinstance(v, torch.Tensor) else v)]
    assert len(vals) == n
    return (C.c_float * n)(*vals)


# ----------------------------------------------------------------------------- sampler
def max_dilate(t, w, dilation, lo, hi):
    _chk(t, "t"), _chk(w, "w")
    n, s = w.shape
    t_out = torch.empty(n, 3 * s + 1, device=t.device, dtype=_F32)
    w_out = torch.empty(n, 3 * s, device=t.device, dtype=_F32)
    _lib.call_unless_empty(n, "hos_max_dilate", _p(t), _p(w), n, s, dilation, lo, hi, _p(t_out), _p(w_out), _stream())
    return t_out, w_out


def sample_intervals(t, logits, u_base, jitter, max_jitter, lo, hi, want_aux=False):
    _chk(t, "t"), _chk(logits, "logits"), _chk(u_base, "u_base"), _chk(jitter, "jitter")
    n, m = logits.shape
    s = u_base.numel()
    out = torch.empty(n, s + 1, device=t.device, dtype=_F32)
    centers = torch.empty(n, s, device=t.device, dtype=_F32) if want_aux else None
    idx = torch.empty(n, s, device=t.device, dtype=torch.int32) if want_aux else None
    jc = 0 if jitter is None else jitter.shape[-1]
    _lib.call_unless_empty(n, "hos_sample_intervals", _p(t), _p(logits), _p(u_base), _p(jitter), jc, max_jitter, n, m, s,
              lo, hi, _p(out), _p(centers), _p(idx), _stream())
    return (out, centers, idx) if want_aux else out


def invert_cdf(t, cw, u_base, jitter, max_jitter, lo, hi):
    """helper.invert_cdf on a given CDF cw [N,M+1] -> (t_out [N,S+1], centers [N,S], idx int32 [N,S])."""
    _chk(t, "t"), _chk(cw, "cw"), _chk(u_base, "u_base"), _chk(jitter, "jitter")
    n, m1 = cw.shape
    s = u_base.numel()
    out = torch.empty(n, s + 1, device=t.device, dtype=_F32)
    centers = torch.empty(n, s, device=t.device, dtype=_F32)
    idx = torch.empty(n, s, device=t.device, dtype=torch.int32)
    jc = 0 if jitter is None else jitter.shape[-1]
    _lib.call_unless_empty(n, "hos_invert_cdf", _p(t), _p(cw), _p(u_base), _p(jitter), jc, max_jitter, n, m1 - 1, s, lo, hi,
                           _p(out), _p(centers), _p(idx), _stream())
    return out, centers, idx


def resample_level(sdist, weights, dilate, dilation, anneal, padding, u_base, jitter, max_jitter,
                   lo, hi, s_near, s_far):
    _chk(sdist, "sdist"), _chk(weights, "weights"), _chk(u_base, "u_base"), _chk(jitter, "jitter")
    n, m = weights.shape
    s = u_base.numel()
    sd = torch.empty(n, s + 1, device=sdist.device, dtype=_F32)
    td = torch.empty(n, s + 1, device=sdist.device, dtype=_F32)
    jc = 0 if jitter is None else jitter.shape[-1]
    _lib.call_unless_empty(n, "hos_resample_level", _p(sdist), _p(weights), n, m, int(dilate), dilation, anneal, padding,
              _p(u_base), _p(jitter), jc, max_jitter, s, lo, hi, s_near, s_far, _p(sd), _p(td), _stream())
    return sd, td


def human_samples(rays_o, rays_d, near, far, t_lin, rand=None):
    for t, nm in ((rays_o, "rays_o"), (rays_d, "rays_d"), (near, "near"), (far, "far"), (t_lin, "t_lin"), (rand, "rand")):
        _chk(t, nm)
    n, s = rays_o.shape[0], t_lin.numel()
    z = torch.empty(n, s, device=rays_o.device, dtype=_F32)
    pts = torch.empty(n, s, 3, device=rays_o.device, dtype=_F32)
    _lib.call_unless_empty(n, "hos_human_samples", _p(rays_o), _p(rays_d), _p(near), _p(far), _p(t_lin), _p(rand), n, s,
              _p(z), _p(pts), _stream())
    return z, pts


# ----------------------------------------------------------------------------- encodings
def tiled_bytes(rows: int, k: int) -> int:
    return ((rows + 127) // 128) * ((k + 63) // 64) * 16384


def ipe_features(tdist, rays_o, rays_d, radii, basis, min_deg=0, max_deg=12, out="fp32", want_aux=False):
    for t, nm in ((tdist, "tdist"), (rays_o, "rays_o"), (rays_d, "rays_d"), (radii, "radii"), (basis, "basis")):
        _chk(t, nm)
    n, s1 = tdist.shape
    s = s1 - 1
    b = basis.shape[1]
    width = 2 * (max_deg - min_deg) * b
    dev = tdist.device
    if out == "fp32":
        feat, ld, code = torch.empty(n * s, width, device=dev, dtype=_F32), width, 0
    elif out == "fp16":
        feat, ld, code = torch.empty(n * s, width, device=dev, dtype=torch.float16), width, 1
    elif out == "f16op":      # fp16 row-major operand plane of the layer GEMMs (fast pairwise evaluation)
        feat, ld, code = torch.empty(n * s, width, device=dev, dtype=torch.float16), width, 4
    elif out == "split":      # [2, rows, width] fp16: hi plane, residual plane
        feat, ld, code = torch.empty(2, n * s, width, device=dev, dtype=torch.float16), width, 3
    elif out == "tiled":
        feat, ld, code = torch.empty(tiled_bytes(n * s, width), device=dev, dtype=torch.uint8), 0, 2
    else:
        raise ValueError(out)
    means = torch.empty(n * s, 3, device=dev, dtype=_F32) if want_aux else None
    lvar = torch.empty(n * s, b, device=dev, dtype=_F32) if want_aux else None
    _lib.call_unless_empty(n * s, "hos_ipe_features", _p(tdist), _p(rays_o), _p(rays_d), _p(radii), _p(basis), n, s, b,
              min_deg, max_deg, _p(feat), ld, code, _p(means), _p(lvar), _stream())
    return (feat, means, lvar) if want_aux else feat


def ipe_from_gaussians(means, covs, basis, min_deg=0, max_deg=12):
    """contract + lift + IPE of caller-supplied Gaussians: means [..., 3], covs [..., 3, 3] -> fp32 [rows, 2 * deg * B]."""
    _chk(means, "means"), _chk(covs, "covs"), _chk(basis, "basis")
    rows = means.numel() // 3
    b = basis.shape[1]
    width = 2 * (max_deg - min_deg) * b
    feat = torch.empty(rows, width, device=means.device, dtype=_F32)
    _lib.call_unless_empty(rows, "hos_ipe_from_gaussians", _p(means), _p(covs), _p(basis), rows, b, min_deg, max_deg, _p(feat), width,
                           _stream())
    return feat


def ipe_features_fast(tdist, rays_o, rays_d, radii, basis_host):
    """504 IPE features per sample, tiled fp16, generation column order (for TiledLinear(ipe_inputs=...))."""
    for t, nm in ((tdist, "tdist"), (rays_o, "rays_o"), (rays_d, "rays_d"), (radii, "radii")):
        _chk(t, nm)
    n, s = tdist.shape[0], tdist.shape[1] - 1
    feat = torch.empty(tiled_bytes(n * s, 504), device=tdist.device, dtype=torch.uint8)
    _lib.call_unless_empty(n * s, "hos_ipe_features_fast", _p(tdist), _p(rays_o), _p(rays_d), _p(radii), basis_host, n, s,
                           _p(feat), _stream())
    return feat


def pos_enc(x, min_deg, max_deg, append_identity=True):
    _chk(x, "x")
    n = x.shape[0]
    width = (3 if append_identity else 0) + 6 * (max_deg - min_deg)
    out = torch.empty(n, width, device=x.device, dtype=_F32)
    _lib.call_unless_empty(n, "hos_pos_enc", _p(x), n, min_deg, max_deg, int(append_identity), _p(out), _stream())
    return out


def fourier_embed(x, n_freqs, include_input, hann_w=None, out="fp32"):
    _chk(x, "x"), _chk(hann_w, "hann_w")
    p = x.shape[0]
    width = (3 if include_input else 0) + 6 * n_freqs
    if out == "fp32":
        o, ld, code = torch.empty(p, width, device=x.device, dtype=_F32), width, 0
    else:
        o, ld, code = torch.empty(tiled_bytes(p, width), device=x.device, dtype=torch.uint8), 0, 2
    _lib.call_unless_empty(p, "hos_fourier_embed", _p(x), p, n_freqs, int(include_input), _p(hann_w), _p(o), ld, code, _stream())
    return o


# ----------------------------------------------------------------------------- LBS
def lbs_warp(pts, R, T, vol, bbox_min, bbox_scale):
    _chk(pts, "pts"), _chk(R, "R"), _chk(T, "T"), _chk(vol, "vol")
    p = pts.numel() // 3
    bones = R.shape[0]
    g = vol.shape[-1]
    assert vol.shape[0] >= bones and vol.shape[1] == g and vol.shape[2] == g
    x = torch.empty(p, 3, device=pts.device, dtype=_F32)
    m = torch.empty(p, device=pts.device, dtype=_F32)
    _lib.call_unless_empty(p, "hos_lbs_warp", _p(pts), _p(R), _p(T), _p(vol), _host3(bbox_min), _host3(bbox_scale), p, bones, g,
              _p(x), _p(m), _stream())
    return x, m


def lbs_warp_backward(pts, R, T, vol, bbox_min, bbox_scale, g_x, g_mask=None):
    """-> (g_vol like vol, g_R [bones,3,3], g_T [bones,3]) for upstream g_x [P,3] (and g_mask [P])."""
    for t, nm in ((pts, "pts"), (R, "R"), (T, "T"), (vol, "vol"), (g_x, "g_x"), (g_mask, "g_mask")):
        _chk(t, nm)
    p, bones, g = pts.numel() // 3, R.shape[0], vol.shape[-1]
    g_vol, g_R, g_T = torch.zeros_like(vol), torch.zeros_like(R), torch.zeros_like(T)
    _lib.call_unless_empty(p, "hos_lbs_warp_backward", _p(pts), _p(R), _p(T), _p(vol), _host3(bbox_min), _host3(bbox_scale), p, bones,
                           g, _p(g_x), _p(g_mask), _p(g_vol), _p(g_R), _p(g_T), _stream())
    return g_vol, g_R, g_T


def lbs_forward_backward(cnl_pts, R_fwd, T_fwd, vol, bbox_min, bbox_scale, g_x):
    """-> (g_vol, g_R, g_T, g_pts [P,3]) of the forward warp."""
    for t, nm in ((cnl_pts, "cnl_pts"), (R_fwd, "R_fwd"), (T_fwd, "T_fwd"), (vol, "vol"), (g_x, "g_x")):
        _chk(t, nm)
    p, bones, g = cnl_pts.numel() // 3, R_fwd.shape[0], vol.shape[-1]
    g_vol, g_R, g_T = torch.zeros_like(vol), torch.zeros_like(R_fwd), torch.zeros_like(T_fwd)
    g_pts = torch.zeros(p, 3, device=cnl_pts.device, dtype=_F32)
    _lib.call_unless_empty(p, "hos_lbs_forward_backward", _p(cnl_pts), _p(R_fwd), _p(T_fwd), _p(vol), _host3(bbox_min),
                           _host3(bbox_scale), p, bones, g, _p(g_x), _p(g_vol), _p(g_R), _p(g_T), _p(g_pts), _stream())
    return g_vol, g_R, g_T, g_pts


def lbs_forward(cnl_pts, R_fwd, T_fwd, vol, bbox_min, bbox_scale):
    _chk(cnl_pts, "cnl_pts"), _chk(R_fwd, "R_fwd"), _chk(T_fwd, "T_fwd"), _chk(vol, "vol")
    p = cnl_pts.numel() // 3
    bones = R_fwd.shape[0]
    g = vol.shape[-1]
    x = torch.empty(p, 3, device=cnl_pts.device, dtype=_F32)
    m = torch.empty(p, device=cnl_pts.device, dtype=_F32)
    _lib.call_unless_empty(p, "hos_lbs_forward", _p(cnl_pts), _p(R_fwd), _p(T_fwd), _p(vol), _host3(bbox_min),
                           _host3(bbox_scale), p, bones, g, _p(x), _p(m), _stream())
    return x, m


# ----------------------------------------------------------------------------- fp32 MLP blocks
def linear_f32(x1, w, b, act=0, x2=None, x2_row_div=1, k1=None, k2=None):
    """y = act([x1[:, :k1] | x2[:, :k2]] @ w.T + b)."""
    _chk(x1, "x1"), _chk(w, "w"), _chk(b, "b"), _chk(x2, "x2")
    m = x1.shape[0]
    k1 = x1.shape[1] if k1 is None else k1
    k2 = 0 if x2 is None else (x2.shape[1] if k2 is None else k2)
    n = w.shape[0]
    assert w.shape[1] == k1 + k2, (w.shape, k1, k2)
    y = torch.empty(m, n, device=x1.device, dtype=_F32)
    _lib.call_unless_empty(m, "hos_linear_f32_ex", _p(x1), x1.stride(0), k1, _p(x2), 0 if x2 is None else x2.stride(0), k2,
              x2_row_div, _p(w), _p(b), m, n, act, _p(y), n, _stream())
    return y


def head_f32(x, w, b, post=0, shift=0.0, add=None):
    _chk(x, "x"), _chk(w, "w"), _chk(b, "b"), _chk(add, "add")
    m, k = x.shape
    n = w.shape[0]
    assert w.shape[1] == k
    y = torch.empty(m, n, device=x.device, dtype=_F32)
    _lib.call_unless_empty(m, "hos_head_f32", _p(x), x.stride(0), k, _p(w), _p(b), m, n, post, shift, _p(add), _p(y), n, _stream())
    return y


# ----------------------------------------------------------------------------- tcgen05 MLP
class FusedMLP:
    """Owner of an opaque ``hos_mlp_t`` (repacked fp16 weights for the tcgen05 kernel)."""

    def __init__(self, in_dim, layers, heads):
        lib = _lib.load()
        self.in_dim = in_dim
        self.layers = layers
        self.heads = heads
        la = (_lib.MlpLayer * len(layers))(*[_lib.MlpLayer(**l) for l in layers])
        ha = (_lib.MlpHead * max(len(heads), 1))(*[_lib.MlpHead(**h) for h in heads])
        self._h = lib.hos_mlp_create(in_dim, len(layers), la, len(heads), ha)
        if not self._h:
            raise RuntimeError("hos_mlp_create failed: " + lib.hos_last_error().decode())
        self.kblocks = lib.hos_mlp_in_kblocks(self._h)

    def set_layer(self, i, w, b):
        _chk(w, "w"), _chk(b, "b")
        l = self.layers[i]
        assert tuple(w.shape) == (l["out_dim"], l["in_h"] + l["in_x"]), (i, w.shape, l)
        _lib.call("hos_mlp_set_layer", self._h, i, _p(w), _p(b), _stream())

    def _apply_handle_options(self):
        """Kernel variant / debug timeline are fields of the handle: push the host-side defaults when they changed."""
        tl = MLP_TIMELINE
        key = (MLP_VARIANT, None if tl is None else tl.data_ptr())
        if getattr(self, "_opt_key", (0, None)) != key:
            lib = _lib.load()
            _lib.check(lib.hos_mlp_set_variant(self._h, MLP_VARIANT), "hos_mlp_set_variant")
            _lib.check(lib.hos_mlp_debug_timeline(self._h, None if tl is None else tl.data_ptr()), "hos_mlp_debug_timeline")
            self._opt_key = key

    def set_ipe_input(self, enable=True):
        """Select the fused-IPE weight-column order; call before set_layer()."""
        _lib.call("hos_mlp_set_ipe_input", self._h, int(enable))
        self.ipe = bool(enable)

    def forward_ipe(self, tdist, rays_o, rays_d, radii, basis_host, rowbias=None, rowbias_div=1):
        """Fused prologue: ray intervals in, head outputs out (features never leave the SM)."""
        for t, nm in ((tdist, "tdist"), (rays_o, "rays_o"), (rays_d, "rays_d"), (radii, "radii"), (rowbias, "rowbias")):
            _chk(t, nm)
        n, s = tdist.shape[0], tdist.shape[1] - 1
        rows = n * s
        outs = [None, None]
        for h in self.heads:
            outs[h["out_slot"]] = torch.empty(rows, h["out_dim"], device=tdist.device, dtype=_F32)
        if PROFILE is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self._apply_handle_options()
        _lib.call_unless_empty(rows, "hos_mlp_forward_ipe", self._h, _p(tdist), _p(rays_o), _p(rays_d), _p(radii), basis_host, n, s,
                  _p(rowbias), rowbias_div, _p(outs[0]), _p(outs[1]), _stream())
        if PROFILE is not None:
            e1.record()
            PROFILE.append((len(self.layers), rows, e0, e1))
        return outs

    def set_bias(self, i, b):
        _chk(b, "b")
        assert b.numel() == self.layers[i]["out_dim"]
        _lib.call("hos_mlp_set_bias", self._h, i, _p(b), _stream())

    def set_head(self, i, w, b):
        _chk(w, "w"), _chk(b, "b")
        _lib.call("hos_mlp_set_head", self._h, i, _p(w), _p(b), _stream())

    def forward(self, x_tiled, rows, rowbias=None, rowbias_div=1, add=None):
        _chk(x_tiled, "x_tiled", torch.uint8), _chk(rowbias, "rowbias"), _chk(add, "add")
        assert x_tiled.numel() >= tiled_bytes(rows, self.kblocks * 64), "x_tiled too small for this MLP"
        outs = [None, None]
        for h in self.heads:
            outs[h["out_slot"]] = torch.empty(rows, h["out_dim"], device=x_tiled.device, dtype=_F32)
        if PROFILE is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self._apply_handle_options()
        _lib.call_unless_empty(rows, "hos_mlp_forward", self._h, _p(x_tiled), rows, _p(rowbias), rowbias_div, _p(add),
                  _p(outs[0]), _p(outs[1]), _stream())
        if PROFILE is not None:
            e1.record()
            PROFILE.append((len(self.layers), rows, e0, e1))
        return outs

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().hos_mlp_destroy(self._h)
                self._h = None
        except Exception:
            pass


GEMM_CLUSTER = 0      # A/B switch for new TiledLinear handles: 0 automatic, 2 / 4 force the pair / quad kernel


class TiledLinear:
    """One wide nn.Linear (n_out a multiple of 256) on the tensor cores: tiled fp16 in, tiled fp16 out
    (``hos_gemm_*``).  ``x2`` is the skip-connection input whose columns follow (or, ``x_first``, precede) x1's in W."""

    def __init__(self, n_out, k1, k2=0, x_first=False, ipe_inputs=0):
        lib = _lib.load()
        self.n_out, self.k1, self.k2 = n_out, k1, k2
        self._h = lib.hos_gemm_create(n_out, k1, k2, int(x_first), int(ipe_inputs))
        if not self._h:
            raise RuntimeError("hos_gemm_create failed: " + lib.hos_last_error().decode())
        self.head_dim = 0
        if GEMM_CLUSTER:
            self.set_cluster(GEMM_CLUSTER)

    def set_cluster(self, cluster_size: int):
        """0 automatic, 2 CTA-pair kernel, 4 quad kernel (weight stages multicast between two pairs)."""
        _lib.check(_lib.load().hos_gemm_set_cluster(self._h, int(cluster_size)), "hos_gemm_set_cluster")

    def set_weight(self, w, b):
        _chk(w, "w"), _chk(b, "b")
        assert tuple(w.shape) == (self.n_out, self.k1 + self.k2), (w.shape, self.n_out, self.k1, self.k2)
        _lib.call("hos_gemm_set_weight", self._h, _p(w), _p(b), _stream())

    def set_head(self, w, b):
        _chk(w, "w"), _chk(b, "b")
        assert w.shape[1] == self.n_out and w.shape[0] <= 4
        self.head_dim = w.shape[0]
        _lib.call("hos_gemm_set_head", self._h, self.head_dim, _p(w), _p(b), _stream())

    def forward(self, x1_tiled, rows, x2_tiled=None, relu=True, want_y=True, head_post=None, head_shift=0.0):
        _chk(x1_tiled, "x1_tiled", torch.uint8), _chk(x2_tiled, "x2_tiled", torch.uint8)
        assert x1_tiled.numel() >= tiled_bytes(rows, self.k1)
        assert (self.k2 == 0) == (x2_tiled is None)
        dev = x1_tiled.device
        y = torch.empty(tiled_bytes(rows, self.n_out), device=dev, dtype=torch.uint8) if want_y else None
        head = torch.empty(rows, self.head_dim, device=dev, dtype=_F32) if head_post is not None else None
        if PROFILE is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.call_unless_empty(rows, "hos_gemm_forward", self._h, _p(x1_tiled), _p(x2_tiled), rows, int(relu), _p(y), _p(head),
                               0 if head_post is None else int(head_post), float(head_shift), _stream())
        if PROFILE is not None:
            e1.record()
            PROFILE.append((("gemm", self.n_out, self.k1 + self.k2), rows, e0, e1))
        return y, head

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().hos_gemm_destroy(self._h)
                self._h = None
        except Exception:
            pass


MLP_VARIANT = 0          # host-side default that FusedMLP applies to ITS handle before a launch (A/B measurements, tests)
MLP_TIMELINE = None      # optional int64 CUDA tensor (>= 1024): debug stamps of the next FusedMLP launches


def set_mlp_variant(variant: int):
    """0 = automatic, 1 = single-CTA tcgen05 kernel, 2 = cluster-pair (cta_group::2) kernel, 3 = cluster-pair kernel with one
    tile pair in flight (A/B against the duo schedule of narrow networks).  The choice is a field of each ``hos_mlp_t`` handle
    (``hos_mlp_set_variant(mlp, v)``); this sets the value FusedMLP objects apply."""
    global MLP_VARIANT
    assert variant in (0, 1, 2, 3)
    MLP_VARIANT = int(variant)


def pack_rows_f16(x, k=None):
    _chk(x, "x")
    rows = x.shape[0]
    k = x.shape[1] if k is None else k
    out = torch.empty(tiled_bytes(rows, k), device=x.device, dtype=torch.uint8)
    _lib.call_unless_empty(rows, "hos_pack_rows_f16", _p(x), rows, x.stride(0), k, _p(out), _stream())
    return out


# ----------------------------------------------------------------------------- composite
def composite_mip360(density, tdist, dirs, rgb=None, opaque_background=False, bg=1.0):
    _chk(density, "density"), _chk(tdist, "tdist"), _chk(dirs, "dirs"), _chk(rgb, "rgb")
    n, s = density.shape
    w = torch.empty(n, s, device=density.device, dtype=_F32)
    out = torch.empty(n, 3, device=density.device, dtype=_F32) if rgb is not None else None
    _lib.call_unless_empty(n, "hos_composite_mip360", _p(density), _p(tdist), _p(dirs), _p(rgb), n, s, int(opaque_background),
              float(bg), _p(w), _p(out), _stream())
    return w, out


def composite_mip360_backward(density, tdist, dirs, rgb=None, g_weights=None, g_rgb_out=None, opaque_background=False, bg=1.0):
    """Gradients of composite_mip360 w.r.t. density (and the per-sample rgb) for upstream dL/dweights and / or
    dL/d(composited rgb)."""
    for t, nm in ((density, "density"), (tdist, "tdist"), (dirs, "dirs"), (rgb, "rgb"), (g_weights, "g_weights"), (g_rgb_out, "g_rgb_out")):
        _chk(t, nm)
    n, s = density.shape
    gd = torch.empty(n, s, device=density.device, dtype=_F32)
    gc = torch.empty(n, s, 3, device=density.device, dtype=_F32) if (rgb is not None and g_rgb_out is not None) else None
    _lib.call_unless_empty(n, "hos_composite_mip360_backward", _p(density), _p(tdist), _p(dirs), _p(rgb), _p(g_weights),
                           _p(g_rgb_out), n, s, int(opaque_background), float(bg), _p(gd), _p(gc), _stream())
    return gd, gc


def composite_nerf(raw, mask, z, dirs, bgcolor=None, activate=True):
    _chk(raw, "raw"), _chk(mask, "mask"), _chk(z, "z"), _chk(dirs, "dirs")
    n, s = z.shape
    dev = raw.device
    rgb = torch.empty(n, 3, device=dev, dtype=_F32)
    acc = torch.empty(n, device=dev, dtype=_F32)
    w = torch.empty(n, s, device=dev, dtype=_F32)
    depth = torch.empty(n, device=dev, dtype=_F32)
    bg = None if bgcolor is None else _host3(bgcolor)
    _lib.call_unless_empty(n, "hos_composite_nerf", _p(raw), _p(mask), _p(z), _p(dirs), bg, n, s, int(activate), _p(rgb), _p(acc),
              _p(w), _p(depth), _stream())
    return rgb, acc, w, depth


def composite_s3(bkg_rgb, bkg_density, bkg_tdist, human_rgb, human_density, pts_mask, newsmpl_pts, M,
                 rays_o, rays_d, thre_fg=5e-3, want_human_w=True):
    """S3 model.py:1524-1596.  ``human_w`` [n,Sh]: composite weights of the human samples in depth order (the
    reference's ``weights_onlyfg[human_pts_idx].reshape(...)``), rows of background rays zero."""
    for t, nm in ((bkg_rgb, "bkg_rgb"), (bkg_density, "bkg_density"), (bkg_tdist, "bkg_tdist"),
                  (human_rgb, "human_rgb"), (human_density, "human_density"), (pts_mask, "pts_mask"),
                  (newsmpl_pts, "newsmpl_pts"), (rays_o, "rays_o"), (rays_d, "rays_d")):
        _chk(t, nm)
    n, sb = bkg_density.shape
    sh = human_density.shape[1]
    dev = bkg_rgb.device
    rgb = torch.empty(n, 3, device=dev, dtype=_F32)
    is_fg = torch.empty(n, device=dev, dtype=torch.uint8)
    hw = torch.empty(n, sh, device=dev, dtype=_F32) if want_human_w else None
    flag = torch.empty(1, device=dev, dtype=torch.int32)       # scratch of the batch-wide small-direction test
    _lib.call_unless_empty(n, "hos_composite_s3", _p(bkg_rgb), _p(bkg_density), _p(bkg_tdist), _p(human_rgb), _p(human_density),
              _p(pts_mask), _p(newsmpl_pts), _host3(M, 16), _p(rays_o), _p(rays_d), n, sb, sh, thre_fg,
              _p(rgb), _p(is_fg), _p(hw), _p(flag), _stream())
    return rgb, is_fg.bool(), hw


# ----------------------------------------------------------------------------- stage-1 loss terms (forward values)
def lossfun_distortion(t, w):
    """helper.lossfun_distortion (S1 helper.py:122-128): t [N,S+1], w [N,S] -> [N]."""
    _chk(t, "t"), _chk(w, "w")
    n, s = w.shape
    if tuple(t.shape) != (n, s + 1):
        raise RuntimeError(f"hosnerf_b200: lossfun_distortion expects t [N,S+1] and w [N,S], got {tuple(t.shape)} / {tuple(w.shape)}")
    out = torch.empty(n, device=w.device, dtype=_F32)
    _lib.call_unless_empty(n, "hos_lossfun_distortion", _p(t), _p(w), n, s, _p(out), _stream())
    return out


def lossfun_outer(t, w, t_env, w_env, want_rows=False):
    """helper.lossfun_outer (S1 helper.py:92-120): fine histogram (t, w) against the proposal envelope
    (t_env, w_env) -> loss [N,S] (and its row sums [N] when ``want_rows``)."""
    _chk(t, "t"), _chk(w, "w"), _chk(t_env, "t_env"), _chk(w_env, "w_env")
    n, s = w.shape
    se = w_env.shape[1]
    if tuple(t.shape) != (n, s + 1) or tuple(t_env.shape) != (n, se + 1) or w_env.shape[0] != n:
        raise RuntimeError("hosnerf_b200: lossfun_outer expects t [N,S+1], w [N,S], t_env [N,Se+1], w_env [N,Se]")
    loss = torch.empty(n, s, device=w.device, dtype=_F32)
    rows = torch.empty(n, device=w.device, dtype=_F32) if want_rows else None
    _lib.call_unless_empty(n, "hos_lossfun_outer", _p(t), _p(w), _p(t_env), _p(w_env), n, s, se, _p(loss), _p(rows), _stream())
    return (loss, rows) if want_rows else loss


def lossfun_distortion_backward(t, w, g_scalar=1.0, g_ray=None):
    """d(g * lossfun_distortion)/dw: [N,S]; g = g_scalar (times g_ray[ray] when given)."""
    _chk(t, "t"), _chk(w, "w"), _chk(g_ray, "g_ray")
    n, s = w.shape
    out = torch.empty(n, s, device=w.device, dtype=_F32)
    _lib.call_unless_empty(n, "hos_lossfun_distortion_backward", _p(t), _p(w), _p(g_ray), float(g_scalar), n, s, _p(out), _stream())
    return out


def lossfun_outer_backward(t, w, t_env, w_env, g_scalar=1.0):
    """d(g_scalar * sum(lossfun_outer))/dw_env: [N,S_env] (the fine histogram is a constant)."""
    _chk(t, "t"), _chk(w, "w"), _chk(t_env, "t_env"), _chk(w_env, "w_env")
    n, s = w.shape
    se = w_env.shape[1]
    out = torch.empty(n, se, device=w.device, dtype=_F32)
    _lib.call_unless_empty(n, "hos_lossfun_outer_backward", _p(t), _p(w), _p(t_env), _p(w_env), float(g_scalar), n, s, se,
                           _p(out), _stream())
    return out


def reduce_scaled(x, scale, y=None):
    """scale * sum(x) (or scale * sum((x - y)^2)) as a 0-d CUDA tensor: one CTA, fixed order, double accumulation."""
    _chk(x, "x"), _chk(y, "y")
    if y is not None and y.numel() != x.numel():
        raise RuntimeError("hosnerf_b200: reduce_scaled: x and y differ in size")
    out = torch.zeros((), device=x.device, dtype=_F32)
    _lib.call_unless_empty(x.numel(), "hos_reduce_scaled", _p(x), _p(y), x.numel(), float(scale), _p(out), _stream())
    return out


# ----------------------------------------------------------------------------- row-major fp16 layer GEMMs (csrc/gemm_tc.cu)
_F16 = torch.float16


def _chk16(t, name):
    """fp16 CUDA matrix, unit column stride, 16-byte aligned base and pitch (what a TMA tensor map needs)."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != _F16 or t.dim() != 2:
        raise RuntimeError(f"hosnerf_b200: `{name}` must be a 2-D CUDA fp16 tensor")
    if t.stride(1) != 1 or (t.stride(0) * 2) % 16 != 0 or t.data_ptr() % 16 != 0:
        raise RuntimeError(f"hosnerf_b200: `{name}` needs unit column stride and a 16-byte aligned base / pitch (stride {t.stride()})")
    return t


def split16(x):
    """fp32 -> (hi, lo) fp16 planes with hi + lo = x to ~22 bits."""
    hi = x.to(_F16)
    return hi, (x - hi.float()).to(_F16)


def _pair(t):
    return t if isinstance(t, (tuple, list)) else (t, None)


def gemm_tma(a0, w0, n, a1=None, w1=None, bias=None, relu=False, mask=None, mode=0, out16=True, out_lo=False, out32=False,
             y16=None, rowbias=None, rowbias_div=1, head=None):
    """mode 0: Y = act([A0 | A1] [W0 | W1]^T + bias) with W* [n, k*];  mode 1: Y = (A0 W0) .* (mask > 0) with W0 [k0, n].
    Operands are fp16 matrices or (hi, lo) pairs of them (split precision: all operands must then be pairs).
    ``head`` = (W [hn, n] fp32, b [hn] or None, post, shift): an fp32 head evaluated in the epilogue; its [rows, hn] result is
    appended to the returned tuple.  Returns (y_hi or None, y_lo or None, y_f32 or None[, head_out])."""
    (a0h, a0l), (w0h, w0l) = _pair(a0), _pair(w0)
    (a1h, a1l), (w1h, w1l) = _pair(a1), _pair(w1)
    for t, nm in ((a0h, "a0"), (a0l, "a0_lo"), (w0h, "w0"), (w0l, "w0_lo"), (a1h, "a1"), (a1l, "a1_lo"), (w1h, "w1"),
                  (w1l, "w1_lo"), (mask, "mask")):
        _chk16(t, nm)
    _chk(bias, "bias")
    rows, k0 = a0h.shape
    k1 = 0 if a1h is None else a1h.shape[1]
    dev = a0h.device
    if mode == 0:
        assert w0h.shape == (n, k0) and (w1h is None or w1h.shape == (n, k1)), (w0h.shape, n, k0, k1)
    else:
        assert w0h.shape[0] == k0 and w0h.shape[1] >= n and a1h is None, (w0h.shape, n, k0)
    n_pad = (n + 7) // 8 * 8
    if out16 and y16 is None:
        y16 = torch.empty(rows, n_pad, device=dev, dtype=_F16)
    ylo = torch.empty(rows, n_pad, device=dev, dtype=_F16) if (out16 and out_lo) else None
    y32 = torch.empty(rows, n_pad, device=dev, dtype=_F32) if out32 else None
    d = _lib.GemmTmaDesc()
    d.mode, d.rows = mode, rows
    d.a0_hi, d.a0_lo, d.k0, d.lda0 = _p(a0h), _p(a0l), k0, a0h.stride(0)
    if a1h is not None:
        d.a1_hi, d.a1_lo, d.k1, d.lda1 = _p(a1h), _p(a1l), k1, a1h.stride(0)
        d.w1_hi, d.w1_lo, d.ldw1 = _p(w1h), _p(w1l), w1h.stride(0)
    d.w0_hi, d.w0_lo, d.ldw0 = _p(w0h), _p(w0l), w0h.stride(0)
    d.n, d.bias, d.relu = n_pad, _p(bias), int(relu)
    if bias is not None:
        assert bias.numel() >= n_pad or n_pad == n, "bias shorter than the padded width"
    if mask is not None:
        d.mask, d.ld_mask = _p(mask), mask.stride(0)
    if rowbias is not None:
        _chk(rowbias, "rowbias")
        assert rowbias.shape[-1] == n_pad and rowbias.shape[0] * rowbias_div >= rows
        d.rowbias, d.rowbias_div = _p(rowbias), int(rowbias_div)
    if y16 is not None:
        d.y_hi, d.y_lo, d.ldy = _p(y16), _p(ylo), y16.stride(0)
    if y32 is not None:
        d.y_f32, d.ldy32 = _p(y32), y32.stride(0)
    hout = None
    if head is not None:
        hw, hb, post, shift = head
        _chk(hw, "head_w"), _chk(hb, "head_b")
        assert hw.shape == (hw.shape[0], n_pad) and hw.shape[0] <= 4
        hout = torch.empty(rows, hw.shape[0], device=dev, dtype=_F32)
        d.hn, d.head_w, d.head_b, d.head_post, d.head_shift, d.head_out = hw.shape[0], _p(hw), _p(hb), int(post), float(shift), _p(hout)
    _lib.call_unless_empty(rows, "hos_gemm_tma", C.byref(d), _stream())
    return (y16, ylo, y32) if head is None else (y16, ylo, y32, hout)


def wgrad_tma(p, q, out, transpose_out=False, colsum=None):
    """out[i, j] += sum_rows p[row, i] q[row, j]  (out fp32 [m, nq], or [nq, m] with transpose_out; any column-sliced view).
    ``colsum`` (fp32 [m], optional): += column sums of p in the same pass (bias gradient when p = dL/dZ)."""
    _chk16(p, "p"), _chk16(q, "q")
    assert out.dtype == _F32 and out.is_cuda and out.stride(1) == 1
    rows, m = p.shape
    nq = q.shape[1]
    assert q.shape[0] == rows and tuple(out.shape) == ((nq, m) if transpose_out else (m, nq)), (p.shape, q.shape, out.shape)
    assert colsum is None or (colsum.dtype == _F32 and colsum.is_cuda and colsum.numel() == m and colsum.is_contiguous())
    if rows == 0:
        return out
    # one launch: blockIdx.y of the kernel walks the 256 x 512 output blocks
    _lib.call("hos_wgrad_tma", p.data_ptr(), m, p.stride(0), q.data_ptr(), nq, q.stride(0), rows, out.data_ptr(), out.stride(0),
              int(transpose_out), None if colsum is None else colsum.data_ptr(), _stream())
    return out


def colsum_f16(x, out, g=None):
    """out[h, c] += sum_rows g[row, h] x[row, c]  (g None: out[c] += column sums)."""
    _chk16(x, "x"), _chk(g, "g")
    rows, n = x.shape
    hn = 0 if g is None else g.shape[1]
    assert out.dtype == _F32 and out.is_cuda and (n % 2) == 0
    ld_out = out.stride(0) if out.dim() == 2 else n
    _lib.call_unless_empty(rows, "hos_colsum_f16", _p(x), rows, n, x.stride(0), _p(g), hn, out.data_ptr(), ld_out, _stream())
    return out


def head_dgrad(g, W, n, add=None, mask=None):
    """fp16 Y[rows, n] = (g [rows, hn] @ W [hn, n] + add) .* (mask > 0)."""
    _chk(g, "g"), _chk16(add, "add"), _chk16(mask, "mask")
    assert W.dtype == _F32 and W.is_cuda and W.stride(1) == 1 and g.dim() == 2 and W.shape[0] == g.shape[1]
    rows, hn = g.shape
    if W.stride(0) % 4 != 0 or W.data_ptr() % 16 != 0:
        W = W.contiguous()
    y = torch.empty(rows, n, device=g.device, dtype=_F16)
    _lib.call_unless_empty(rows, "hos_head_dgrad", _p(g), hn, W.data_ptr(), W.stride(0), _p(add), 0 if add is None else add.stride(0),
                           _p(mask), 0 if mask is None else mask.stride(0), rows, n, _p(y), n, _stream())
    return y
