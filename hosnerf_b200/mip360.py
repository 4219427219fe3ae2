"""Drop-in replacements for the reference's background branch
(S1/src/model/mipnerf360/model.py): ``MipNeRF360MLP`` / ``NeRFMLP`` / ``PropMLP`` /
``MipNeRF360`` / ``LitMipNeRF360.render_rays``.

Same constructor arguments, same ``nn.Parameter`` names and shapes (so Lightning
checkpoints, Adam param groups and ``load_state_dict`` keep working), same
``forward(batch, train_frac, randomized, is_train, near, far)`` call surface and
return structure - but the body launches the kernels of libhosnerf_b200.so instead
of ~150 ATen ops per level.  There is no PyTorch fallback: tensors must live on a
B200.

Two precision modes (``precision=`` keyword / ``set_precision``):
  * ``"fp32"`` - FFMA MLP kernels, every stage fp32: the mode held to the
                 "1e-4 rel" parity gate against the oracle.
  * ``"fp16"`` - IPE features + MLP on the tcgen05 tensor cores (fp16 operands, fp32
                 accumulate, activations never leave the SM); widths <= 256.
  * ``"fp16x3"`` - split precision on the tensor cores: every fp32 operand is carried as an
                 fp16 hi + lo pair and each layer runs as three tcgen05 passes (hi*hi + lo*hi +
                 hi*lo, fp32 accumulate; ``hos_gemm_tma``), features from the accurate fp32
                 encoder.  Meets the same 1e-4 gate as ``"fp32"`` at tensor-core speed.
Sampler and composite are identical (fp32) in both modes.

Training: under autograd ``MipNeRF360.forward(is_train=True)`` is one ``train.RenderFn`` node (fp16 tensor-core layer
GEMMs with saved activations; backward = composite backward + MLP dgrad / wgrad kernels), see train.py.
"""
from __future__ import annotations

import ctypes
import json
import math
import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.init as init

from . import ops
from .geopoly import generate_basis

try:  # the reference decorates these classes with gin; keep the seam when gin is installed
    import gin  # type: ignore
    _configurable = gin.configurable
except Exception:  # pragma: no cover - gin is absent in the build image
    def _configurable(*a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda f: f

EPS = 1.1920929e-07
_DEFAULT_PRECISION = "fp32"
PRECISIONS = ("fp32", "fp16", "fp16x3")
# fp16 mode: generate the IPE features inside the tcgen05 MLP kernel (False: materialise them in HBM
# with hos_ipe_features first - kept for A/B measurements and as the accurate-sin/cos variant)
FUSE_IPE = os.environ.get("HOSNERF_UNFUSED_IPE", "0") != "1"
MERGE_BOTTLENECK = os.environ.get("HOSNERF_KEEP_BOTTLENECK", "0") != "1"


def set_precision(p: str):
    global _DEFAULT_PRECISION
    assert p in PRECISIONS
    _DEFAULT_PRECISION = p


def select_state_index(n_embeds: int, time, transitions_times, eps: float = 1e-5) -> int:
    """Which state embedding is active at ``time`` - S1 model.py:137-206
    (strict ``<`` on the first transition, ``<=`` afterwards)."""
    if n_embeds == 1:
        return 0
    t = float(time)
    if t < float(transitions_times[0]) - eps:
        return 0
    for k in range(1, n_embeds - 1):
        if t <= float(transitions_times[k]) + eps:
            return k
    return n_embeds - 1


def _read_transitions(basedir):
    path = os.path.join(basedir, "transitions_times.json") if basedir else ""
    if path and os.path.exists(path):
        with open(path, "r") as f:
            infos = json.load(f)
        return np.stack([np.array(infos[k]["time"], dtype=np.float32) for k in infos], axis=0)
    return None


@_configurable()
class MipNeRF360MLP(nn.Module):
    """Parameter container + kernel dispatch for one proposal / NeRF MLP
    (S1 model.py:28-264)."""

    def __init__(self, basedir, netdepth: int = 8, netwidth: int = 256, bottleneck_width: int = 256,
                 netdepth_condition: int = 1, netwidth_condition: int = 128, min_deg_point: int = 0,
                 max_deg_point: int = 12, skip_layer: int = 4, skip_layer_dir: int = 4,
                 num_rgb_channels: int = 3, num_density_channels: int = 1, deg_view: int = 4,
                 bottleneck_noise: float = 0.0, density_bias: float = -1.0, density_noise: float = 0.0,
                 rgb_premultiplier: float = 1.0, rgb_bias: float = 0.0, rgb_padding: float = 0.001,
                 basis_shape: str = "icosahedron", basis_subdivision: int = 2, disable_rgb: bool = False):
        for name, value in vars().items():
            if name not in ["self", "__class__"]:
                setattr(self, name, value)
        super().__init__()
        self.register_buffer("pos_basis_t", generate_basis(basis_shape, basis_subdivision))
        self.ipe_size = ((max_deg_point - min_deg_point) * 2) * self.pos_basis_t.shape[-1]
        view_pos_size = (deg_view * 2 + 1) * 3
        self.embedding_size = 64
        pos_size = self.ipe_size + self.embedding_size

        self.transitions_times = _read_transitions(basedir)
        n_states = 1 if self.transitions_times is None else self.transitions_times.shape[0] + 1
        self.bkgd_stateembeds = nn.ParameterList(
            [nn.Parameter(torch.randn(self.embedding_size), requires_grad=True) for _ in range(n_states)])

        def lin(i, o):
            m = nn.Linear(i, o)
            init.kaiming_uniform_(m.weight)
            return m

        layers = [lin(pos_size, netwidth)]
        for idx in range(netdepth - 1):
            layers.append(lin(netwidth + pos_size if (idx % skip_layer == 0 and idx > 0) else netwidth, netwidth))
        self.pts_linear = nn.ModuleList(layers)
        self.density_layer = lin(netwidth, num_density_channels)
        if not disable_rgb:
            self.bottleneck_layer = nn.Linear(netwidth, bottleneck_width)
            views = [lin(bottleneck_width + view_pos_size, netwidth_condition)]
            for idx in range(netdepth_condition - 1):
                views.append(lin(netwidth_condition + view_pos_size if (idx % skip_layer_dir == 0 and idx > 0)
                                 else netwidth_condition, netwidth_condition))
            self.views_linear = nn.ModuleList(views)
            self.rgb_layer = nn.Linear(netwidth_condition, num_rgb_channels)
            init.kaiming_uniform_(self.bottleneck_layer.weight)
            init.kaiming_uniform_(self.rgb_layer.weight)
        self._cache = {}

    # ------------------------------------------------------------------ helpers
    def _skip_inputs(self, i: int) -> bool:
        """True if pts_linear[i] consumes cat([x, inputs]) (S1 model.py:215-216)."""
        return i >= 1 and (i - 1) % self.skip_layer == 0 and (i - 1) > 0

    def _state_index(self, time) -> int:
        return select_state_index(len(self.bkgd_stateembeds), time, self.transitions_times)

    def _versions(self):
        """Cache key of everything derived from the parameters: in-place version AND storage identity
        (``module.to(dev)``, ``.half()`` and ``param.data = ...`` replace the storage without bumping ``_version``)."""
        ps = list(self.parameters()) + [self.pos_basis_t]
        return tuple((p._version, p.data_ptr(), p.dtype) for p in ps)

    def _slot(self, kind: str, state_idx: int, ver):
        """Per-state cache slot (a multi-state sequence alternates between embeddings: one slot per state instead of
        rebuilding - and freeing device buffers a captured CUDA graph may still point into - on every switch)."""
        d = self._cache.setdefault(kind, {})
        ent = d.get(state_idx)
        return d, (ent[1] if ent is not None and ent[0] == ver else None)

    def _check_supported(self):
        if self.netdepth_condition != 1 or self.num_density_channels != 1:
            raise NotImplementedError("hosnerf_b200: netdepth_condition != 1 / density channels != 1 are not built")

    def _folded(self, state_idx: int, _ver=None):
        """fp32 weights with the (constant per call) state embedding folded into the bias of
        every layer that sees the encoded input: W[:, ipe:ipe+64] @ e  (S1 model.py:208-209)."""
        ver = _ver if _ver is not None else self._versions()
        slots, hit = self._slot("f32", state_idx, ver)
        if hit is not None:
            return hit
        self._check_supported()
        e = self.bkgd_stateembeds[state_idx].detach()
        F, nw = self.ipe_size, self.netwidth
        out = {"layers": []}
        for i, m in enumerate(self.pts_linear):
            W, b = m.weight.detach(), m.bias.detach()
            if i == 0:
                out["layers"].append((W[:, :F].contiguous(), (b + W[:, F:] @ e).contiguous(), False))
            elif self._skip_inputs(i):
                Wc = torch.cat([W[:, :nw], W[:, nw:nw + F]], dim=1).contiguous()
                out["layers"].append((Wc, (b + W[:, nw + F:] @ e).contiguous(), True))
            else:
                out["layers"].append((W.contiguous(), b.contiguous(), False))
        out["density"] = (self.density_layer.weight.detach().contiguous(), self.density_layer.bias.detach().contiguous())
        if not self.disable_rgb:
            bw = self.bottleneck_width
            Wv, bv = self.views_linear[0].weight.detach(), self.views_linear[0].bias.detach()
            out["bottleneck"] = (self.bottleneck_layer.weight.detach().contiguous(),
                                 self.bottleneck_layer.bias.detach().contiguous())
            out["views"] = (Wv.contiguous(), bv.contiguous(), Wv[:, :bw].contiguous(), Wv[:, bw:].contiguous())
            Wr = self.rgb_layer.weight.detach() * self.rgb_premultiplier
            br = self.rgb_layer.bias.detach() * self.rgb_premultiplier + self.rgb_bias
            out["rgb"] = (Wr.contiguous(), br.contiguous())
        slots[state_idx] = (ver, out)
        return out

    def _fused(self, state_idx: int, _ver=None):
        """tcgen05 program + uploaded fp16 weights (rebuilt when any parameter changes)."""
        ver = _ver if _ver is not None else self._versions()
        slots, hit = self._slot("f16", state_idx, ver)
        if hit is not None:
            return hit
        if self.netwidth > 256 or self.netwidth % 64 != 0:
            raise NotImplementedError(
                f"hosnerf_b200: the fused tcgen05 MLP kernel supports widths <= 256 (got {self.netwidth}); "
                "wide networks run layer by layer (_wide)")
        f = self._folded(state_idx, ver)
        F, nw = self.ipe_size, self.netwidth
        layers, heads = [], []
        for i in range(self.netdepth):
            layers.append(dict(out_dim=nw, in_h=0 if i == 0 else nw, in_x=F if (i == 0 or f["layers"][i][2]) else 0,
                               x_first=0, relu=1, rowbias=0, head=-1))
        heads.append(dict(out_dim=1, post=1, shift=float(self.density_bias), out_slot=0))
        layers[-1]["head"] = 0
        # The bottleneck layer has no activation (S1 model.py:238-248), so bottleneck -> view layer is one
        # linear map:  Wv[:, :bw] (Wb h + bb) = (Wv[:, :bw] Wb) h + Wv[:, :bw] bb.  Merging them on the
        # host removes a 256x256 layer (7.7 % of the NeRF MLP's FLOPs) from the tensor-core program.
        merge = MERGE_BOTTLENECK and not self.disable_rgb
        if not self.disable_rgb:
            if not merge:
                layers.append(dict(out_dim=self.bottleneck_width, in_h=nw, in_x=0, x_first=0, relu=0, rowbias=0, head=-1))
            layers.append(dict(out_dim=self.netwidth_condition, in_h=nw if merge else self.bottleneck_width, in_x=0,
                               x_first=0, relu=1, rowbias=1, head=1))
            heads.append(dict(out_dim=self.num_rgb_channels, post=2, shift=float(self.rgb_padding), out_slot=1))
        mlp = ops.FusedMLP(F, layers, heads)
        fuse = (FUSE_IPE and self.pos_basis_t.shape[1] == 21 and self.max_deg_point - self.min_deg_point == 12
                and self.min_deg_point == 0)
        if fuse:   # feature generation inside the MLP kernel: weights take the generation column order
            mlp.set_ipe_input(True)
            b = self.pos_basis_t.detach().float().cpu().contiguous().reshape(-1).tolist()
            mlp.basis_host = (ctypes.c_float * len(b))(*b)
        mlp.fused_ipe = fuse
        for i in range(self.netdepth):
            W, b, _ = f["layers"][i]
            mlp.set_layer(i, W, b)
        mlp.set_head(0, *f["density"])
        if not self.disable_rgb:
            Wb, bb = f["bottleneck"]
            Wv1, bv = f["views"][2], f["views"][1]
            if merge:
                mlp.set_layer(self.netdepth, (Wv1 @ Wb).contiguous(), None)
                mlp.view_bias = (bv + Wv1 @ bb).contiguous()
            else:
                mlp.set_layer(self.netdepth, Wb, bb)
                mlp.set_layer(self.netdepth + 1, Wv1, None)
                mlp.view_bias = bv
            mlp.set_head(1, *f["rgb"])
        mlp.folded = f                      # the view-term tensors a captured graph points into live as long as the program
        slots[state_idx] = (ver, mlp)
        return mlp

    def _wide(self, state_idx: int, _ver=None):
        """Wide networks (netwidth a multiple of 256 above 256 - the reference default NeRFMLP is 1024 wide,
        S1 model.py:267-275): every pts_linear layer is one tensor-core GEMM over tiled fp16 activations
        (``ops.TiledLinear``), the density head rides in the last one's epilogue, and the narrow bottleneck + view
        layers run on the fused kernel."""
        ver = _ver if _ver is not None else self._versions()
        slots, hit = self._slot("wide", state_idx, ver)
        if hit is not None:
            return hit
        if self.netwidth % 256 != 0:
            raise NotImplementedError(f"hosnerf_b200: fp16 mode needs netwidth <= 256 or a multiple of 256 (got {self.netwidth}); "
                                      "use precision='fp32' for this network")
        f = self._folded(state_idx, ver)
        F, nw = self.ipe_size, self.netwidth
        fast = (FUSE_IPE and self.pos_basis_t.shape[1] == 21 and self.max_deg_point - self.min_deg_point == 12
                and self.min_deg_point == 0)      # features from the fast generator (kernel column order)
        lins = []
        for i in range(self.netdepth):
            W, b, skip = f["layers"][i]
            lin = ops.TiledLinear(nw, F if i == 0 else nw, F if skip else 0,       # _folded orders skip weights [h | x]
                                  ipe_inputs=(1 if i == 0 else (2 if skip else 0)) if fast else 0)
            lin.set_weight(W, b)
            lins.append((lin, skip))
        lins[-1][0].set_head(*f["density"])
        tail = None
        if not self.disable_rgb:
            bw, cw = self.bottleneck_width, self.netwidth_condition
            if bw % 64 != 0 or bw > 256 or cw % 64 != 0 or cw > 256:
                raise NotImplementedError("hosnerf_b200: bottleneck / condition widths must be multiples of 64 up to 256")
            layers = [dict(out_dim=bw, in_h=0, in_x=nw, x_first=0, relu=0, rowbias=0, head=-1),
                      dict(out_dim=cw, in_h=bw, in_x=0, x_first=0, relu=1, rowbias=1, head=0)]
            heads = [dict(out_dim=self.num_rgb_channels, post=2, shift=float(self.rgb_padding), out_slot=0)]
            tail = ops.FusedMLP(nw, layers, heads)
            tail.set_layer(0, *f["bottleneck"])
            tail.set_layer(1, f["views"][2], None)
            tail.set_head(0, *f["rgb"])
            tail.view_bias = f["views"][1]
        out = {"lins": lins, "tail": tail, "fast": fast}
        if fast:
            bh = self.pos_basis_t.detach().float().cpu().contiguous().reshape(-1).tolist()
            out["basis_host"] = (ctypes.c_float * len(bh))(*bh)
        slots[state_idx] = (ver, out)
        return out

    def _x3(self, state_idx: int, _ver=None):
        """Split-precision operands of every layer: (hi, lo) fp16 pairs of the folded fp32 weights (``precision="fp16x3"``)."""
        ver = _ver if _ver is not None else self._versions()
        slots, hit = self._slot("x3", state_idx, ver)
        if hit is not None:
            return hit
        f = self._folded(state_idx, ver)
        F, nw = self.ipe_size, self.netwidth
        if nw % 8 != 0:
            raise NotImplementedError("hosnerf_b200: fp16x3 mode needs netwidth % 8 == 0")
        sp = lambda W: tuple(t.contiguous() for t in ops.split16(W.contiguous()))
        out = {"layers": []}
        for i, (W, b, skip) in enumerate(f["layers"]):
            if skip:          # _folded orders the skip layer's columns [h | x]
                out["layers"].append((sp(W[:, :nw]), sp(W[:, nw:nw + F]), b))
            else:
                out["layers"].append((sp(W), None, b))
        if not self.disable_rgb:
            # bottleneck -> view layer is one linear map (no activation in between, S1 model.py:238-248): merged in fp64,
            #   Wv[:, :bw] (Wb h + bb) = (Wv[:, :bw] Wb) h + Wv[:, :bw] bb,  which removes one 256-wide layer pass
            Wb, bb = f["bottleneck"]
            Wv1, bv = f["views"][2], f["views"][1]
            out["views"] = sp((Wv1.double() @ Wb.double()).float())
            out["view_bias"] = (bv.double() + Wv1.double() @ bb.double()).float().contiguous()
        slots[state_idx] = (ver, out)
        return out

    # ------------------------------------------------------------------ evaluation
    def eval_samples(self, tdist, rays_o, rays_d, radii, viewdirs, time, precision: str):
        """tdist [N,S+1] -> density [N,S], rgb [N,S,3] (zeros for proposal MLPs).  Covers
        cast_rays + contract + IPE + MLP (S1 model.py:410-424 and 126-264)."""
        n, s = tdist.shape[0], tdist.shape[1] - 1
        st = self._state_index(time)
        ver = self._versions()                  # one pass over the parameters per call (the caches below all key on it)
        f = self._folded(st, ver)
        basis = self.pos_basis_t
        if precision == "fp16x3":
            x3 = self._x3(st, ver)
            rows = n * s
            fs = ops.ipe_features(tdist, rays_o, rays_d, radii, basis, self.min_deg_point, self.max_deg_point, "split")
            feat = (fs[0], fs[1])
            x, dens = feat, None
            nl = len(x3["layers"])
            for i, (Wh, Wx, b) in enumerate(x3["layers"]):
                last = i == nl - 1
                kw = dict(bias=b, relu=True, out_lo=True, out16=not (last and self.disable_rgb))
                if last:                    # density layer from the un-rounded fp32 accumulator, same pass
                    kw["head"] = (f["density"][0], f["density"][1], 1, float(self.density_bias))
                if Wx is not None:
                    kw.update(a1=feat, w1=Wx)
                res = ops.gemm_tma(x, Wh, self.netwidth, **kw)
                x = (res[0], res[1])
                if last:
                    dens = res[3]
            density = dens.view(n, s)
            if self.disable_rgb:
                return density, torch.zeros(n, s, 3, device=tdist.device)
            de = ops.pos_enc(viewdirs, 0, self.deg_view, True)
            rowbias = ops.linear_f32(de, f["views"][3], x3["view_bias"])              # per-ray view term + bias, fp32
            rgb = ops.gemm_tma(x, x3["views"], self.netwidth_condition, relu=True, rowbias=rowbias, rowbias_div=s,
                               out16=False, head=(f["rgb"][0], f["rgb"][1], 2, float(self.rgb_padding)))[3].view(n, s, 3)
            return density, rgb
        if precision == "fp16" and self.netwidth > 256:
            wide = self._wide(st, ver)
            rows = n * s
            if wide["fast"]:
                feat = ops.ipe_features_fast(tdist, rays_o, rays_d, radii, wide["basis_host"])
            else:
                feat = ops.ipe_features(tdist, rays_o, rays_d, radii, basis, self.min_deg_point, self.max_deg_point, "tiled")
            x = feat
            dens = None
            for i, (lin, skip) in enumerate(wide["lins"]):
                last = i == len(wide["lins"]) - 1
                x, head = lin.forward(x, rows, x2_tiled=feat if skip else None, relu=True,
                                      want_y=not (last and self.disable_rgb),
                                      head_post=1 if last else None, head_shift=float(self.density_bias))
                if last:
                    dens = head
            density = dens.view(n, s)
            if self.disable_rgb:
                return density, torch.zeros(n, s, 3, device=tdist.device)
            tail = wide["tail"]
            de = ops.pos_enc(viewdirs, 0, self.deg_view, True)
            rowbias = ops.linear_f32(de, f["views"][3], tail.view_bias)            # per-ray view term + bias
            rgb = tail.forward(x, rows, rowbias=rowbias, rowbias_div=s)[0]
            return density, rgb.view(n, s, 3)
        if precision == "fp16":
            mlp = self._fused(st, ver)
            rowbias = None
            if not self.disable_rgb:
                de = ops.pos_enc(viewdirs, 0, self.deg_view, True)
                rowbias = ops.linear_f32(de, f["views"][3], mlp.view_bias)        # per-ray view term + bias
            if mlp.fused_ipe:
                dens, rgb = mlp.forward_ipe(tdist, rays_o, rays_d, radii, mlp.basis_host, rowbias=rowbias, rowbias_div=s)
            else:
                feat = ops.ipe_features(tdist, rays_o, rays_d, radii, basis, self.min_deg_point, self.max_deg_point,
                                        "tiled")
                dens, rgb = mlp.forward(feat, n * s, rowbias=rowbias, rowbias_div=s)
            density = dens.view(n, s)
            rgb = rgb.view(n, s, 3) if rgb is not None else torch.zeros(n, s, 3, device=tdist.device)
            return density, rgb
        feat = ops.ipe_features(tdist, rays_o, rays_d, radii, basis, self.min_deg_point, self.max_deg_point, "fp32")
        x = feat
        for i, (W, b, skip) in enumerate(f["layers"]):
            x = ops.linear_f32(x, W, b, act=1, x2=feat if skip else None)
        density = ops.head_f32(x, *f["density"], post=1, shift=float(self.density_bias)).view(n, s)
        if self.disable_rgb:
            return density, torch.zeros(n, s, 3, device=tdist.device)
        bott = ops.linear_f32(x, *f["bottleneck"], act=0)
        de = ops.pos_enc(viewdirs, 0, self.deg_view, True)
        v = ops.linear_f32(bott, f["views"][0], f["views"][1], act=1, x2=de, x2_row_div=s)
        rgb = ops.head_f32(v, *f["rgb"], post=2, shift=float(self.rgb_padding)).view(n, s, 3)
        return density, rgb

    def forward(self, gaussians, viewdirs, randomized, is_train, time):
        """The reference's own entry of the MLP (S1 model.py:223-264): ``gaussians = (means [N,S,3], covs [N,S,3,3])`` as
        cast_rays produces them (un-contracted), ``viewdirs`` [N,3] -> ``{"density": [N,S], "rgb": [N,S,3]}``.  Runs the
        un-fused fp32 kernels (encoder from the given Gaussians, FFMA layers); ``MipNeRF360.forward`` does not come through
        here - it evaluates samples straight from ray intervals (``eval_samples``)."""
        means, covs = gaussians
        if not means.is_cuda:
            raise RuntimeError("hosnerf_b200.MipNeRF360MLP: inputs must be CUDA tensors (no CPU fallback)")
        if randomized and (self.density_noise > 0 or self.bottleneck_noise > 0):
            raise NotImplementedError("hosnerf_b200: density_noise / bottleneck_noise > 0 with randomized=True is not built")
        n, s = means.shape[0], means.shape[1]
        st = self._state_index(time)
        f = self._folded(st)
        feat = ops.ipe_from_gaussians(means.contiguous().float(), covs.contiguous().float(), self.pos_basis_t,
                                      self.min_deg_point, self.max_deg_point)
        x = feat
        for W, b, skip in f["layers"]:
            x = ops.linear_f32(x, W, b, act=1, x2=feat if skip else None)
        density = ops.head_f32(x, *f["density"], post=1, shift=float(self.density_bias)).view(n, s)
        if self.disable_rgb:
            return {"density": density, "rgb": torch.zeros_like(means)}
        bott = ops.linear_f32(x, *f["bottleneck"], act=0)
        de = ops.pos_enc(viewdirs.contiguous().float(), 0, self.deg_view, True)
        v = ops.linear_f32(bott, f["views"][0], f["views"][1], act=1, x2=de, x2_row_div=s)
        rgb = ops.head_f32(v, *f["rgb"], post=2, shift=float(self.rgb_padding)).view(n, s, 3)
        return {"density": density, "rgb": rgb}


@_configurable()
class NeRFMLP(MipNeRF360MLP):
    def __init__(self, basedir, netdepth: int = 8, netwidth: int = 1024):
        super().__init__(basedir, netdepth=netdepth, netwidth=netwidth)


@_configurable()
class PropMLP(MipNeRF360MLP):
    def __init__(self, basedir, netdepth: int = 4, netwidth: int = 256):
        super().__init__(basedir, netdepth=netdepth, netwidth=netwidth, disable_rgb=True)


@_configurable()
class MipNeRF360(nn.Module):
    """S1 model.py:291-461 (``stage3=True``: the S3 variant :379-540 - 0-d ``times``,
    ``tdist`` in the history, no renderings)."""

    def __init__(self, basedir, num_prop_samples: int = 64, num_nerf_samples: int = 32, num_levels: int = 3,
                 bg_intensity_range: Tuple[float] = (1.0, 1.0), anneal_slope: int = 10,
                 stop_level_grad: bool = True, use_viewdirs: bool = True, ray_shape: str = "cone",
                 disable_integration: bool = False, single_jitter: bool = True,
                 dilation_multiplier: float = 0.5, dilation_bias: float = 0.0025, num_glo_features: int = 0,
                 num_glo_embeddings: int = 1000, learned_exposure_scaling: bool = False,
                 near_anneal_rate: Optional[float] = None, near_anneal_init: float = 0.95,
                 single_mlp: bool = False, resample_padding: float = 0.0, use_gpu_resampling: bool = False,
                 opaque_background: bool = False,
                 # --- extensions (keyword-only in spirit; the reference binds these through gin) ---
                 nerf_netwidth: Optional[int] = None, prop_netwidth: Optional[int] = None,
                 precision: Optional[str] = None, stage3: bool = False):
        for name, value in vars().items():
            if name not in ["self", "__class__"]:
                setattr(self, name, value)
        super().__init__()
        if ray_shape != "cone" or disable_integration:
            raise NotImplementedError("hosnerf_b200: only ray_shape='cone' with integration is built")
        prop_kw = {} if prop_netwidth is None else {"netwidth": prop_netwidth}
        nerf_kw = {} if nerf_netwidth is None else {"netwidth": nerf_netwidth}
        self.mlps = nn.ModuleList([PropMLP(basedir, **prop_kw) for _ in range(num_levels - 1)]
                                  + [NeRFMLP(basedir, **nerf_kw)])
        self._u_cache = {}
        # True: the stratified-sampling jitter is drawn by the device generator (the reference draws it on the host and copies
        # it over, helper.py:325-328) - no host work per step, and the step can be captured in a CUDA graph (train.GraphedStep)
        self.device_rng = False

    # quantiles of helper.sample (deterministic_center=True): generated on the HOST with the same
    # torch.linspace call as the reference so both sides invert the CDF at bit-identical u.
    def _u_base(self, s: int, randomized: bool, device):
        key = (s, randomized, str(device))
        if key not in self._u_cache:
            if not randomized:
                pad = 1 / (2 * s)
                u = torch.linspace(pad, 1 - pad - EPS, s)
                mj = 0.0
            else:
                u_max = EPS + (1 - EPS) / s
                mj = (1 - u_max) / (s - 1) - EPS
                u = torch.linspace(0, 1 - u_max, s)
            self._u_cache[key] = (u.to(device), float(mj))
        return self._u_cache[key]

    def _host_time(self, time):
        """``time`` only selects the state embedding: resolve a device tensor to a host float ONCE per call (every MLP of
        every level would otherwise synchronise on it), and not at all for single-state models."""
        if isinstance(time, torch.Tensor) and time.is_cuda and any(len(m.bkgd_stateembeds) > 1 for m in self.mlps):
            return float(time.reshape(-1)[0]) if time.numel() else 0.0
        return time

    def _background(self, randomized: bool) -> float:
        """Background intensity composited behind the last sample (S1 model.py:437-444): the midpoint of
        ``bg_intensity_range`` when it is not randomised."""
        lo, hi = self.bg_intensity_range[0], self.bg_intensity_range[1]
        if lo == hi:
            return float(lo)
        if randomized:
            raise NotImplementedError("hosnerf_b200: random background intensity is not built")
        return (lo + hi) / 2.0

    def _check_noise(self, randomized: bool):
        if randomized and any(m.density_noise > 0 or m.bottleneck_noise > 0 for m in self.mlps):
            raise NotImplementedError("hosnerf_b200: density_noise / bottleneck_noise > 0 with randomized=True is not built "
                                      "(the reference adds rand_like noise, S1 model.py:226-241)")

    def forward(self, batch, train_frac, randomized, is_train, near, far, rands=None):
        """S1 model.py:331-461.  Under autograd with ``is_train`` the call is one ``train.RenderFn`` node: the weights of
        every level and the final rgb carry a ``grad_fn`` whose backward runs the library's composite / MLP backward
        kernels, so the reference's ``training_step`` works on this module unchanged."""
        rays_o = batch["rays_o"]
        if not rays_o.is_cuda:
            raise RuntimeError("hosnerf_b200.MipNeRF360: inputs must be CUDA tensors (no CPU fallback)")
        if torch.is_grad_enabled() and is_train and any(p.requires_grad for p in self.parameters()):
            return self._forward_autograd(batch, train_frac, randomized, near, far, rands)
        return self._forward_impl(batch, train_frac, randomized, is_train, near, far, rands)

    def _forward_autograd(self, batch, train_frac, randomized, near, far, rands):
        from . import train
        named = [(k, p) for k, p in self.named_parameters()]
        names = tuple(k for k, _ in named)
        res = train.RenderFn.apply(self, batch, train_frac, randomized, near, far, rands, names, *[p for _, p in named])
        L = self.num_levels
        ws = res[1:1 + L]
        aux = res[1 + L:1 + 4 * L]
        tds = res[1 + 4 * L:]
        n = batch["rays_o"].shape[0]
        bg = self._background(randomized)
        history, renderings = [], []
        for lvl in range(L):
            last = lvl == L - 1
            h = {"density": aux[3 * lvl], "rgb": aux[3 * lvl + 1], "sdist": aux[3 * lvl + 2], "weights": ws[lvl]}
            if self.stage3:
                h["tdist"] = tds[lvl]
            elif last:
                renderings.append({"rgb": res[0]})
            else:
                renderings.append({"rgb": (torch.clip(1 - ws[lvl].sum(-1, keepdim=True), min=0) * bg).expand(n, 3)})
            history.append(h)
        return renderings, history

    def _forward_impl(self, batch, train_frac, randomized, is_train, near, far, rands=None, train_ctx=None):
        rays_o = batch["rays_o"]
        precision = self.precision or _DEFAULT_PRECISION
        n = rays_o.shape[0]
        dev = rays_o.device
        rays_o = rays_o.contiguous().float()
        rays_d = batch["rays_d"].contiguous().float()
        viewdirs = batch["viewdirs"].contiguous().float()
        radii = batch["radii"].reshape(-1).contiguous().float()
        time = self._host_time(batch["times"] if self.stage3 else batch["times"][0:1])
        s_near, s_far = float(np.float32(1 / near)), float(np.float32(1 / far))
        if self.near_anneal_rate is None:
            lo = 0.0
        else:
            lo = max(min(1 - train_frac / self.near_anneal_rate, 1), 0)
        hi = 1.0
        # level-0 inputs (one interval [lo, hi] of weight 1 per ray) are constants of (n, lo, hi): built once, read-only
        k0 = (n, float(lo), float(hi), str(dev))
        if self._u_cache.get("lvl0_key") != k0:
            self._u_cache["lvl0_key"] = k0
            self._u_cache["lvl0"] = (torch.cat([torch.full((n, 1), lo, device=dev), torch.full((n, 1), hi, device=dev)], dim=-1),
                                     torch.ones(n, 1, device=dev))
        sdist, weights = self._u_cache["lvl0"]
        prod = 1
        history, renderings = [], []
        anneal = (self.anneal_slope * train_frac) / ((self.anneal_slope - 1) * train_frac + 1) \
            if self.anneal_slope > 0 else 1.0
        bg = self._background(randomized)
        self._check_noise(randomized)
        for lvl in range(self.num_levels):
            is_prop = lvl < self.num_levels - 1
            s = self.num_prop_samples if is_prop else self.num_nerf_samples
            dilation = self.dilation_bias + self.dilation_multiplier * (hi - lo) / prod
            prod *= s
            dilate = lvl > 0 and (self.dilation_bias > 0 or self.dilation_multiplier > 0)
            u_base, max_jitter = self._u_base(s, randomized, dev)
            jitter = None
            if randomized:
                d = 1 if self.single_jitter else s
                # the reference draws on the host: torch.rand(t.shape[:-1] + (d,))  (helper.py:325-328)
                if rands is None and getattr(self, "device_rng", False):
                    r = torch.rand(n, d, device=dev)         # device generator: no host draw / copy (CUDA-graph capturable)
                else:
                    r = torch.rand(n, d) if rands is None else rands[lvl]
                jitter = r.to(dev, torch.float32).contiguous()
            sdist, tdist = ops.resample_level(sdist, weights, dilate, dilation, float(anneal),
                                              float(self.resample_padding), u_base, jitter, max_jitter,
                                              float(lo), float(hi), s_near, s_far)
            if train_ctx is not None:          # fp16 layer GEMMs with saved activations (train.mlp_forward_train)
                from . import train
                mlp = self.mlps[lvl]
                density, rgb, c = train.mlp_forward_train(mlp, tdist, rays_o, rays_d, radii, viewdirs, mlp._state_index(time))
                train_ctx.append(c)
                if rgb is None:
                    rgb = torch.zeros(n, s, 3, device=dev)
            else:
                density, rgb = self.mlps[lvl].eval_samples(tdist, rays_o, rays_d, radii, viewdirs, time, precision)
            last = not is_prop
            weights, rgb_out = ops.composite_mip360(density, tdist, rays_d, rgb if (last and not self.stage3) else None,
                                                    self.opaque_background, bg)
            res = {"density": density, "rgb": rgb, "sdist": sdist, "weights": weights}
            if getattr(self, "_keep_tdist", False) or train_ctx is not None:
                res["_tdist"] = tdist          # for the backward of the composite (LitMipNeRF360.loss_gradients)
            if self.stage3:
                res["tdist"] = tdist
            else:
                if last:
                    renderings.append({"rgb": rgb_out})
                else:   # proposal levels render rgb = 0 -> background weight only (S1 model.py:446-453)
                    renderings.append({"rgb": (torch.clip(1 - weights.sum(-1, keepdim=True), min=0) * bg).expand(n, 3)})
            history.append(res)
        return renderings, history


    # ------------------------------------------------------------------ one C call per batch
    def fused_render_supported(self, precision=None) -> bool:
        """True when ``render_fused`` can replace ``forward`` for an rgb-only render: tcgen05 mode, every level's MLP
        runs with the fused IPE prologue (widths <= 256), stage-1 composite."""
        precision = precision or self.precision or _DEFAULT_PRECISION
        if precision != "fp16" or self.stage3 or ops.PROFILE is not None or not FUSE_IPE:
            return False
        if self.num_levels > 4:                 # BKG_MAX_LEVELS of include/hosnerf_b200.h
            return False
        if self.bg_intensity_range[0] != self.bg_intensity_range[1]:
            return False
        for i, m in enumerate(self.mlps):
            if m.netwidth > 256 or m.pos_basis_t.shape[1] != 21 or m.max_deg_point - m.min_deg_point != 12 or m.min_deg_point != 0:
                return False
            if (i == self.num_levels - 1) == bool(m.disable_rgb):
                return False
        return True

    @staticmethod
    def fused_workspace(cfg, n, dev):
        from . import _lib
        nbytes = ctypes.c_size_t(0)
        _lib.call("hos_render_bkg_workspace", ctypes.byref(cfg), n, ctypes.byref(nbytes))
        return torch.empty(nbytes.value, dtype=torch.uint8, device=dev)

    def render_fused(self, batch, train_frac, randomized, near, far, rands=None, want_hist=False, mlp_events=None,
                     workspace=None):
        """The rgb of ``forward(...)[0][-1]`` through ONE library call (``hos_render_bkg``: the level loop of S1
        model.py:331-461 sequenced in C on the current stream).  Same kernels, same order, same results as
        ``forward``; what is saved is the per-level Python and allocator work.  Returns rgb [N,3] (and the final
        level's (sdist, weights) when ``want_hist``).  ``mlp_events``: optional list of 2 * num_levels recorded-once
        ``torch.cuda.Event(enable_timing=True)``; the library records them around each level's MLP launch."""
        from . import _lib
        rays_o = batch["rays_o"]
        if not rays_o.is_cuda:
            raise RuntimeError("hosnerf_b200.MipNeRF360: inputs must be CUDA tensors (no CPU fallback)")
        if not self.fused_render_supported():
            raise RuntimeError("hosnerf_b200.MipNeRF360.render_fused: configuration not covered (see fused_render_supported)")
        n, dev = rays_o.shape[0], rays_o.device
        rays_o = rays_o.contiguous().float()
        rays_d = batch["rays_d"].contiguous().float()
        viewdirs = batch["viewdirs"].contiguous().float()
        radii = batch["radii"].reshape(-1).contiguous().float()
        time = self._host_time(batch["times"][0:1])
        lo = 0.0 if self.near_anneal_rate is None else max(min(1 - train_frac / self.near_anneal_rate, 1), 0)
        anneal = (self.anneal_slope * train_frac) / ((self.anneal_slope - 1) * train_frac + 1) if self.anneal_slope > 0 else 1.0
        keep = []                                  # tensors the config points into
        cfg = _lib.BkgConfig()
        cfg.n_levels = self.num_levels
        prod = 1
        for lvl, m in enumerate(self.mlps):
            st = m._state_index(time)
            ver = m._versions()
            mlp = m._fused(st, ver)
            L = cfg.levels[lvl]
            L.mlp = mlp._h
            s = self.num_prop_samples if lvl < self.num_levels - 1 else self.num_nerf_samples
            u_base, mj = self._u_base(s, randomized, dev)
            L.n_samples, L.u_base, L.max_jitter = s, u_base.data_ptr(), mj
            L.dilation = self.dilation_bias + self.dilation_multiplier * (1.0 - lo) / prod
            L.dilate = int(lvl > 0 and (self.dilation_bias > 0 or self.dilation_multiplier > 0))
            prod *= s
            if randomized:
                d = 1 if self.single_jitter else s
                r = (torch.rand(n, d) if rands is None else rands[lvl]).to(dev, torch.float32).contiguous()
                keep.append(r)
                L.jitter, L.jitter_cols = r.data_ptr(), d
            if not m.disable_rgb:
                f = m._folded(st, ver)
                L.view_W, L.view_b, L.view_dim = f["views"][3].data_ptr(), mlp.view_bias.data_ptr(), m.netwidth_condition
                cfg.deg_view = m.deg_view
            cfg.basis_host = mlp.basis_host
        cfg.s_near, cfg.s_far = float(np.float32(1 / near)), float(np.float32(1 / far))
        cfg.dom_lo, cfg.dom_hi = float(lo), 1.0
        cfg.anneal, cfg.resample_padding = float(anneal), float(self.resample_padding)
        if mlp_events is not None:
            assert len(mlp_events) == 2 * self.num_levels
            evs = (ctypes.c_void_p * len(mlp_events))(*[e.cuda_event for e in mlp_events])
            keep.append(evs)
            cfg.mlp_events = ctypes.cast(evs, ctypes.POINTER(ctypes.c_void_p))
        cfg.opaque_background, cfg.bg = int(self.opaque_background), float(self.bg_intensity_range[0])
        rgb = torch.empty(n, 3, device=dev)
        s_last = self.num_nerf_samples
        sd = torch.empty(n, s_last + 1, device=dev) if want_hist else None
        wt = torch.empty(n, s_last, device=dev) if want_hist else None
        if n:
            ws = workspace                  # a caller that captures this call in a CUDA graph owns its workspace
            if ws is None:
                wkey = ("ws", n, str(dev), self.num_levels, self.num_prop_samples, self.num_nerf_samples)
                if self._u_cache.get("ws_key") != wkey:
                    self._u_cache["ws_key"], self._u_cache["ws"] = wkey, self.fused_workspace(cfg, n, dev)
                ws = self._u_cache["ws"]
            _lib.call("hos_render_bkg", ctypes.byref(cfg), rays_o.data_ptr(), rays_d.data_ptr(), viewdirs.data_ptr(),
                      radii.data_ptr(), n, ws.data_ptr(), ws.numel(), rgb.data_ptr(),
                      sd.data_ptr() if want_hist else None, wt.data_ptr() if want_hist else None,
                      ops._stream())
            _lib.LAUNCHES += 1 + 3 * self.num_levels + 2      # level-0 histogram; resample + MLP + composite per level; view term
        return (rgb, sd, wt) if want_hist else rgb


try:
    import pytorch_lightning as _pl  # type: ignore
    _LitBase = _pl.LightningModule
except Exception:  # pragma: no cover - Lightning is absent in the build image
    _LitBase = nn.Module


class LitMipNeRF360(_LitBase):
    """S1 model.py:464-627: the ``render_rays`` call surface named by BASELINE.json.  Metrics,
    image dumping and the optimiser schedule are the reference's own glue and stay there
    (INTEGRATION.md shows the two-line swap in ``utils/select_option.py``)."""

    def __init__(self, basedir, lr_init: float = 2.0e-3, lr_final: float = 2.0e-5, lr_delay_steps: int = 512,
                 lr_delay_mult: float = 0.01, data_loss_mult: float = 1.0, interlevel_loss_mult: float = 1.0,
                 distortion_loss_mult: float = 0.01, use_multiscale: bool = False, charb_padding: float = 0.001,
                 **model_kwargs):
        for name, value in vars().items():
            if name not in ["self", "__class__", "model_kwargs"]:
                setattr(self, name, value)
        super().__init__()
        self.model = MipNeRF360(basedir, **model_kwargs)
        self.near, self.far = 0.1, 1e6
        self._train_frac = 1.0

    def setup(self, stage=None):
        dm = self.trainer.datamodule
        self.near, self.far, self.white_bkgd = dm.near, dm.far, getattr(dm, "white_bkgd", False)

    def _frac(self):
        tr = getattr(self, "trainer", None) if isinstance(self, nn.Module) else None
        try:
            return self.global_step / self.trainer.max_steps
        except Exception:
            return self._train_frac

    def render_rays(self, batch, batch_idx):
        with torch.no_grad():
            if self.model.fused_render_supported():        # one library call for the whole level loop
                rgb = self.model.render_fused(batch, self._frac(), False, self.near, self.far)
            else:
                rendered, _ = self.model(batch, self._frac(), False, False, self.near, self.far)
                rgb = rendered[-1]["rgb"]
        return {"target": batch.get("target"), "rgb": rgb}

    _RAY_KEYS = ("rays_o", "rays_d", "viewdirs", "radii")

    def _graphed_render(self, hb, dev):
        """Copy one HOST batch into the static inputs of a captured CUDA graph of ``render_fused`` and replay it (the
        step is nine dependent launches: one graph launch instead).  Returns the graph's static rgb [n,3]."""
        from . import _lib
        m = self.model
        n = hb["rays_o"].shape[0]
        frac = self._frac()
        times = hb["times"]                                   # stays on the host: only selects the state embedding
        states = tuple(x._state_index(times[0:1]) for x in m.mlps)
        key = (n, str(dev), frac, self.near, self.far, states, tuple(x._versions() for x in m.mlps))
        cache = self.__dict__.setdefault("_graphs", {})
        ent = cache.get(key)
        if ent is None:
            if len(cache) >= 4:                               # a render has one or two batch sizes; weights change rarely
                cache.pop(next(iter(cache)))
            static = {k: torch.empty(tuple(hb[k].shape), dtype=torch.float32, device=dev) for k in self._RAY_KEYS}
            static["times"] = times
            for k in self._RAY_KEYS:
                static[k].copy_(hb[k], non_blocking=True)
            launches0 = _lib.LAUNCHES
            m.render_fused(static, frac, False, self.near, self.far)     # packs weights / sizes the workspace before capture
            per_replay = _lib.LAUNCHES - launches0
            ws = torch.empty_like(m._u_cache["ws"])          # the graph's own workspace: never resized under it
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):     # other threads (NCCL watchdog) may call CUDA
                rgb = m.render_fused(static, frac, False, self.near, self.far, workspace=ws)
            _lib.LAUNCHES -= per_replay                      # the capture pass launched nothing
            # The graph bakes in raw device pointers: packed weights / parameter blocks of each level's FusedMLP, its folded
            # view-term tensors, the quantile tables.  The entry owns references to all of them, so a later switch to
            # another state embedding (or a parameter update) can never free memory this graph still reads.
            holds = [x._fused(st, x._versions()) for x, st in zip(m.mlps, states)]
            holds += [m._u_cache[k] for k in list(m._u_cache) if isinstance(k, tuple)]
            ent = cache[key] = (g, static, rgb, per_replay, ws, holds)
        g, static, rgb, per_replay = ent[:4]
        for k in self._RAY_KEYS:
            static[k].copy_(hb[k], non_blocking=True)
        g.replay()
        _lib.LAUNCHES += per_replay
        return rgb

    def render_rays_stream(self, host_batches, depth: int = 2, graph: bool = True):
        """``render_rays`` over an iterable of HOST ray batches (pinned tensors), as the eval loops do chunk by chunk
        (S1 model.py:516-560), without a host/device round trip per chunk: the copy-in and the kernels of chunk i+1 are
        enqueued before the rgb of chunk i is awaited.  With ``graph`` (and a model whose levels all run on the fused
        tcgen05 path) every chunk is one CUDA-graph replay of ``hos_render_bkg`` on static input buffers.  Yields one
        pinned host tensor [n,3] per batch, in order; a yielded tensor is a staging buffer that is overwritten once the
        next item is requested, so copy what you keep."""
        dev = next(self.model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("hosnerf_b200: render_rays_stream needs the module on a CUDA device (no CPU path)")
        slots, pending = [None] * depth, []
        for i, hb in enumerate(host_batches):
            with torch.no_grad():
                if graph and hb["rays_o"].shape[0] > 0 and self.model.fused_render_supported():
                    rgb = self._graphed_render(hb, dev)
                else:
                    batch = {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) and k != "times" else v)
                             for k, v in hb.items()}
                    rgb = self.render_rays(batch, i)["rgb"]
            j = i % depth
            if len(pending) == depth:             # slot j still belongs to chunk i - depth: hand it out first
                ev, out = pending.pop(0)
                ev.synchronize()
                yield out
            if slots[j] is None or slots[j].shape != rgb.shape:
                slots[j] = torch.empty(rgb.shape, dtype=rgb.dtype, pin_memory=True)
            slots[j].copy_(rgb, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending.append((ev, slots[j]))
        for ev, out in pending:
            ev.synchronize()
            yield out

    def validation_step(self, batch, batch_idx):
        return self.render_rays(batch, batch_idx)

    def test_step(self, batch, batch_idx):
        return self.render_rays(batch, batch_idx)

    def interlevel_loss(self, ray_history):
        """S1 model.py:609-618: sum over proposal levels of mean(lossfun_outer(final level | level))."""
        c, w = ray_history[-1]["sdist"].contiguous(), ray_history[-1]["weights"].contiguous()
        total = None
        for lvl in ray_history[:-1]:
            _, rows = ops.lossfun_outer(c, w, lvl["sdist"].contiguous(), lvl["weights"].contiguous(), want_rows=True)
            term = ops.reduce_scaled(rows, 1.0 / max(w.numel(), 1))
            total = term if total is None else total + term
        return total if total is not None else torch.zeros((), device=w.device)

    def distortion_loss(self, ray_history):
        """S1 model.py:620-625: mean over rays of lossfun_distortion(final level)."""
        c, w = ray_history[-1]["sdist"].contiguous(), ray_history[-1]["weights"].contiguous()
        return ops.reduce_scaled(ops.lossfun_distortion(c, w), 1.0 / max(w.shape[0], 1))

    def loss_terms(self, batch, randomized: bool = True, rands=None):
        """Forward value of the stage-1 training objective (S1 model.py:488-512) on one ray batch:
        Charbonnier data term + interlevel + distortion, and the PSNR the reference logs.  No autograd graph
        is built (backward kernels are not part of this round), so this evaluates the loss, it does not train."""
        with torch.no_grad():
            rendered, hist = self.model(batch, self._frac(), randomized, True, self.near, self.far, rands=rands)
            rgb = rendered[-1]["rgb"].contiguous()
            mse = ops.reduce_scaled(rgb, 1.0 / max(rgb.numel(), 1), y=batch["target"].to(rgb.dtype).contiguous())
            data = torch.sqrt(mse + self.charb_padding ** 2) * self.data_loss_mult
            inter = self.interlevel_loss(hist)
            dist = self.distortion_loss(hist)
            loss = data + inter * self.interlevel_loss_mult + dist * self.distortion_loss_mult
            psnr = -10.0 * torch.log(mse) / math.log(10.0)
        return {"loss": loss, "rgbloss": mse, "interlevel": inter, "distortion": dist, "psnr": psnr}

    def loss_gradients(self, batch, randomized: bool = True, rands=None):
        """The loss side of the stage-1 backward pass (S1 model.py:488-512), on device kernels: forward render, the
        objective, and its gradient with respect to what the MLPs produced - per level dL/ddensity [N,S] and, for the
        final level, dL/drgb [N,S,3].  Chain: loss terms -> dL/dweights (``hos_lossfun_*_backward``) and
        dL/d(composited rgb) -> ``hos_composite_mip360_backward``.  The sample positions are constants (the reference
        detaches them, model.py:405-406).  The MLP dgrad / wgrad that would consume these gradients is not built yet."""
        m = self.model
        with torch.no_grad():
            m._keep_tdist = True
            try:
                rendered, hist = m(batch, self._frac(), randomized, True, self.near, self.far, rands=rands)
            finally:
                m._keep_tdist = False
            rgb = rendered[-1]["rgb"].contiguous()
            target = batch["target"].to(rgb.dtype).contiguous()
            n = rgb.shape[0]
            mse = ops.reduce_scaled(rgb, 1.0 / max(rgb.numel(), 1), y=target)
            data = torch.sqrt(mse + self.charb_padding ** 2) * self.data_loss_mult
            inter, dist = self.interlevel_loss(hist), self.distortion_loss(hist)
            loss = data + inter * self.interlevel_loss_mult + dist * self.distortion_loss_mult
            # d(data term)/d(composited rgb) = mult / (2 sqrt(mse + pad^2)) * 2 (rgb - target) / numel
            g_out = ((rgb - target) * (self.data_loss_mult / (torch.sqrt(mse + self.charb_padding ** 2) * rgb.numel()))).contiguous()
            c, w = hist[-1]["sdist"].contiguous(), hist[-1]["weights"].contiguous()
            rays_d = batch["rays_d"].contiguous().float()
            bg = m._background(randomized)
            grads = []
            for lvl, h in enumerate(hist):
                last = lvl == len(hist) - 1
                tdist = h.pop("_tdist")
                if last:
                    g_w = ops.lossfun_distortion_backward(c, w, g_scalar=self.distortion_loss_mult / max(n, 1))
                    gd, gc = ops.composite_mip360_backward(h["density"].contiguous(), tdist, rays_d, h["rgb"].contiguous(), g_w, g_out,
                                                           m.opaque_background, bg)
                    grads.append({"density": gd, "rgb": gc})
                else:
                    g_w = ops.lossfun_outer_backward(c, w, h["sdist"].contiguous(), h["weights"].contiguous(),
                                                     g_scalar=self.interlevel_loss_mult / max(w.numel(), 1))
                    gd, _ = ops.composite_mip360_backward(h["density"].contiguous(), tdist, rays_d, None, g_w, None,
                                                          m.opaque_background, bg)
                    grads.append({"density": gd})
        return {"loss": loss, "rgbloss": mse, "interlevel": inter, "distortion": dist, "ray_history": hist, "grads": grads}

    def training_objective(self, batch, randomized: bool = True, rands=None):
        """The stage-1 objective of S1 model.py:491-514 with an autograd graph: Charbonnier data term + interlevel +
        distortion.  ``loss.backward()`` runs ``train.RenderFn.backward`` (composite / MLP backward kernels)."""
        from . import train
        rendered, hist = self.model(batch, self._frac(), randomized, True, self.near, self.far, rands=rands)
        rgb = rendered[-1]["rgb"]
        target = batch["target"].to(rgb.dtype)
        mse = torch.mean((rgb - target) ** 2)
        loss = torch.sqrt(mse + self.charb_padding ** 2) * self.data_loss_mult
        c, w = hist[-1]["sdist"].detach(), hist[-1]["weights"].detach()
        inter = 0.0
        for lvl in hist[:-1]:
            inter = inter + train.lossfun_outer_sum(c, w, lvl["sdist"].detach(), lvl["weights"]) / max(w.numel(), 1)
        dist = train.lossfun_distortion(hist[-1]["sdist"].detach(), hist[-1]["weights"]).mean()
        loss = loss + inter * self.interlevel_loss_mult + dist * self.distortion_loss_mult
        return {"loss": loss, "rgbloss": mse.detach(), "interlevel": inter, "distortion": dist,
                "psnr": -10.0 * torch.log(mse.detach()) / math.log(10.0)}

    def training_step(self, batch, batch_idx):
        """S1 model.py:491-514."""
        out = self.training_objective(batch, randomized=True)
        if hasattr(self, "log") and getattr(self, "_trainer", None) is not None:
            self.log("train/loss", out["loss"].item(), on_step=True, prog_bar=True)
            self.log("train/psnr", out["psnr"].item(), on_step=True, prog_bar=True)
        return out["loss"]

    def lr_at(self, step: int, max_steps: int) -> float:
        """Learning-rate schedule of S1 model.py:541-569: log-linear decay lr_init -> lr_final with a sine warm-up."""
        if self.lr_delay_steps > 0:
            delay = self.lr_delay_mult + (1 - self.lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / self.lr_delay_steps, 0), 1))
        else:
            delay = 1.0
        t = min(max(step / max_steps, 0), 1)
        return delay * math.exp(math.log(self.lr_init) * (1 - t) + math.log(self.lr_final) * t)

    def optimizer_step(self, epoch=None, batch_idx=None, optimizer=None, optimizer_idx=None, optimizer_closure=None,
                       on_tpu=None, using_native_amp=None, using_lbfgs=None, step=None, max_steps=None):
        """S1 model.py:541-569 (``step`` / ``max_steps`` default to the Lightning trainer's / gin's values)."""
        if step is None:
            step = self.trainer.global_step
        if max_steps is None:
            try:
                import gin  # type: ignore
                max_steps = gin.query_parameter("run.max_steps")
            except Exception:
                max_steps = self.trainer.max_steps
        lr = self.lr_at(step, max_steps)
        for pg in optimizer.param_groups:
            pg["lr"] = lr
        optimizer.step(closure=optimizer_closure)

    def configure_optimizers(self):
        return torch.optim.Adam(params=self.parameters(), lr=self.lr_init, betas=(0.9, 0.999))

    # ------------------------------------------------------------------ eval tail (S1 src/model/interface.py:28-51)
    def alter_gather_cat(self, outputs, key, image_sizes):
        """Per-rank step outputs -> whole images: concatenate, all_gather over the ranks, undo the strided ray sharding
        (ray i rendered by rank i mod W) and cut into [h, w, 3] images (interface.py:28-39).  One NCCL all_gather."""
        from . import dist as hd
        each = torch.cat([o[key] for o in outputs])
        total = sum(h * w for h, w in image_sizes)
        full = hd.gather_rays(each, total, "strided").detach()
        ret, cur = [], 0
        for h, w in image_sizes:
            ret.append(full[cur:cur + h * w].reshape(h, w, 3))
            cur += h * w
        return ret

    @torch.no_grad()
    def psnr_each(self, preds, gts):
        """interface.py:41-51."""
        out = []
        for pred, gt in zip(preds, gts):
            mse = torch.mean((torch.clip(pred, 0, 1) - torch.clip(gt, 0, 1)) ** 2)
            out.append(-10.0 * torch.log(mse) / math.log(10.0))
        return torch.stack(out)

    def validation_epoch_end(self, outputs):
        """S1 model.py:571-581 without the metric packages this image lacks (SSIM / LPIPS need piqa + VGG weights): gathers the
        rendered images of all ranks and returns / logs the mean PSNR."""
        sizes = self.trainer.datamodule.val_image_sizes
        rgbs = self.alter_gather_cat(outputs, "rgb", sizes)
        targets = self.alter_gather_cat(outputs, "target", sizes)
        psnr = self.psnr_each(rgbs, targets).mean()
        if hasattr(self, "log"):
            self.log("val/psnr", psnr.item(), on_epoch=True, sync_dist=True)
        return psnr
