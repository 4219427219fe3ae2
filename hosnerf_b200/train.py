"""Training path of the background branch: forward with saved activations and the hand-written backward of
``MipNeRF360MLP`` (S1 model.py:212-259) on the tensor cores - what the reference gets from autograd when
``training_step`` (S1 model.py:491-514) calls ``loss.backward()``.

Every 256/1024-wide layer is one ``hos_gemm_tma`` launch over row-major fp16 activations (forward: bias + ReLU in the
epilogue, density / rgb heads from the fp32 accumulator; data gradient: W consumed as an MN-major operand, ReLU mask in the
epilogue) and one ``hos_wgrad_tma`` launch (reduction over the rows through MN-major descriptors, fp32 atomics); bias
gradients and the <= 4-wide heads run on two small SIMT kernels.  Gradients travel in fp16 with a power-of-two loss scale
chosen on the device (no host synchronisation) and are unscaled in fp32.

The sample positions are constants of the backward pass (``stop_level_grad``, S1 model.py:405-406) and the Gaussians are
detached by the reference itself (helper.py:57-60), so nothing flows into the sampler or the encoder: the chain is
loss -> composite -> (density, rgb) -> MLP parameters (+ the state embedding, through the bias it is folded into).
"""
from __future__ import annotations

import torch

from . import ops

_F16 = torch.float16


def _h(t):
    return t.detach().to(_F16).contiguous()


def mlp_forward_train(m, tdist, rays_o, rays_d, radii, viewdirs, state_idx: int):
    """One MipNeRF360MLP on [N, S] intervals with everything the backward needs kept: returns (density [N,S], rgb [N,S,3]
    or None, ctx)."""
    n, s = tdist.shape[0], tdist.shape[1] - 1
    rows = n * s
    F, nw = m.ipe_size, m.netwidth
    e = m.bkgd_stateembeds[state_idx].detach()
    feat = ops.ipe_features(tdist, rays_o, rays_d, radii, m.pos_basis_t, m.min_deg_point, m.max_deg_point, "f16op")
    hs, x = [], feat
    dens = None
    nl = len(m.pts_linear)
    w16 = []                     # fp16 copy of every layer's hidden-input weight block: reused by the data-gradient GEMMs
    for i, lin in enumerate(m.pts_linear):
        W, b = lin.weight.detach(), lin.bias.detach()
        last = i == nl - 1
        kw = dict(relu=True)
        if last:
            kw["head"] = (m.density_layer.weight.detach().contiguous(), m.density_layer.bias.detach().contiguous(), 1,
                          float(m.density_bias))
        if i == 0:
            wh = _h(W[:, :F])
            res = ops.gemm_tma(x, wh, nw, bias=(b + W[:, F:] @ e).contiguous(), **kw)
        elif m._skip_inputs(i):
            wh = _h(W[:, :nw])
            res = ops.gemm_tma(x, wh, nw, a1=feat, w1=_h(W[:, nw:nw + F]), bias=(b + W[:, nw + F:] @ e).contiguous(), **kw)
        else:
            wh = _h(W)
            res = ops.gemm_tma(x, wh, nw, bias=b.contiguous(), **kw)
        w16.append(wh)
        x = res[0]
        hs.append(x)
        if last:
            dens = res[3]
    ctx = {"m": m, "state": state_idx, "n": n, "s": s, "feat": feat, "hs": hs, "density": dens, "w16": w16}
    if m.disable_rgb:
        return dens.view(n, s), None, ctx
    bw = m.bottleneck_width
    wb16 = _h(m.bottleneck_layer.weight)
    bott = ops.gemm_tma(x, wb16, bw, bias=m.bottleneck_layer.bias.detach().contiguous())[0]
    de = ops.pos_enc(viewdirs, 0, m.deg_view, True)
    Wv, bv = m.views_linear[0].weight.detach(), m.views_linear[0].bias.detach()
    rowterm = ops.linear_f32(de, Wv[:, bw:].contiguous(), bv.contiguous())
    Wr = (m.rgb_layer.weight.detach() * m.rgb_premultiplier).contiguous()
    br = (m.rgb_layer.bias.detach() * m.rgb_premultiplier + m.rgb_bias).contiguous()
    wv16 = _h(Wv[:, :bw])
    v, _, _, rgb = ops.gemm_tma(bott, wv16, m.netwidth_condition, relu=True, rowbias=rowterm, rowbias_div=s,
                                head=(Wr, br, 2, float(m.rgb_padding)))
    ctx.update(bott=bott, de=de, v=v, rgb=rgb, wb16=wb16, wv16=wv16)
    return dens.view(n, s), rgb.view(n, s, 3), ctx


def _loss_scale(*gs):
    """Power-of-two scale (device tensor) that lifts the largest upstream gradient to ~2^8: the fp16 gradient planes keep
    ~3.5 decades below it before flushing and two decades of head-room above it."""
    mx = torch.stack([g.detach().abs().max() for g in gs if g is not None]).max().clamp_min(1e-30)
    return torch.exp2(torch.floor(8.0 - torch.log2(mx))).clamp(2.0 ** -20, 2.0 ** 40)


def mlp_backward(ctx, g_density, g_rgb, grads: dict):
    """Accumulate parameter gradients of one MLP into ``grads`` (name -> fp32 tensor shaped like the parameter), given
    dL/ddensity [N,S] and dL/drgb [N,S,3] (or None for a proposal MLP)."""
    m = ctx["m"]
    rows = ctx["n"] * ctx["s"]
    F, nw = m.ipe_size, m.netwidth
    dev = g_density.device
    feat, hs = ctx["feat"], ctx["hs"]
    h_last = hs[-1]
    e = m.bkgd_stateembeds[ctx["state"]].detach()

    scale = _loss_scale(g_density, g_rgb)
    inv = 1.0 / scale
    # density = softplus(raw + bias): d/draw = sigmoid = 1 - exp(-density)
    g_raw = (g_density.reshape(rows, 1) * (1.0 - torch.exp(-ctx["density"])) * scale).contiguous()
    Wd = m.density_layer.weight.detach().contiguous()
    # scaled gradients of this MLP: views of ONE zeroed buffer (one memset, one unscaling multiply at the end)
    own = [(k, p) for k, p in m.named_parameters() if p.requires_grad]
    flat = torch.zeros(sum(p.numel() for _, p in own), device=dev, dtype=torch.float32)
    sg, off = {}, 0
    for k, p in own:
        sg[k] = flat[off:off + p.numel()].view(p.shape)
        off += p.numel()
    touched = set()

    def sacc(name, shape):
        touched.add(name)
        return sg[name]

    ops.colsum_f16(h_last, sacc("density_layer.weight", Wd.shape), g=g_raw)
    sacc("density_layer.bias", (1,)).add_(g_raw.sum(0))
    add = None
    if g_rgb is not None and not m.disable_rgb:
        bw, cw = m.bottleneck_width, m.netwidth_condition
        pad = float(m.rgb_padding)
        sgm = (ctx["rgb"] + pad) / (1.0 + 2.0 * pad)
        g_pre = (g_rgb.reshape(rows, 3) * ((1.0 + 2.0 * pad) * scale) * sgm * (1.0 - sgm)).contiguous()     # wrt premult * lin + bias
        Wr_eff = (m.rgb_layer.weight.detach() * m.rgb_premultiplier).contiguous()
        v, bott, de = ctx["v"], ctx["bott"], ctx["de"]
        gWr = torch.zeros(3, cw, device=dev)
        ops.colsum_f16(v, gWr, g=g_pre)
        sacc("rgb_layer.weight", (3, cw)).add_(gWr * m.rgb_premultiplier)
        sacc("rgb_layer.bias", (3,)).add_(g_pre.sum(0) * m.rgb_premultiplier)
        g_zv = ops.head_dgrad(g_pre, Wr_eff, cw, mask=v)                                    # [rows, cw] fp16, ReLU-masked
        Wv = m.views_linear[0].weight.detach()
        gWv = sacc("views_linear.0.weight", Wv.shape)
        ops.wgrad_tma(bott, g_zv, gWv[:, :bw], transpose_out=True)
        g_row = g_zv.view(ctx["n"], ctx["s"], cw).sum(1, dtype=torch.float32)               # per-ray term
        gWv[:, bw:].add_(g_row.t() @ de)
        sacc("views_linear.0.bias", (cw,)).add_(g_row.sum(0))
        g_bott = ops.gemm_tma(g_zv, ctx["wv16"], bw, mode=1)[0]                              # [rows, bw]
        Wb = m.bottleneck_layer.weight.detach()
        ops.wgrad_tma(g_bott, h_last, sacc("bottleneck_layer.weight", Wb.shape), colsum=sacc("bottleneck_layer.bias", (bw,)))
        add = ops.gemm_tma(g_bott, ctx["wb16"], nw, mode=1)[0]                               # d/dh_last through the bottleneck
    g_z = ops.head_dgrad(g_raw, Wd, nw, add=add, mask=h_last)                                # + density head, ReLU mask of the last layer
    g_e = sacc("bkgd_stateembeds.%d" % ctx["state"], e.shape)
    for i in range(len(m.pts_linear) - 1, -1, -1):
        lin = m.pts_linear[i]
        W = lin.weight.detach()
        gW = sacc(f"pts_linear.{i}.weight", W.shape)
        gb = sacc(f"pts_linear.{i}.bias", (nw,))         # bias gradient = column sums of dL/dZ: rides in the weight-gradient pass
        if i == 0:
            ops.wgrad_tma(g_z, feat, gW[:, :F], colsum=gb)
            emb0 = F
        elif m._skip_inputs(i):
            ops.wgrad_tma(g_z, hs[i - 1], gW[:, :nw], colsum=gb)
            ops.wgrad_tma(g_z, feat, gW[:, nw:nw + F])
            emb0 = nw + F
        else:
            ops.wgrad_tma(g_z, hs[i - 1], gW, colsum=gb)
            emb0 = None
        if emb0 is not None:        # the embedding rides in the bias: rank-1 weight gradient, and its own gradient
            gW[:, emb0:].add_(torch.outer(gb, e))
            g_e.add_(W[:, emb0:].t() @ gb)
        if i > 0:
            g_z = ops.gemm_tma(g_z, ctx["w16"][i], nw, mode=1, mask=hs[i - 1])[0]
    flat.mul_(inv)
    for k in touched:
        if k in grads:
            grads[k] = grads[k] + sg[k]
        else:
            grads[k] = sg[k]
    return grads


class RenderFn(torch.autograd.Function):
    """``MipNeRF360.forward`` as one autograd node: outputs the composited rgb of the final level and every level's weights
    (what the stage-1 objective reads, S1 model.py:491-514, 609-625); the backward runs composite backward -> MLP backward on
    the library's kernels and hands one gradient per parameter to autograd, so ``loss.backward()``, Lightning and any torch
    optimiser work unchanged."""

    @staticmethod
    def forward(ctx, model, batch, train_frac, randomized, near, far, rands, names, *params):
        saved = []
        rend, hist = model._forward_impl(batch, train_frac, randomized, True, near, far, rands, train_ctx=saved)
        ctx.model, ctx.saved, ctx.names = model, saved, names
        ctx.rays_d = batch["rays_d"].contiguous().float()
        ctx.bg = model._background(randomized)
        ctx.levels = len(hist)
        tdists = [h.pop("_tdist") for h in hist]
        # the context keeps detached aliases: the tensors returned below become outputs of this node, and a context that
        # holds its own outputs is a reference cycle - the node (and the parameters' AccumulateGrad nodes, with the stream
        # they were created on) would outlive the step until the garbage collector runs
        ctx.tdists = [t.detach() for t in tdists]
        ctx.hist = [{k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in h.items()} for h in hist]
        ctx.stage3 = model.stage3
        L = len(hist)
        aux = []
        for h in hist:
            aux += [h["density"], h["rgb"], h["sdist"]]
        rgb_final = rend[-1]["rgb"] if not model.stage3 else torch.zeros(0, device=ctx.rays_d.device)
        # stage 3 composites outside (together with the human samples): there the final level's per-sample density / rgb
        # carry the gradient instead of the composited colour
        nondiff = [a for i, a in enumerate(aux) if not (model.stage3 and i in (3 * (L - 1), 3 * (L - 1) + 1))] + tdists
        if model.stage3:
            nondiff.append(rgb_final)
        ctx.mark_non_differentiable(*nondiff)
        return tuple([rgb_final] + [h["weights"] for h in hist] + aux + tdists)

    @staticmethod
    def backward(ctx, *gouts):
        m = ctx.model
        L = ctx.levels
        grads = {}
        g_out, g_ws = gouts[0], gouts[1:1 + L]
        g_dens_last = g_rgb_last = None
        if ctx.stage3:
            g_out = None
            g_dens_last, g_rgb_last = gouts[1 + L + 3 * (L - 1)], gouts[1 + L + 3 * (L - 1) + 1]
        sink = getattr(m, "_grad_sink", None)        # dist.FlatGrads: accumulate in place + bucketed asynchronous all-reduce
        for lvl in range(L - 1, -1, -1):             # the NeRF MLP (largest bucket) first: its collective overlaps the rest
            h = ctx.hist[lvl]
            last = lvl == L - 1
            g_w = g_ws[lvl]
            gd = gc = None
            want_rgb = last and not ctx.stage3 and g_out is not None
            if g_w is not None or want_rgb:
                g_w = (g_w if g_w is not None else torch.zeros_like(h["weights"])).contiguous().float()
                gd, gc = ops.composite_mip360_backward(h["density"].contiguous(), ctx.tdists[lvl], ctx.rays_d,
                                                       h["rgb"].contiguous() if want_rgb else None, g_w,
                                                       g_out.contiguous().float() if want_rgb else None, m.opaque_background, ctx.bg)
            if last and ctx.stage3:
                if g_dens_last is not None:
                    gd = g_dens_last.contiguous().float() if gd is None else gd + g_dens_last
                gc = None if g_rgb_last is None else g_rgb_last.contiguous().float()
            if gd is None and gc is None:
                continue
            if gd is None:
                gd = torch.zeros_like(h["density"])
            sub = {}
            mlp_backward(ctx.saved[lvl], gd.contiguous(), gc, sub)
            for k, v in sub.items():
                if sink is not None:
                    sink.add_(f"mlps.{lvl}.{k}", v)
                else:
                    grads[f"mlps.{lvl}.{k}"] = v
            if sink is not None:
                sink.reduce_bucket(lvl)
        ctx.saved = None
        return (None,) * 8 + tuple(grads.get(nm) for nm in ctx.names)


class _DistortionFn(torch.autograd.Function):
    """helper.lossfun_distortion (S1 helper.py:122-128) per ray, gradient w.r.t. the weights only (positions are detached)."""

    @staticmethod
    def forward(ctx, t, w):
        t, w = t.contiguous(), w.contiguous()
        ctx.save_for_backward(t, w)
        return ops.lossfun_distortion(t, w)

    @staticmethod
    def backward(ctx, g):
        t, w = ctx.saved_tensors
        return None, ops.lossfun_distortion_backward(t, w, 1.0, g_ray=g.contiguous())


class _OuterSumFn(torch.autograd.Function):
    """sum(helper.lossfun_outer(t, w, t_env, w_env)) (S1 helper.py:92-120), gradient w.r.t. the envelope weights."""

    @staticmethod
    def forward(ctx, t, w, t_env, w_env):
        t, w, t_env, w_env = t.contiguous(), w.contiguous(), t_env.contiguous(), w_env.contiguous()
        ctx.save_for_backward(t, w, t_env, w_env)
        _, rows = ops.lossfun_outer(t, w, t_env, w_env, want_rows=True)
        return ops.reduce_scaled(rows, 1.0)

    @staticmethod
    def backward(ctx, g):
        t, w, t_env, w_env = ctx.saved_tensors
        return None, None, None, ops.lossfun_outer_backward(t, w, t_env, w_env, 1.0) * g


def lossfun_distortion(t, w):
    return _DistortionFn.apply(t, w)


def lossfun_outer_sum(t, w, t_env, w_env):
    return _OuterSumFn.apply(t, w, t_env, w_env)


# ============================================================================ human-object branch (S2 / S3 network.py)
class LbsWarpFn(torch.autograd.Function):
    """``Network._sample_motion_fields`` (network.py:304-354): forward and backward on the library's kernels.  Gradient to
    the motion-weight volume and the bone maps; the sample points are data."""

    @staticmethod
    def forward(ctx, pts, R, T, vol, bbox_min, bbox_scale):
        pts, R, T, vol = pts.contiguous(), R.contiguous(), T.contiguous(), vol.contiguous()
        ctx.save_for_backward(pts, R, T, vol)
        ctx.bbox = (bbox_min, bbox_scale)
        return ops.lbs_warp(pts, R, T, vol, bbox_min, bbox_scale)

    @staticmethod
    def backward(ctx, g_x, g_m):
        pts, R, T, vol = ctx.saved_tensors
        g_x = torch.zeros(pts.numel() // 3, 3, device=pts.device) if g_x is None else g_x.contiguous().float()
        g_vol, g_R, g_T = ops.lbs_warp_backward(pts, R, T, vol, ctx.bbox[0], ctx.bbox[1], g_x,
                                                None if g_m is None else g_m.contiguous().float())
        return None, g_R, g_T, g_vol, None, None


class LbsForwardFn(torch.autograd.Function):
    """``Network._sample_motion_fields_forward`` (network.py:357-398) of the cycle / flow side paths: gradient to the volume, the
    forward bone maps and the canonical points."""

    @staticmethod
    def forward(ctx, cnl_pts, R, T, vol, bbox_min, bbox_scale):
        cnl_pts, R, T, vol = cnl_pts.contiguous(), R.contiguous(), T.contiguous(), vol.contiguous()
        ctx.save_for_backward(cnl_pts, R, T, vol)
        ctx.bbox = (bbox_min, bbox_scale)
        return ops.lbs_forward(cnl_pts, R, T, vol, bbox_min, bbox_scale)[0]

    @staticmethod
    def backward(ctx, g_x):
        pts, R, T, vol = ctx.saved_tensors
        g_vol, g_R, g_T, g_pts = ops.lbs_forward_backward(pts, R, T, vol, ctx.bbox[0], ctx.bbox[1], g_x.contiguous().float())
        return g_pts, g_R, g_T, g_vol, None, None


def _pad8(t):
    """fp16 copy of a matrix with its column count padded to a multiple of 8 (16-byte row pitch for the tensor maps)."""
    k = t.shape[1]
    kp = (k + 7) // 8 * 8
    out = torch.zeros(t.shape[0], kp, device=t.device, dtype=_F16)
    out[:, :k] = t
    return out


class MlpFn(torch.autograd.Function):
    """A ReLU MLP with one skip concatenation and a <= 4-wide linear head, forward and backward on the tensor-core layer
    kernels: the canonical MLP (mlp_rgb_sigma.py:16-58) and the non-rigid motion MLPs (mlp_offset.py:16-70).

    ``x`` [P, kx] is the per-point encoded input (differentiable); per-call constants that the reference concatenates to
    it (pose condition code, state embedding) ride in the bias of the layers that read them - ``const`` [kc] with its
    column range in those layers' weights - and receive their gradient from the bias gradient.
    spec: column ranges inside the weights:  first layer ``x_cols`` / ``c_cols``;  skip layer (index ``skip``) ``sh_cols``
    (hidden part), ``sx_cols`` (encoded input), ``sc_cols`` (constant part or None).
    Returns the head's pre-activation [P, n_out] (fp32)."""

    @staticmethod
    def forward(ctx, x, const, spec, *params):
        nl = (len(params) - 2) // 2
        Ws, bs = params[0:2 * nl:2], params[1:2 * nl:2]
        Wo, bo = params[-2], params[-1]
        width = Ws[0].shape[0]
        kx = x.shape[1]
        c = const.detach().reshape(-1).float()
        x16 = _pad8(x.detach())
        hs, h = [], None

        def cbias(W, b, cols):
            return (b if cols is None else b + W[:, cols[0]:cols[1]] @ c).contiguous()
        for i in range(nl):
            W, b = Ws[i].detach(), bs[i].detach()
            if i == 0:
                (xa, xb) = spec["x_cols"]
                h = ops.gemm_tma(x16, _pad8(W[:, xa:xb]), width, bias=cbias(W, b, spec["c_cols"]), relu=True)[0]
            elif i == spec["skip"]:
                (ha, hb), (xa, xb) = spec["sh_cols"], spec["sx_cols"]
                h = ops.gemm_tma(h, _h(W[:, ha:hb]), width, a1=x16, w1=_pad8(W[:, xa:xb]), bias=cbias(W, b, spec["sc_cols"]),
                                 relu=True)[0]
            else:
                h = ops.gemm_tma(h, _h(W), width, bias=b.contiguous(), relu=True)[0]
            hs.append(h)
        # the head is <= 4 wide: evaluated from the fp16 activations in fp32 (SIMT kernel)
        out = ops.head_f32(h.float(), Wo.detach().contiguous(), bo.detach().contiguous(), post=0)
        ctx.spec, ctx.nl, ctx.kx = spec, nl, kx
        ctx.hs, ctx.x16, ctx.c = hs, x16, c
        ctx.params = params
        ctx.need_x = x.requires_grad
        return out

    @staticmethod
    def backward(ctx, g_out):
        spec, nl = ctx.spec, ctx.nl
        params = ctx.params
        Ws = params[0:2 * nl:2]
        Wo = params[-2]
        hs, x16, c = ctx.hs, ctx.x16, ctx.c
        width = Ws[0].shape[0]
        kx, kxp = ctx.kx, x16.shape[1]
        dev = g_out.device
        g_out = g_out.contiguous().float()
        scale = _loss_scale(g_out)
        inv = 1.0 / scale
        g = (g_out * scale).contiguous()
        gWo = torch.zeros(Wo.shape[0], width, device=dev)
        ops.colsum_f16(hs[-1], gWo, g=g)
        gbo = g.sum(0)
        g_z = ops.head_dgrad(g, Wo.detach().contiguous(), width, mask=hs[-1])
        gWs, gbs = [None] * nl, [None] * nl
        g_c = torch.zeros_like(c)
        g_x = torch.zeros(x16.shape[0], kxp, device=dev) if ctx.need_x else None

        def input_part(W, gW, gb, x_cols, c_cols, want_colsum):
            nonlocal g_c, g_x
            xa, xb = x_cols
            xg = torch.zeros(width, kxp, device=dev)
            ops.wgrad_tma(g_z, x16, xg, colsum=gb if want_colsum else None)
            gW[:, xa:xb] = xg[:, :kx]
            if c_cols is not None:       # the constant rides in the bias: rank-1 weight gradient, and its own gradient
                gW[:, c_cols[0]:c_cols[1]] = torch.outer(gb, c)
                g_c += W[:, c_cols[0]:c_cols[1]].t() @ gb
            if g_x is not None:
                g_x += ops.gemm_tma(g_z, _pad8(W[:, xa:xb]), kxp, mode=1, out16=False, out32=True)[2]
        for i in range(nl - 1, -1, -1):
            W = Ws[i].detach()
            gW = torch.zeros_like(W, dtype=torch.float32)
            gb = torch.zeros(width, device=dev)
            if i == 0:
                input_part(W, gW, gb, spec["x_cols"], spec["c_cols"], True)
            elif i == spec["skip"]:
                ha, hb = spec["sh_cols"]
                hg = torch.zeros(width, width, device=dev)
                ops.wgrad_tma(g_z, hs[i - 1], hg, colsum=gb)
                gW[:, ha:hb] = hg
                input_part(W, gW, gb, spec["sx_cols"], spec["sc_cols"], False)
                g_z = ops.gemm_tma(g_z, _h(W[:, ha:hb]), width, mode=1, mask=hs[i - 1])[0]
            else:
                ops.wgrad_tma(g_z, hs[i - 1], gW, colsum=gb)
                g_z = ops.gemm_tma(g_z, _h(W), width, mode=1, mask=hs[i - 1])[0]
            gWs[i], gbs[i] = gW * inv, gb * inv
        out = [None if g_x is None else (g_x[:, :kx] * inv), g_c * inv, None]
        for i in range(nl):
            out += [gWs[i], gbs[i]]
        out += [gWo * inv, gbo * inv]
        ctx.hs = ctx.x16 = None
        return tuple(out)


def canonical_mlp_train(net, pe, emb):
    """CanonicalMLP (mlp_rgb_sigma.py:16-58) on pe [P, 63] with the state embedding ``emb`` [64]: raw [P, 4]."""
    m = net.cnl_mlp
    lins = m.linears()
    kpe, kin, w = pe.shape[1], m.input_ch, m.mlp_width
    spec = dict(x_cols=(0, kpe), c_cols=(kpe, kin), skip=m.skip_layer_indices()[0], sx_cols=(0, kpe), sc_cols=(kpe, kin),
                sh_cols=(kin, kin + w))
    params = []
    for lin in lins:
        params += [lin.weight, lin.bias]
    params += [m.output_linear[0].weight, m.output_linear[0].bias]
    return MlpFn.apply(pe, emb, spec, *params)


def non_rigid_mlp_train(mlp, pe, cond, xyz):
    """NonRigidMotionMLP (mlp_offset.py:16-70): xyz + offset(pe, cond)."""
    lins = mlp.linears()
    cc, kpe, w = mlp.condition_code_size, mlp.pos_embed_size, mlp.mlp_width
    spec = dict(x_cols=(cc, cc + kpe), c_cols=(0, cc), skip=mlp.skip_layer_indices()[0], sh_cols=(0, w), sx_cols=(w, w + kpe),
                sc_cols=None)
    params = []
    for lin in lins[:-1]:
        params += [lin.weight, lin.bias]
    params += [lins[-1].weight, lins[-1].bias]
    return xyz + MlpFn.apply(pe, cond.reshape(-1), spec, *params)


# ============================================================================ whole-step CUDA graph
class GraphedStep:
    """``fn()`` - zero the gradient buffers, forward, objective, ``backward()`` - captured ONCE in a CUDA graph and replayed.

    A training chunk is ~1400 small launches; at a few thousand rays per GPU the host cannot enqueue them as fast as the
    GPU retires them (23 ms of enqueue for ~8 ms of kernels at 1024 rays), which is what caps strong scaling.  A replay is
    one launch.  Requirements on ``fn`` (checked by CUDA itself: a violation aborts the capture with an error):
      * static shapes and no host reads: ``Network.static_shapes = True``, ``train_hosnerf_chunk(dense=True)``, ray batches
        written in place into the same device tensors (``copy_``), ``times`` / ``iter_val`` as host scalars (their value is
        frozen into the graph: re-capture when the Hann window or the state index changes);
      * gradients accumulate into persistent buffers (``dist.FlatGrads`` or ``zero_grad(set_to_none=False)``);
      * collectives and the optimiser step stay outside (``fn`` ends with ``backward()``).
    The first ``warmup`` calls run eagerly (lazy initialisation, weight-cache fills, allocator warm-up), the next one captures.
    Returns whatever ``fn`` returns; after capture these are the graph's static output tensors, overwritten by every replay."""

    def __init__(self, fn, warmup: int = 3):
        self.fn, self.warmup = fn, warmup
        self.calls, self.graph, self.out = 0, None, None
        self.launches_per_step = 0

    def __call__(self):
        from . import _lib
        if self.graph is not None:
            self.graph.replay()
            _lib.LAUNCHES += self.launches_per_step
            return self.out
        self.calls += 1
        if self.calls <= self.warmup:
            return self.fn()
        import gc
        gc.collect()               # autograd graphs of earlier eager steps still alive in reference cycles would hand their
        torch.cuda.synchronize()   # AccumulateGrad nodes (bound to the eager stream) to the captured backward
        g = torch.cuda.CUDAGraph()
        l0 = _lib.LAUNCHES
        # thread_local: only this thread is held to the capture rules (the NCCL watchdog thread polls its events meanwhile);
        # the autograd worker's launches still land in the graph because they go to the capturing stream
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self.out = self.fn()
        self.launches_per_step = _lib.LAUNCHES - l0
        _lib.LAUNCHES = l0
        self.graph = g
        return self()
