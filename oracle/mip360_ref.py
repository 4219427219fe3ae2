"""Oracle (CPU, fp32 torch) for the background branch: proposal resampling,
conical-frustum Gaussians, contraction, integrated positional encoding, the
Prop/NeRF MLPs and the mip-360 alpha composite.  TEST INFRASTRUCTURE ONLY.

Restates S1/src/model/mipnerf360/helper.py and model.py (S3 copies differ only
where noted).  Network weights are passed as a plain ``state_dict``-style
mapping with the reference's key names (``mlps.{i}.pts_linear.{j}.weight`` ...).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1.1920929e-07  # fp32 machine epsilon, S1 helper.py:18


# --------------------------------------------------------------------------
# a1  ray warp (S1 helper.py:146-150)
# --------------------------------------------------------------------------
def s_to_t(s, near, far):
    s_near, s_far = 1 / near, 1 / far
    return 1 / (s * s_far + (1 - s) * s_near)


# --------------------------------------------------------------------------
# a2  max-dilate (S1 helper.py:130-143, 152-164) + caller trimming (model.py:373-382)
# --------------------------------------------------------------------------
def max_dilate_weights(t, w, dilation, lo, hi, renormalize=True):
    """t [N,S+1], w [N,S] -> (t' [N,3S+1], w' [N,3S]).  The caller drops the first
    and last element of each (model.py:381-382)."""
    width = torch.clip(t[..., 1:] - t[..., :-1], min=EPS)
    pdf = w / width                                               # helper.py:152-153
    left = t[..., :-1] - dilation
    right = t[..., 1:] + dilation
    knots = torch.sort(torch.cat([t, left, right], dim=-1), dim=-1).values
    knots = torch.clip(knots, lo, hi)
    inside = (left[..., None, :] <= knots[..., None]) & (right[..., None, :] > knots[..., None])
    pdf_max = torch.where(inside, pdf[..., None, :], torch.zeros_like(pdf[..., None, :]))
    pdf_max = pdf_max.max(dim=-1).values[..., :-1]
    w_out = pdf_max * (knots[..., 1:] - knots[..., :-1])          # helper.py:155-156
    if renormalize:
        w_out = w_out / torch.clip(w_out.sum(dim=-1, keepdim=True), min=EPS)
    return knots, w_out


def anneal_bias(train_frac, slope):
    """model.py:384-388."""
    if slope > 0:
        return (slope * train_frac) / ((slope - 1) * train_frac + 1)
    return 1.0


def resample_logits(sdist, weights, anneal, padding=0.0):
    """model.py:390-394: -inf on empty intervals."""
    return torch.where(sdist[..., 1:] > sdist[..., :-1],
                       anneal * torch.log(weights + padding),
                       torch.full_like(weights, -torch.inf))


# --------------------------------------------------------------------------
# a3  inverse-CDF interval sampling (S1 helper.py:166-196, 306-359)
# --------------------------------------------------------------------------
def cdf_from_logits(logits):
    """softmax -> cumsum (clamped at 1) padded with 0 / 1; helper.py:166-173,193-194."""
    w = F.softmax(logits, dim=-1)
    c = torch.cumsum(w[..., :-1], dim=-1).clip(max=1.0)
    shape = c.shape[:-1] + (1,)
    return torch.cat([torch.zeros(shape).type_as(c), c, torch.ones(shape).type_as(c)], dim=-1)


def interval_index(u, cw):
    """Integer output implied by sorted_interp: idx = #{j : cw[j] <= u} - 1
    (the last knot with cw <= u).  int64 [N,S]."""
    return (u[..., None, :] >= cw[..., :, None]).sum(dim=-2) - 1


def sorted_interp(x, xp, fp):
    """Piecewise-linear lookup, mask-max/min formulation; helper.py:175-190."""
    ge = x[..., None, :] >= xp[..., :, None]
    f_lo = torch.max(torch.where(ge, fp[..., None], fp[..., :1, None]), dim=-2).values
    f_hi = torch.min(torch.where(~ge, fp[..., None], fp[..., -1:, None]), dim=-2).values
    x_lo = torch.max(torch.where(ge, xp[..., None], xp[..., :1, None]), dim=-2).values
    x_hi = torch.min(torch.where(~ge, xp[..., None], xp[..., -1:, None]), dim=-2).values
    frac = torch.clip(torch.nan_to_num((x - x_lo) / (x_hi - x_lo), 0), 0, 1)
    return f_lo + frac * (f_hi - f_lo)


def sample_u(num_samples, shape_prefix, randomized, single_jitter, rand=None):
    """Quantiles at which the CDF is inverted (deterministic_center=True branch of
    helper.py:314-330, the only one sample_intervals uses).  ``rand`` replaces the
    in-place ``torch.rand`` draw so a test can feed the same numbers to the kernel."""
    if not randomized:
        pad = 1 / (2 * num_samples)
        u = torch.linspace(pad, 1 - pad - EPS, num_samples)
        return torch.broadcast_to(u, tuple(shape_prefix) + (num_samples,))
    u_max = EPS + (1 - EPS) / num_samples
    max_jitter = (1 - u_max) / (num_samples - 1) - EPS
    d = 1 if single_jitter else num_samples
    if rand is None:
        rand = torch.rand(tuple(shape_prefix) + (d,))
    return torch.linspace(0, 1 - u_max, num_samples) + rand * max_jitter


def sample_intervals(t, logits, num_samples, lo, hi, randomized=False, single_jitter=True,
                     rand=None, return_aux=False):
    """helper.py:336-359.  t [N,M+1], logits [N,M] -> [N,S+1]."""
    u = sample_u(num_samples, t.shape[:-1], randomized, single_jitter, rand).type_as(t)
    cw = cdf_from_logits(logits)
    centers = sorted_interp(u, cw, t)
    mid = (centers[..., 1:] + centers[..., :-1]) / 2
    first = torch.clip(2 * centers[..., :1] - mid[..., :1], min=lo)
    last = torch.clip(2 * centers[..., -1:] - mid[..., -1:], max=hi)
    out = torch.cat([first, mid, last], dim=-1)
    if return_aux:
        return out, {"u": u, "cw": cw, "centers": centers, "idx": interval_index(u, cw)}
    return out


# --------------------------------------------------------------------------
# a5  conical frustum -> Gaussian (S1 helper.py:242-302, diag=False path)
# --------------------------------------------------------------------------
def cast_rays(tdist, origins, directions, radii):
    t0, t1 = tdist[..., :-1], tdist[..., 1:]
    mu = (t0 + t1) / 2
    hw = (t1 - t0) / 2
    t_mean = mu + (2 * mu * hw**2) / (3 * mu**2 + hw**2).clip(min=EPS)
    denom = (3 * mu**2 + hw**2).clip(min=EPS)
    t_var = (hw**2) / 3 - (4 / 15) * hw**4 * (12 * mu**2 - hw**2) / denom**2
    r_var = (mu**2) / 4 + (5 / 12) * hw**2 - (4 / 15) * (hw**4) / denom
    r_var = r_var * radii**2
    d = directions
    mean = d[..., None, :] * t_mean[..., None]
    d_sq = torch.sum(d**2, dim=-1, keepdim=True).clip(min=1e-10)
    dd = d[..., :, None] * d[..., None, :]
    eye = torch.eye(3).type_as(d)
    null = eye - d[..., :, None] * (d / d_sq)[..., None, :]
    cov = t_var[..., None, None] * dd[..., None, :, :] + r_var[..., None, None] * null[..., None, :, :]
    return mean + origins[..., None, :], cov


# --------------------------------------------------------------------------
# a6  scene contraction + Jacobian push-forward (S1 helper.py:26-60)
# --------------------------------------------------------------------------
def _contract_pt(x):
    m = torch.sum(x**2, dim=-1, keepdim=True).clip(min=1e-32)
    return torch.where(m <= 1, x, ((2 * torch.sqrt(m) - 1) / m) * x)


def contract(mean, cov):
    """Same op sequence as the reference (autodiff Jacobian via vmap(jacrev)) so the
    oracle is pinned to it as tightly as CPU fp32 allows."""
    n, s, _ = mean.shape
    flat = mean.reshape(n * s, 3)
    with torch.inference_mode(False), torch.enable_grad():
        z = torch.func.vjp(_contract_pt, mean)[0]
        jac = torch.func.vmap(torch.func.jacrev(_contract_pt))(flat)
    c = torch.einsum("bij,bjk->bik", jac, cov.reshape(n * s, 3, 3))
    c = torch.einsum("bij,bkj->bik", c, jac)
    return z.reshape(n, s, 3).detach(), c.reshape(n, s, 3, 3).detach()


def contract_closed_form(mean, cov):
    """Analytic Jacobian (what the CUDA kernel evaluates): for r=|x|>1,
    J = a I + b x x^T with a=(2r-1)/r^2, b = 2(1-r)/r^4; identity otherwise.
    Used by tests to bound the autodiff-vs-closed-form gap."""
    m = torch.sum(mean**2, dim=-1, keepdim=True).clip(min=1e-32)
    r = torch.sqrt(m)
    a = (2 * r - 1) / m
    b = 2 * (1 - r) / (m * m)
    inside = m <= 1
    z = torch.where(inside, mean, a * mean)
    J = a[..., None] * torch.eye(3) + b[..., None] * mean[..., :, None] * mean[..., None, :]
    J = torch.where(inside[..., None], torch.eye(3).expand_as(J), J)
    return z, J @ cov @ J.transpose(-1, -2)


# --------------------------------------------------------------------------
# a7  integrated positional encoding (S1 helper.py:62-78, 89-90)
# --------------------------------------------------------------------------
def lift_and_diagonalize(mean, cov, basis):
    return mean @ basis, torch.sum(basis[None, None, ...] * (cov @ basis), dim=-2)


def integrated_pos_enc(mean, var, min_deg, max_deg):
    scales = 2 ** torch.arange(min_deg, max_deg).type_as(mean)
    shape = list(mean.shape[:-1]) + [-1]
    sm = torch.reshape(mean[..., None, :] * scales[:, None], shape)
    sv = torch.reshape(var[..., None, :] * scales[:, None] ** 2, shape)
    phase = torch.cat([sm, sm + 0.5 * np.pi], dim=-1)     # cosine half = sin(x + fl32(pi/2))
    return torch.exp(-0.5 * torch.cat([sv, sv], dim=-1)) * torch.sin(phase)


def ipe_features(tdist, rays_o, rays_d, radii, basis, min_deg=0, max_deg=12):
    """cast_rays -> contract -> lift -> IPE: [N,S+1] -> [N,S,2*(max-min)*21]."""
    mean, cov = cast_rays(tdist, rays_o, rays_d, radii)
    mean, cov = contract(mean, cov)
    lm, lv = lift_and_diagonalize(mean, cov, basis)
    return integrated_pos_enc(lm, lv, min_deg, max_deg)


# --------------------------------------------------------------------------
# a10 view-direction encoding (S1 helper.py:80-87)
# --------------------------------------------------------------------------
def pos_enc(x, min_deg, max_deg, append_identity=True):
    scales = 2 ** torch.arange(min_deg, max_deg).type_as(x)
    xb = torch.reshape(x[..., None, :] * scales[:, None], x.shape[:-1] + (-1,))
    feat = torch.sin(torch.cat([xb, xb + 0.5 * np.pi], dim=-1))
    return torch.cat([x, feat], dim=-1) if append_identity else feat


# --------------------------------------------------------------------------
# a8/a9  state embedding + MLP (S1 model.py:126-264)
# --------------------------------------------------------------------------
def select_state(n_embeds, time, transitions_times, eps=1e-5):
    """Index of the state embedding used at ``time`` (model.py:137-206).
    First boundary is strict ``<  t0 - eps``, later ones ``<= tk + eps``."""
    if n_embeds == 1:
        return 0
    t = float(time)
    if t < float(transitions_times[0]) - eps:
        return 0
    for k in range(1, n_embeds - 1):
        if t <= float(transitions_times[k]) + eps:
            return k
    return n_embeds - 1


def mlp_forward(sd, prefix, feats, viewdirs, state_idx=0, netdepth=8, skip_layer=4,
                disable_rgb=False, density_bias=-1.0, rgb_padding=0.001, deg_view=4):
    """feats [N,S,504] -> density [N,S], rgb [N,S,3]."""
    n, s, _ = feats.shape
    emb = sd[f"{prefix}bkgd_stateembeds.{state_idx}"]
    x = torch.cat([feats, emb.repeat(n, s, 1)], dim=-1)           # model.py:208-209
    inputs = x
    for i in range(netdepth):
        x = F.relu(F.linear(x, sd[f"{prefix}pts_linear.{i}.weight"], sd[f"{prefix}pts_linear.{i}.bias"]))
        if i % skip_layer == 0 and i > 0:
            x = torch.cat([x, inputs], dim=-1)                    # model.py:215-216
    raw = F.linear(x, sd[f"{prefix}density_layer.weight"], sd[f"{prefix}density_layer.bias"])[..., 0]
    density = F.softplus(raw + density_bias)
    if disable_rgb:
        return density, torch.zeros(n, s, 3)
    b = F.linear(x, sd[f"{prefix}bottleneck_layer.weight"], sd[f"{prefix}bottleneck_layer.bias"])
    de = pos_enc(viewdirs, 0, deg_view, True)
    de = torch.broadcast_to(de[..., None, :], b.shape[:-1] + (de.shape[-1],))
    x = torch.cat([b, de], dim=-1)
    x = F.relu(F.linear(x, sd[f"{prefix}views_linear.0.weight"], sd[f"{prefix}views_linear.0.bias"]))
    x = F.linear(x, sd[f"{prefix}rgb_layer.weight"], sd[f"{prefix}rgb_layer.bias"])
    rgb = torch.sigmoid(x) * (1 + 2 * rgb_padding) - rgb_padding
    return density, rgb


# --------------------------------------------------------------------------
# a11/a12  alpha composite (S1 helper.py:198-238)
# --------------------------------------------------------------------------
def alpha_weights(density, tdist, dirs, opaque_background=False):
    delta = (tdist[..., 1:] - tdist[..., :-1]) * torch.norm(dirs[..., None, :], dim=-1)
    dd = density * delta
    if opaque_background:
        dd = torch.cat([dd[..., :-1], torch.full_like(dd[..., -1:], 1e10)], dim=-1)
    alpha = 1 - torch.exp(-dd)
    trans = torch.exp(-torch.cat([torch.zeros_like(dd[..., :1]),
                                  torch.cumsum(dd[..., :-1], dim=-1)], dim=-1))
    return alpha * trans, alpha, trans


def render_rgb(rgbs, weights, bg=1.0):
    acc = weights.sum(dim=-1)
    return (weights[..., None] * rgbs).sum(dim=-2) + torch.clip(1 - acc[..., None], min=0) * bg


# --------------------------------------------------------------------------
# a4/a13  level loop (S1 model.py:331-461; S3 :418-540 adds "tdist", drops renderings)
# --------------------------------------------------------------------------
def mip360_forward(sd, batch, train_frac, randomized, near, far, *, num_prop_samples=64,
                   num_nerf_samples=32, num_levels=3, netdepths=None, opaque_background=True,
                   anneal_slope=10, dilation_multiplier=0.5, dilation_bias=0.0025,
                   single_jitter=True, bg=1.0, transitions_times=None, rands=None,
                   stage3=False):
    """Returns (renderings, ray_history) with the reference's dict keys.  ``rands``
    is an optional list (one per level) of the uniform draws for randomized mode."""
    rays_o, rays_d, viewdirs, radii = batch["rays_o"], batch["rays_d"], batch["viewdirs"], batch["radii"]
    n = rays_o.shape[0]
    time = batch["times"] if stage3 else batch["times"][0:1]
    if netdepths is None:
        netdepths = [4] * (num_levels - 1) + [8]
    s_lo, s_hi = 0.0, 1.0
    sdist = torch.cat([torch.full((n, 1), s_lo), torch.full((n, 1), s_hi)], dim=-1)
    weights = torch.ones(n, 1)
    prod = 1
    history, renderings = [], []
    for lvl in range(num_levels):
        is_prop = lvl < num_levels - 1
        ns = num_prop_samples if is_prop else num_nerf_samples
        dilation = dilation_bias + dilation_multiplier * (s_hi - s_lo) / prod
        prod *= ns
        if lvl > 0 and (dilation_bias > 0 or dilation_multiplier > 0):
            sdist, weights = max_dilate_weights(sdist, weights, dilation, s_lo, s_hi, True)
            sdist, weights = sdist[..., 1:-1], weights[..., 1:-1]
        logits = resample_logits(sdist, weights, anneal_bias(train_frac, anneal_slope))
        sdist = sample_intervals(sdist, logits, ns, s_lo, s_hi, randomized, single_jitter,
                                 None if rands is None else rands[lvl]).detach()
        tdist = s_to_t(sdist, near, far)
        feats = ipe_features(tdist, rays_o, rays_d, radii, sd[f"mlps.{lvl}.pos_basis_t"])
        n_emb = sum(1 for k in sd if k.startswith(f"mlps.{lvl}.bkgd_stateembeds."))
        st = select_state(n_emb, time, transitions_times)
        density, rgb = mlp_forward(sd, f"mlps.{lvl}.", feats, viewdirs, st,
                                   netdepth=netdepths[lvl], disable_rgb=is_prop)
        weights = alpha_weights(density, tdist, rays_d, opaque_background)[0]
        res = {"density": density, "rgb": rgb, "sdist": sdist, "weights": weights}
        if stage3:
            res["tdist"] = tdist
        else:
            renderings.append({"rgb": render_rgb(rgb, weights, bg)})
        history.append(res)
    return renderings, history


# --------------------------------------------------------------------------
# geodesic basis (S1 helper.py:420-494) - restated for the oracle; the product
# has its own generator in hosnerf_b200/geopoly.py and both are checked against
# the golden copy of the reference buffer.
# --------------------------------------------------------------------------
def icosahedron_basis(subdiv=2, tol=1e-4):
    a = (math.sqrt(5) + 1) / 2
    v = np.array([(-1, 0, a), (1, 0, a), (-1, 0, -a), (1, 0, -a), (0, a, 1), (0, a, -1),
                  (0, -a, 1), (0, -a, -1), (a, 1, 0), (-a, 1, 0), (a, -1, 0), (-a, -1, 0)]) / math.sqrt(a + 2)
    faces = [(0, 4, 1), (0, 9, 4), (9, 5, 4), (4, 5, 8), (4, 8, 1), (8, 10, 1), (8, 3, 10), (5, 3, 8),
             (5, 2, 3), (2, 7, 3), (7, 10, 3), (7, 6, 10), (7, 11, 6), (11, 0, 6), (0, 1, 6), (6, 1, 10),
             (9, 0, 11), (9, 11, 2), (9, 2, 5), (7, 2, 11)]
    bary = np.array([(i, j, subdiv - i - j) for i in range(subdiv + 1) for j in range(subdiv + 1 - i)]) / subdiv
    pts = []
    for f in faces:
        p = bary @ v[list(f)]
        pts.append(p / np.linalg.norm(p, axis=1, keepdims=True))
    pts = np.concatenate(pts, 0)
    d2 = np.maximum(0, (pts**2).sum(1)[:, None] + (pts**2).sum(1)[None] - 2 * pts @ pts.T)
    first = np.array([np.min(np.argwhere(row <= tol)) for row in d2])
    pts = pts[np.unique(first)]
    d2n = np.maximum(0, (pts**2).sum(1)[:, None] + (pts**2).sum(1)[None] + 2 * pts @ pts.T)
    keep = np.any(np.triu(d2n < tol), 1)
    pts = pts[keep]
    return torch.from_numpy(pts[:, ::-1].copy().T).to(torch.float32)
