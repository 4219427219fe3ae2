"""Oracle (CPU, fp32 torch) for the human-object branch: stratified samples,
LBS motion field, Hann-windowed / Fourier encodings, non-rigid and canonical
MLPs, NeRF composite and the stage-3 depth-merge composite.
TEST INFRASTRUCTURE ONLY.

Restates S3/core/nets/human_nerf/network.py (+ the component files it loads) and
S3/src/model/mipnerf360/model.py:73-99, 1524-1596.  Weights are passed as a
``state_dict`` mapping with the reference key names.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

PARENT = {1: 0, 2: 0, 3: 0, 4: 1, 5: 2, 6: 3, 7: 4, 8: 5, 9: 6, 10: 7, 11: 8, 12: 9, 13: 9,
          14: 9, 15: 12, 16: 13, 17: 14, 18: 16, 19: 17, 20: 18, 21: 19, 22: 20, 23: 21,
          24: 23, 25: 22}


# --------------------------------------------------------------------------
# a14  samples along the ray (network.py:401-424)
# --------------------------------------------------------------------------
def z_samples(near, far, n_samples, rand=None):
    """near/far [n,1] -> z [n,S]; ``rand`` [n,S] switches on stratified jitter
    (network.py:416-424, the reference draws it with torch.rand)."""
    t = torch.linspace(0.0, 1.0, steps=n_samples).to(near)
    z = (near * (1.0 - t) + far * t).expand([near.shape[0], n_samples])
    if rand is not None:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * rand
    return z


# --------------------------------------------------------------------------
# motion bases (S3/core/utils/network_util.py:106-174) - per-frame, host side
# --------------------------------------------------------------------------
def motion_bases(dst_Rs, dst_Ts, cnl_gtfms):
    """[26,3,3],[26,3],[26,4,4] -> backward (R,T) and forward (R,T) bone maps.
    Shapes (leading batch of 1, batched matmul / inverse) are kept exactly as in the
    reference because the 4x4 inverse amplifies last-bit differences of the chain."""
    dst_Rs, dst_Ts, cnl = dst_Rs[None], dst_Ts[None], cnl_gtfms[None]
    nb = dst_Rs.shape[1]
    glob = torch.zeros_like(cnl)
    local = torch.zeros(size=(1, nb, 4, 4), dtype=dst_Rs.dtype)
    local[:, :, :3, :3] = dst_Rs
    local[:, :, :3, 3] = dst_Ts
    local[:, :, 3, 3] = 1.0
    glob[:, 0, :, :] = local[:, 0, :, :]
    for i in range(1, nb):
        glob[:, i, :, :] = torch.matmul(glob[:, PARENT[i], :, :].clone(), local[:, i, :, :])
    glob = glob.view(-1, 4, 4)
    cnl = cnl.view(-1, 4, 4)
    back = torch.matmul(cnl, torch.inverse(glob)).view(-1, nb, 4, 4)
    fwd = torch.matmul(glob, torch.inverse(cnl)).view(-1, nb, 4, 4)
    return back[0, :, :3, :3], back[0, :, :3, 3], fwd[0, :, :3, :3], fwd[0, :, :3, 3]


def rodrigues(rvec):
    """network_util.py:66-92."""
    theta = torch.sqrt(1e-5 + torch.sum(rvec**2, dim=1))
    r = rvec / theta[:, None]
    c, s = torch.cos(theta), torch.sin(theta)
    return torch.stack((
        r[:, 0]**2 + (1. - r[:, 0]**2) * c,
        r[:, 0] * r[:, 1] * (1. - c) - r[:, 2] * s,
        r[:, 0] * r[:, 2] * (1. - c) + r[:, 1] * s,
        r[:, 0] * r[:, 1] * (1. - c) + r[:, 2] * s,
        r[:, 1]**2 + (1. - r[:, 1]**2) * c,
        r[:, 1] * r[:, 2] * (1. - c) - r[:, 0] * s,
        r[:, 0] * r[:, 2] * (1. - c) - r[:, 1] * s,
        r[:, 1] * r[:, 2] * (1. - c) + r[:, 0] * s,
        r[:, 2]**2 + (1. - r[:, 2]**2) * c), dim=1).view(-1, 3, 3)


def _seq(sd, prefix, x, idxs, last_linear=True):
    """Apply Linear(+ReLU) blocks stored at ``prefix{idx}``; ReLU after all but the
    last when ``last_linear``."""
    for k, i in enumerate(idxs):
        x = F.linear(x, sd[f"{prefix}{i}.weight"], sd[f"{prefix}{i}.bias"])
        if not (last_linear and k == len(idxs) - 1):
            x = F.relu(x)
    return x


def pose_refine(sd, posevec, dst_Rs, dst_Ts, mlp_depth=4):
    """mlp_delta_body_pose.py:14-73 + network.py:590-605."""
    h = _seq(sd, "pose_decoder.block_mlps.", posevec[None], [0, 2, 4][: mlp_depth - 1], last_linear=False)
    rvec = _seq(sd, "pose_decoder.block_mlps_dstR.", h, [0, 2]).view(-1, 3)
    dR = rodrigues(rvec)
    dT = _seq(sd, "pose_decoder.block_mlps_dstT.", h, [0, 2]).view(-1, 3)
    Rs = torch.cat([dst_Rs[:1], torch.matmul(dst_Rs[1:], dR)], 0)
    Ts = torch.cat([dst_Ts[:1], dst_Ts[1:] + dT], 0)
    return Rs, Ts


def motion_weight_volume(sd, priors):
    """deconv_vol_decoder.py:34-42 + network_util.py:21-59: const embedding ->
    Linear+LeakyReLU -> 4x (ConvTranspose3d+LeakyReLU) -> ConvTranspose3d;
    softmax(logits + log prior) over the 27 channels.  -> [27,32,32,32]."""
    p = "mweight_vol_decoder."
    h = F.leaky_relu(F.linear(sd[p + "const_embedding"][None], sd[p + "decoder.block_mlp.0.weight"],
                              sd[p + "decoder.block_mlp.0.bias"]), 0.2).view(-1, 1024, 1, 1, 1)
    idxs = sorted({int(k.split(".")[3]) for k in sd if k.startswith(p + "decoder.block_conv.")})
    for j, i in enumerate(idxs):
        h = F.conv_transpose3d(h, sd[f"{p}decoder.block_conv.{i}.weight"],
                               sd[f"{p}decoder.block_conv.{i}.bias"], stride=2, padding=1)
        if j < len(idxs) - 1:
            h = F.leaky_relu(h, 0.2)
    return F.softmax(h + torch.log(priors[None]), dim=1)[0]


# --------------------------------------------------------------------------
# a15  LBS motion field (network.py:304-354)
# --------------------------------------------------------------------------
def lbs_warp(pts, Rs, Ts, vol, bbox_min, bbox_scale):
    """pts [P,3]; Rs [B,3,3]; Ts [B,3]; vol [B+1,G,G,G] -> x_skel [P,3], mask [P,1]."""
    w_list, pos_list = [], []
    for i in range(vol.shape[0] - 1):
        pos = torch.matmul(Rs[i], pts.T).T + Ts[i]
        g = (pos - bbox_min[None, :]) * bbox_scale[None, :] - 1.0
        w = F.grid_sample(vol[None, i:i + 1], g[None, None, None, :, :],
                          padding_mode="zeros", align_corners=True)
        w_list.append(w[0, 0, 0, 0, :, None])
        pos_list.append(pos)
    w = torch.cat(w_list, dim=-1)
    wsum = torch.sum(w, dim=-1, keepdim=True)
    x = torch.sum(torch.stack([w[:, i:i + 1] * pos_list[i] for i in range(w.shape[-1])], 0), 0)
    return x / wsum.clamp(min=0.0001), wsum


def lbs_forward(cnl_pts, Rs_f, Ts_f, vol, bbox_min, bbox_scale):
    """Forward warp used by the cycle/flow side paths (network.py:357-398)."""
    g = (cnl_pts - bbox_min[None, :]) * bbox_scale[None, :] - 1.0
    w = F.grid_sample(vol[None, :-1], g[None, None, None, :, :], padding_mode="zeros",
                      align_corners=True)[0, :, 0, 0, :].permute(1, 0)
    wsum = torch.sum(w, dim=-1, keepdim=True)
    x = torch.sum(torch.stack([w[:, i:i + 1] * (torch.matmul(Rs_f[i], cnl_pts.T).T + Ts_f[i])
                               for i in range(w.shape[-1])], dim=0), dim=0)
    return x / wsum.clamp(min=0.0001), wsum


# --------------------------------------------------------------------------
# a16/a18  encodings (embedders/hannw_fourier.py:15-71, fourier.py:13-57)
# --------------------------------------------------------------------------
def hann_weights(n_freqs, iter_val, kick_in, full_band):
    """w_k = (1 - cos(pi * clamp(alpha - k, 0, 1))) / 2, alpha = m * max(it-kick,0) / (full-kick)."""
    kick = torch.tensor(kick_in, dtype=torch.float32)
    t = torch.clamp(iter_val - kick, min=0.)
    alpha = n_freqs * t / (full_band - kick)
    return [(1. - torch.cos(np.pi * torch.clamp(alpha - k, min=0., max=1.))) / 2. for k in range(n_freqs)]


def hann_embed(x, n_freqs, iter_val, kick_in, full_band):
    freqs = 2. ** torch.linspace(0., n_freqs - 1, steps=n_freqs)
    ws = hann_weights(n_freqs, iter_val, kick_in, full_band)
    out = []
    for f, w in zip(freqs, ws):
        out += [w * torch.sin(x * f), w * torch.cos(x * f)]
    return torch.cat(out, -1)


def fourier_embed(x, n_freqs):
    freqs = 2. ** torch.linspace(0., n_freqs - 1, steps=n_freqs)
    out = [x]
    for f in freqs:
        out += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(out, -1)


# --------------------------------------------------------------------------
# a17/a19  MLPs (mlp_offset.py:16-70, mlp_rgb_sigma.py:16-58)
# --------------------------------------------------------------------------
def non_rigid_mlp(sd, prefix, pos_embed, xyz, cond, depth=6, skips=(4,)):
    h = torch.cat([cond.expand(xyz.shape[0], cond.shape[-1]), pos_embed], dim=-1)
    for layer in range(depth):
        if layer in skips and layer > 0:
            h = torch.cat([h, pos_embed], dim=-1)
        h = F.relu(F.linear(h, sd[f"{prefix}block_mlps.{2 * layer}.weight"], sd[f"{prefix}block_mlps.{2 * layer}.bias"]))
    off = F.linear(h, sd[f"{prefix}block_mlps.{2 * depth}.weight"], sd[f"{prefix}block_mlps.{2 * depth}.bias"])
    return xyz + off


def canonical_mlp(sd, pos_embed, depth=8, skips=(4,)):
    h = pos_embed
    for layer in range(depth):
        if (layer - 1) in skips:                      # skip i -> concat before linear i+1
            h = torch.cat([pos_embed, h], dim=-1)
        h = F.relu(F.linear(h, sd[f"cnl_mlp.pts_linears.{2 * layer}.weight"], sd[f"cnl_mlp.pts_linears.{2 * layer}.bias"]))
    return F.linear(h, sd["cnl_mlp.output_linear.0.weight"], sd["cnl_mlp.output_linear.0.bias"])


# --------------------------------------------------------------------------
# a20  S2 NeRF composite (S2 network.py:273-299)
# --------------------------------------------------------------------------
def raw2outputs_s2(raw, mask, z_vals, rays_d, bgcolor):
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], dim=-1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    alpha = (1.0 - torch.exp(-F.relu(raw[..., 3]) * dists)) * mask[:, :, 0]
    T = torch.cumprod(torch.cat([torch.ones((alpha.shape[0], 1)), 1. - alpha + 1e-10], dim=-1), dim=-1)[:, :-1]
    w = alpha * T
    rgb_map = torch.sum(w[..., None] * rgb, -2)
    depth = torch.sum(w * z_vals, -1)
    acc = torch.sum(w, -1)
    return rgb_map + (1. - acc[..., None]) * bgcolor[None, :] / 255., acc, w, depth


# --------------------------------------------------------------------------
# a21  S3 composite (S3 model.py:73-99 and 1524-1596)
# --------------------------------------------------------------------------
def raw2outputs_s3(raw, z_vals, rays_d, pts_mask=None, bgcolor=None):
    """raw already activated: [..,:3] rgb in [0,1], [..,3] sigma >= 0."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], dim=-1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    alpha = 1.0 - torch.exp(-raw[..., 3] * dists)
    if pts_mask is not None:
        alpha = alpha * pts_mask[..., 0]
    T = torch.cumprod(torch.cat([torch.ones((alpha.shape[0], 1)), 1. - alpha + 1e-10], dim=-1), dim=-1)[:, :-1]
    w = alpha * T
    rgb_map = torch.sum(w[..., None] * raw[..., :3], -2)
    depth = torch.sum(w * z_vals, -1)
    acc = torch.sum(w, -1)
    if bgcolor is not None:
        rgb_map = rgb_map + (1. - acc[..., None]) * bgcolor[None, :] / 255.
    return rgb_map, acc, w, depth


def human_depth_on_bkg_ray(newsmpl_pts, M, rays_o, rays_d):
    """model.py:1524-1545 (non-degenerate branch): depth of the human samples along
    the background ray = mean_xyz((p_world - o) / (d + 1e-10))."""
    hom = torch.cat([newsmpl_pts, torch.ones_like(newsmpl_pts[..., 0:1])], -1)
    world = torch.einsum("ji,bni->bnj", M, hom)[..., :3]
    return torch.mean((world - rays_o[..., None, :]) / (rays_d[..., None, :] + 1e-10), dim=-1)


def composite_s3(bkg_rgb, bkg_density, bkg_tdist, human_rgb, human_density, pts_mask,
                 newsmpl_pts, M, rays_o, rays_d, thre_fg=5e-3):
    """Merge background samples (tdist[..., :-1]) and human samples by depth on
    foreground rays, composite; background-only rays composite the bkg samples
    alone.  Returns rgb [n,3], idx_fg [n], human_weights [n_fg,S_h]."""
    zh_all = human_depth_on_bkg_ray(newsmpl_pts, M, rays_o, rays_d)
    idx_fg = torch.sum(pts_mask, dim=-1) > thre_fg
    idx_bg = ~idx_fg
    rgb = torch.zeros(pts_mask.shape[0], 3)
    zb = bkg_tdist[..., :-1]
    bkg = torch.cat([bkg_rgb, bkg_density[..., None]], -1)
    hum = torch.cat([human_rgb, human_density[..., None]], -1)
    z_all = torch.cat([zb[idx_fg], zh_all[idx_fg]], -1)
    z_sorted, order = torch.sort(z_all, -1)
    both = torch.cat([bkg[idx_fg], hum[idx_fg]], 1)
    both = torch.gather(both, 1, order[..., None].expand(-1, -1, 4))
    m = torch.cat([torch.ones_like(zb[idx_fg]), pts_mask[idx_fg]], -1)
    m = torch.gather(m, 1, order)[..., None]
    rgb_fg, _, w_fg, _ = raw2outputs_s3(both, z_sorted, rays_d[idx_fg], m)
    is_h = order >= zb.shape[1]
    human_w = w_fg[is_h].reshape(-1, human_rgb.shape[1])
    rgb[idx_fg] = rgb_fg
    mb = torch.ones_like(zb[idx_bg])[..., None]
    rgb_bg, _, _, _ = raw2outputs_s3(bkg[idx_bg], zb[idx_bg], rays_d[idx_bg], mb)
    rgb[idx_bg] = rgb_bg
    return rgb, idx_fg, human_w, order


# --------------------------------------------------------------------------
# a22  Network.forward, eval path and the train-mode outputs that share kernels
# --------------------------------------------------------------------------
def network_forward(sd, batch, *, n_samples=128, kick_in_iter=100000, full_band_iter=200000,
                    pose_kick_in_iter=20000, rand=None, stage2=False, transitions_times=None):
    """Restates Network.forward/_render_rays (S3 network.py:427-698; S2 returns the
    composited rgb/alpha/depth/weights instead, S2 network.py:537-556).  With ``is_train`` and
    ``time > 0.005`` the flow side path (network.py:474-502, 609-631) adds ``deform_pts_prev_final``
    and the S3 dict drops z_vals / rays_d (network.py:538-547)."""
    rays_o, rays_d = batch["rays"][0].reshape(-1, 3).float(), batch["rays"][1].reshape(-1, 3).float()
    iter_val = batch["iter_val"]
    dst_Rs, dst_Ts = batch["dst_Rs"], batch["dst_Ts"]
    if iter_val >= pose_kick_in_iter:
        dst_Rs, dst_Ts = pose_refine(sd, batch["dst_posevec"], dst_Rs, dst_Ts)
    cond = batch["dst_posevec"][None]
    if iter_val < kick_in_iter:
        cond = torch.zeros_like(cond) * cond
    Rb, Tb, Rf, Tf = motion_bases(dst_Rs, dst_Ts, batch["cnl_gtfms"])
    vol = motion_weight_volume(sd, batch["motion_weights_priors"])
    z = z_samples(batch["near"], batch["far"], n_samples, rand)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    flat = pts.reshape(-1, 3)
    x_skel, mask = lbs_warp(flat, Rb, Tb, vol, batch["cnl_bbox_min_xyz"], batch["cnl_bbox_scale_xyz"])
    pe = hann_embed(x_skel, 6, iter_val, kick_in_iter, full_band_iter)
    cnl = non_rigid_mlp(sd, "non_rigid_mlp.", pe, x_skel, cond)
    n_emb = sum(1 for k in sd if k.startswith("human_stateembeds."))
    from .mip360_ref import select_state
    emb = sd[f"human_stateembeds.{select_state(n_emb, batch['time'], transitions_times)}"]
    xin = torch.cat([fourier_embed(cnl, 10), emb.repeat(cnl.shape[0], 1)], dim=-1)
    raw = canonical_mlp(sd, xin).reshape(pts.shape[0], n_samples, 4)
    mask = mask.reshape(pts.shape[0], n_samples, 1)
    # cycle side path (always on, network.py:505-536)
    sel = (mask > 0.005)[..., 0].reshape(-1)
    if sel.sum() > 0:
        observe = flat[sel]
        xd, _ = lbs_forward(cnl[sel], Rf, Tf, vol, batch["cnl_bbox_min_xyz"], batch["cnl_bbox_scale_xyz"])
        pe_f = hann_embed(xd, 6, iter_val, kick_in_iter, full_band_iter)
        deform = non_rigid_mlp(sd, "non_rigid_forward_mlp.", pe_f, xd, cond)
    else:
        observe = deform = pts[0, 0, :][None, :]
    out = {"deform_pts_final": deform, "observe_pts": observe, "bgcolor": batch["bgcolor"]}
    # flow side path: ALL canonical points warped forward with the PREVIOUS frame's pose (network.py:474-502)
    flow = bool(batch.get("is_train", False)) and float(batch["time"]) > 0.005
    if flow:
        Rs_p, Ts_p = batch["dst_Rs_prev"], batch["dst_Ts_prev"]
        if iter_val >= pose_kick_in_iter:
            Rs_p, Ts_p = pose_refine(sd, batch["dst_posevec_prev"], Rs_p, Ts_p)
        _, _, Rf_p, Tf_p = motion_bases(Rs_p, Ts_p, batch["cnl_gtfms"])
        cond_p = batch["dst_posevec_prev"][None]
        if iter_val < kick_in_iter:
            cond_p = torch.zeros_like(cond_p) * cond_p
        xd_p, _ = lbs_forward(cnl, Rf_p, Tf_p, vol, batch["cnl_bbox_min_xyz"], batch["cnl_bbox_scale_xyz"])
        pe_p = hann_embed(xd_p, 6, iter_val, kick_in_iter, full_band_iter)
        out["deform_pts_prev_final"] = non_rigid_mlp(sd, "non_rigid_forward_mlp.", pe_p, xd_p, cond_p).reshape(pts.shape)
    if stage2:
        rgb, acc, w, depth = raw2outputs_s2(raw, mask, z, rays_d, batch["bgcolor"])
        out.update(rgb=rgb, alpha=acc, depth=depth, weights=w)
    else:
        out.update(human_rgb=torch.sigmoid(raw[..., :3]), human_density=F.relu(raw[..., 3]),
                   newsmpl_pts=pts, pts_mask=mask[..., 0])
        if not flow:
            out.update(z_vals=z, rays_d=rays_d)
    out["_x_skel"] = x_skel.reshape(pts.shape)
    out["_cnl_pts"] = cnl.reshape(pts.shape)
    out["_raw"] = raw
    return out
