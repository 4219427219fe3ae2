"""Stage-3 (complete HOSNeRF) per-chunk render: background branch + human-object branch + depth-merge
composite - the body that S3/src/model/mipnerf360/model.py repeats in training_step (:1501-1596) and in
the four eval loops (e.g. free_view :1324-1427), as one call."""
from __future__ import annotations

import torch

from . import ops


def render_hosnerf_chunk(bkg_model, human_net, batch_bkg: dict, batch_human: dict, newsmpl_to_scale_world,
                         near_bkg: float = 0.1, far_bkg: float = 1e6, train_frac: float = 1.0,
                         randomized: bool = False, thre_fg: float = 5e-3, rands=None, cycle_outputs: bool = False):
    """bkg_model: hosnerf_b200.MipNeRF360(stage3=True); human_net: hosnerf_b200.Network.
    batch_bkg: rays_o, rays_d, viewdirs, radii, times (scale-world frame); batch_human: the kwargs of
    Network.forward (rays in the new-SMPL frame, pose, bbox, ...).  Returns dict(rgb [n,3], idx_fg [n],
    human_weights [n,S_h], ray_history, net_output)."""
    with torch.no_grad():
        _, ray_history = bkg_model(batch_bkg, train_frac, randomized, False, near_bkg, far_bkg, rands=rands)
        # the cycle side path only feeds the training loss: render loops skip it (pass cycle_outputs=True to get it)
        net_output = human_net(**batch_human, cycle_outputs=cycle_outputs)
        h = ray_history[-1]
        n = h["density"].shape[0]
        s_h = net_output["human_density"].shape[-1]
        rgb, idx_fg, human_w = ops.composite_s3(
            h["rgb"].contiguous(), h["density"].contiguous(), h["tdist"].contiguous(),
            net_output["human_rgb"].reshape(n, s_h, 3).contiguous(),
            net_output["human_density"].reshape(n, s_h).contiguous(),
            net_output["pts_mask"].reshape(n, s_h).contiguous(),
            net_output["newsmpl_pts"].reshape(n, s_h, 3).contiguous(),
            newsmpl_to_scale_world, batch_bkg["rays_o"].contiguous().float(), batch_bkg["rays_d"].contiguous().float(),
            thre_fg=thre_fg)
    return {"rgb": rgb, "idx_fg": idx_fg, "human_weights": human_w, "ray_history": ray_history,
            "net_output": net_output}


def cycle_loss(net_output: dict):
    """The cycle-consistency term of the stage-2/3 objective (S3 src/model/mipnerf360/model.py:1705-1707):
    mean over the m foreground points of |observe - deform|^2 / 2, as a 0-d CUDA tensor (forward value; one
    deterministic reduction kernel)."""
    a, b = net_output["observe_pts"].contiguous().float(), net_output["deform_pts_final"].contiguous().float()
    if a.shape != b.shape:
        raise RuntimeError("hosnerf_b200.cycle_loss: observe_pts and deform_pts_final differ in shape")
    return ops.reduce_scaled(a, 0.5 / max(a.shape[0], 1), y=b)


def _composite_ray_set(samples, z, rays_d, mask, graph_safe: bool = False):
    """S3 ``_raw2outputs`` (model.py:73-99) on depth-ordered samples [m, S, 4] (rgb in [0,1], sigma >= 0): torch ops with a
    graph - the [rays, 192] tensors of a training chunk are a few hundred kilobytes.  ``graph_safe``: the transmittance as
    exp(cumsum(log)) - torch's cumprod backward reads ``(x == 0).any()`` back to the host, which a CUDA-graph capture cannot do
    (the factors are >= 1e-10, so the logarithm is finite)."""
    dists = torch.cat([z[..., 1:] - z[..., :-1], torch.full_like(z[..., :1], 1e10)], dim=-1) * torch.norm(rays_d[..., None, :], dim=-1)
    alpha = (1.0 - torch.exp(-samples[..., 3] * dists)) * mask
    if graph_safe:
        T = torch.exp(torch.cumsum(torch.log(1. - alpha + 1e-10), dim=-1) - torch.log(1. - alpha + 1e-10))
    else:
        T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1. - alpha + 1e-10], dim=-1), dim=-1)[:, :-1]
    w = alpha * T
    return torch.sum(w[..., None] * samples[..., :3], -2), w


def train_hosnerf_chunk(bkg_model, human_net, batch_bkg: dict, batch_human: dict, newsmpl_to_scale_world,
                        near_bkg: float = 0.1, far_bkg: float = 1e6, train_frac: float = 1.0, randomized: bool = True,
                        thre_fg: float = 5e-3, rands=None, dense: bool = False):
    """The differentiable body of the stage-3 ``training_step`` (S3 model.py:1501-1596): background branch (one
    ``train.RenderFn`` node: per-level weights and the final level's per-sample density / rgb carry the gradient), human-object
    branch (``Network.forward`` under autograd), depth merge + composite.  The merge itself (sort of 64 + 128 depths per
    foreground ray, gather, transmittance product) is a handful of torch ops on [rays, 192] tensors with autograd; the forward
    kernel ``hos_composite_s3`` is the eval path.  Returns dict(rgb [n,3], idx_fg, human_weights [n_fg, S_h], ray_history,
    net_output).

    ``dense=True`` is the static-shape variant for CUDA-graph capture (``train.GraphedStep``): both composites run on every
    ray and ``idx_fg`` selects between them, ``human_weights`` is [n, S_h] with zero rows on background rays, and nothing is
    read back to the host.  Same rgb and gradients as the indexed form."""
    _, ray_history = bkg_model(batch_bkg, train_frac, randomized, True, near_bkg, far_bkg, rands=rands)
    kw = dict(batch_human)
    kw["is_train"] = True
    net_output = human_net(**kw)
    h = ray_history[-1]
    n = h["density"].shape[0]
    s_h = net_output["human_density"].shape[-1]
    rays_o, rays_d = batch_bkg["rays_o"].float(), batch_bkg["rays_d"].float()
    M = newsmpl_to_scale_world.to(rays_o.device).float()
    pts = net_output["newsmpl_pts"].reshape(n, s_h, 3)
    world = torch.einsum("ji,bni->bnj", M, torch.cat([pts, torch.ones_like(pts[..., :1])], -1))[..., :3]
    if not dense and bool(torch.any(torch.abs(rays_d) < 1e-5)):
        raise NotImplementedError("hosnerf_b200.train_hosnerf_chunk: axis-aligned ray directions (the reference's per-axis branch, "
                                  "S3 model.py:1528-1543) are handled by the eval kernel only")
    zh = torch.mean((world - rays_o[:, None, :]) / (rays_d[:, None, :] + 1e-10), dim=-1)            # depth along the bkg ray
    mask = net_output["pts_mask"].reshape(n, s_h)
    idx_fg = torch.sum(mask, dim=-1) > thre_fg
    zb = h["tdist"][..., :-1]
    bkg = torch.cat([h["rgb"], h["density"][..., None]], -1)
    hum = torch.cat([net_output["human_rgb"].reshape(n, s_h, 3), net_output["human_density"].reshape(n, s_h, 1)], -1)
    if dense:
        nb = zb.shape[1]
        z_sorted, order = torch.sort(torch.cat([zb, zh.detach()], -1), -1)
        both = torch.gather(torch.cat([bkg, hum], 1), 1, order[..., None].expand(-1, -1, 4))
        m = torch.gather(torch.cat([torch.ones_like(zb), mask], -1), 1, order)
        rgb_fg, w_fg = _composite_ray_set(both, z_sorted, rays_d, m, graph_safe=True)
        rgb_bg, _ = _composite_ray_set(bkg, zb, rays_d, torch.ones_like(zb), graph_safe=True)
        rgb = torch.where(idx_fg[:, None], rgb_fg, rgb_bg)
        # the human samples of a ray are already depth-ordered, so un-sorting puts their weights in the reference's order
        w_orig = torch.zeros_like(w_fg).scatter(1, order, w_fg)
        human_w = torch.where(idx_fg[:, None], w_orig[:, nb:], torch.zeros_like(w_orig[:, nb:]))
        return {"rgb": rgb, "idx_fg": idx_fg, "human_weights": human_w, "ray_history": ray_history, "net_output": net_output}
    rgb = torch.zeros(n, 3, device=rays_o.device)
    human_w = torch.zeros(0, s_h, device=rays_o.device)
    if bool(idx_fg.any()):
        z_sorted, order = torch.sort(torch.cat([zb[idx_fg], zh[idx_fg].detach()], -1), -1)
        both = torch.gather(torch.cat([bkg[idx_fg], hum[idx_fg]], 1), 1, order[..., None].expand(-1, -1, 4))
        m = torch.gather(torch.cat([torch.ones_like(zb[idx_fg]), mask[idx_fg]], -1), 1, order)
        rgb_fg, w_fg = _composite_ray_set(both, z_sorted, rays_d[idx_fg], m)
        rgb = rgb.index_put((torch.nonzero(idx_fg)[:, 0],), rgb_fg)
        human_w = w_fg[order >= zb.shape[1]].reshape(-1, s_h)
    if bool((~idx_fg).any()):
        ib = ~idx_fg
        rgb_bg, _ = _composite_ray_set(bkg[ib], zb[ib], rays_d[ib], torch.ones_like(zb[ib]))
        rgb = rgb.index_put((torch.nonzero(ib)[:, 0],), rgb_bg)
    return {"rgb": rgb, "idx_fg": idx_fg, "human_weights": human_w, "ray_history": ray_history, "net_output": net_output}


def flow_loss(ray_grid, newsmpl_to_camera_prev, intrinsics_prev, human_weights, deform_pts_prev_final):
    """The flow term of the stage-3 objective (S3 src/model/mipnerf360/model.py:1680-1688 with img2mae :61-72): the
    previous-frame points (``deform_pts_prev_final`` of the foreground rays, [n_fg, S, 3]) are projected into the previous
    camera, and their weighted distance to the optical flow stored in ``ray_grid`` [n_fg, 5] = (x, y, flow_x, flow_y, valid)
    is averaged.  A dozen elementwise torch ops on [n_fg, S, 2] tensors with a graph: gradient reaches the forward
    non-rigid MLP, the previous frame's bone maps and the volume through ``deform_pts_prev_final`` (and the composite
    through ``human_weights``)."""
    hom = torch.cat([deform_pts_prev_final, torch.ones_like(deform_pts_prev_final[..., :1])], dim=-1)
    cam = torch.einsum("ji,bni->bnj", newsmpl_to_camera_prev, hom)[..., :3]
    p2 = torch.einsum("ji,bni->bnj", intrinsics_prev, cam)
    p2 = p2[..., :-1] / p2[..., -1:]
    grid = ray_grid.unsqueeze(1).repeat(1, p2.shape[1], 1)
    x, y, M = p2 - grid[..., :2], grid[..., 2:4], grid[..., -1].unsqueeze(-1)
    return torch.sum(torch.abs(x - y) * human_weights[..., None] * M) / (torch.sum(M) + 1e-8) / x.shape[-1]
