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
