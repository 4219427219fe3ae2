"""ORACLE (test infrastructure, not a product path): numpy restatement of the reference's stage-1 loss terms,
SURVEY 8f item 1.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.  Pinned to
outputs of the unmodified reference by tests/golden/s1_losses.npz (generator: tests/golden/make_golden_losses.py).

  lossfun_outer       <-  S1 src/model/mipnerf360/helper.py:92-120  (searchsorted, inner_outer, lossfun_outer)
  lossfun_distortion  <-  S1 src/model/mipnerf360/helper.py:122-128
  stage1_objective    <-  S1 src/model/mipnerf360/model.py:488-512, 609-625

The reference finds the bracketing envelope edges with a dense [.., Se+1, S+1] comparison mask; edges ascend, so
the same indices come from counting the envelope edges <= v (np.searchsorted side="right").
"""
import numpy as np

EPS = np.float32(1.1920929e-07)          # helper.py:18


def _bracket(edges, v):
    """helper.searchsorted (helper.py:92-97) for ascending ``edges`` [E] and queries ``v`` [Q]:
    lo = last edge <= v (0 if none), hi = first edge > v (E - 1 if none)."""
    cnt = np.searchsorted(edges, v, side="right")
    return np.maximum(cnt - 1, 0), np.minimum(cnt, edges.shape[0] - 1)


def lossfun_outer(t, w, t_env, w_env):
    """helper.py:100-120: loss [N,S] = clip(w - w_outer, 0)^2 / (w + eps), with w_outer the envelope mass of every
    envelope bin that overlaps the fine interval (prefix sum at hi(t_{i+1}) minus prefix sum at lo(t_i))."""
    t, w, t_env, w_env = (np.asarray(a, np.float32) for a in (t, w, t_env, w_env))
    out = np.empty_like(w)
    for r in range(w.shape[0]):
        # torch's CPU cumsum accumulates float in double and rounds every partial sum to float
        cy = np.concatenate([[0.0], np.cumsum(w_env[r].astype(np.float64))]).astype(np.float32)
        lo, hi = _bracket(t_env[r], t[r])
        w_outer = cy[hi[1:]] - cy[lo[:-1]]
        d = np.maximum(w[r] - w_outer, np.float32(0))
        out[r] = d * d / (w[r] + EPS)
    return out


def lossfun_distortion(t, w):
    """helper.py:122-128: per ray, sum_ij w_i w_j |u_i - u_j| + sum_i w_i^2 (t_{i+1} - t_i) / 3, u = bin centres."""
    t, w = np.asarray(t, np.float32), np.asarray(w, np.float32)
    u = (t[:, 1:] + t[:, :-1]) / np.float32(2)
    pair = np.abs(u[:, :, None] - u[:, None, :])
    inter = (w * (w[:, None, :] * pair).sum(-1, dtype=np.float32)).sum(-1, dtype=np.float32)
    intra = (w * w * (t[:, 1:] - t[:, :-1])).sum(-1, dtype=np.float32) / np.float32(3)
    return inter + intra


def stage1_objective(history, rgb, target, data_loss_mult=1.0, interlevel_loss_mult=1.0, distortion_loss_mult=0.01,
                     charb_padding=0.001):
    """model.py:488-512 with interlevel_loss (609-618) and distortion_loss (620-625).  ``history`` is a list of
    dicts with "sdist" and "weights" (numpy).  Returns dict(loss, rgbloss, interlevel, distortion, psnr)."""
    c, w = history[-1]["sdist"], history[-1]["weights"]
    inter = np.float32(0)
    for lvl in history[:-1]:
        inter = inter + lossfun_outer(c, w, lvl["sdist"], lvl["weights"]).mean(dtype=np.float64).astype(np.float32)
    dist = lossfun_distortion(c, w).mean(dtype=np.float64).astype(np.float32)
    diff = np.asarray(rgb, np.float32) - np.asarray(target, np.float32)
    mse = (diff * diff).mean(dtype=np.float64).astype(np.float32)
    loss = np.sqrt(mse + np.float32(charb_padding) ** 2) * np.float32(data_loss_mult) \
        + inter * np.float32(interlevel_loss_mult) + dist * np.float32(distortion_loss_mult)
    psnr = np.float32(-10.0) * np.log(mse) / np.float32(np.log(10.0))
    return {"loss": np.float32(loss), "rgbloss": mse, "interlevel": inter, "distortion": dist, "psnr": np.float32(psnr)}


# ---------------------------------------------------------------------------------------------------------------
# Differentiable (torch) forms of the same terms: the oracle of the BACKWARD contract.  Gradients flow into `w`
# (and `w_env`); the edges are treated as constants, as in the reference, where sample positions are detached
# (S1 model.py:405-406) and interlevel_loss detaches the fine histogram (:613-614).
def lossfun_outer_t(t, w, t_env, w_env):
    import torch
    cy = torch.cat([torch.zeros_like(w_env[..., :1]), torch.cumsum(w_env, dim=-1)], dim=-1)
    cnt = torch.searchsorted(t_env.contiguous(), t.contiguous(), right=True)
    lo = torch.clamp(cnt - 1, min=0)
    hi = torch.clamp(cnt, max=t_env.shape[-1] - 1)
    w_outer = torch.gather(cy, -1, hi[..., 1:]) - torch.gather(cy, -1, lo[..., :-1])
    return torch.clip(w - w_outer, min=0) ** 2 / (w + float(EPS))


def lossfun_distortion_t(t, w):
    import torch
    u = (t[..., 1:] + t[..., :-1]) / 2
    pair = torch.abs(u[..., :, None] - u[..., None, :])
    inter = torch.sum(w * torch.sum(w[..., None, :] * pair, dim=-1), dim=-1)
    intra = torch.sum(w ** 2 * (t[..., 1:] - t[..., :-1]), dim=-1) / 3
    return inter + intra


def stage1_objective_t(history, rgb, target, data_loss_mult=1.0, interlevel_loss_mult=1.0, distortion_loss_mult=0.01,
                       charb_padding=0.001):
    """Differentiable objective of training_step (model.py:488-512); ``history`` holds torch tensors."""
    import torch
    c, w = history[-1]["sdist"], history[-1]["weights"]
    inter = 0.0
    for lvl in history[:-1]:
        inter = inter + torch.mean(lossfun_outer_t(c.detach(), w.detach(), lvl["sdist"], lvl["weights"]))
    dist = torch.mean(lossfun_distortion_t(c, w))
    mse = torch.mean((rgb - target) ** 2)
    loss = torch.sqrt(mse + charb_padding ** 2) * data_loss_mult + inter * interlevel_loss_mult + dist * distortion_loss_mult
    return {"loss": loss, "rgbloss": mse, "interlevel": inter, "distortion": dist}
