"""Generate tests/golden/s1_losses.npz by running the UNMODIFIED reference loss code
(S1 src/model/mipnerf360/helper.py:92-128 and LitMipNeRF360.interlevel_loss / distortion_loss /
the training_step objective, S1 src/model/mipnerf360/model.py:488-512, 609-625) on

  * the ray histories already stored in s1_forward_default.npz / s1_forward_default_rand.npz / s1_forward_c2.npz
    (outputs of the reference forward pass), and
  * synthetic histograms with the edge cases of the search (coinciding edges, intervals outside the
    envelope, zero weights, a single-bin envelope).

Authoring container only (needs /root/reference):   python tests/golden/make_golden_losses.py
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")

import ref_harness as rh  # noqa: E402

rh.install_stubs()
out = {}

with rh.stage(rh.S1):
    import src.model.mipnerf360.helper as H
    import src.model.mipnerf360.model as M

    Lit = M.LitMipNeRF360
    for name in ("s1_forward_default", "s1_forward_default_rand", "s1_forward_c2"):
        G = np.load(os.path.join(HERE, name + ".npz"))
        levels = sorted({int(k[1]) for k in G.files if k.startswith("L") and k[2] == "_"})
        hist = [{"sdist": torch.from_numpy(G[f"L{i}_sdist"]), "weights": torch.from_numpy(G[f"L{i}_weights"])} for i in levels]
        rgb = torch.from_numpy(G[f"R{levels[-1]}_rgb"])
        target = torch.rand(rgb.shape, generator=torch.Generator().manual_seed(5))
        inter = Lit.interlevel_loss(None, hist)
        dist = Lit.distortion_loss(None, hist)
        mse = H.img2mse(rgb, target)
        # the objective of training_step with the constructor defaults (data 1.0, interlevel 1.0, distortion 0.01, padding 1e-3)
        loss = torch.sqrt(mse + 0.001 ** 2) * 1.0 + inter * 1.0 + dist * 0.01
        out[f"{name}__target"] = target.numpy()
        out[f"{name}__interlevel"] = np.float32(inter)
        out[f"{name}__distortion"] = np.float32(dist)
        out[f"{name}__mse"] = np.float32(mse)
        out[f"{name}__psnr"] = np.float32(H.mse2psnr(mse))
        out[f"{name}__loss"] = np.float32(loss)
        c, w = hist[-1]["sdist"], hist[-1]["weights"]
        out[f"{name}__distortion_rays"] = H.lossfun_distortion(c, w).numpy()
        for i, lvl in enumerate(hist[:-1]):
            out[f"{name}__outer_L{i}"] = H.lossfun_outer(c, w, lvl["sdist"], lvl["weights"]).numpy()

    # synthetic edge cases ---------------------------------------------------------------------------
    g = torch.Generator().manual_seed(23)

    def hist_rand(n, s, lo=0.0, hi=1.0):
        t = torch.sort(torch.rand(n, s + 1, generator=g) * (hi - lo) + lo, dim=-1).values
        w = torch.rand(n, s, generator=g)
        return t, w / w.sum(-1, keepdim=True)

    n = 24
    t, w = hist_rand(n, 39)
    te, we = hist_rand(n, 19)
    te[:8] = t[:8, ::2][:, :20]                 # envelope edges coincide with fine edges (>= / < tie handling)
    t[8:12] = t[8:12] * 0.4 + 0.3               # fine histogram strictly inside the envelope range
    te[8:12, 0], te[8:12, -1] = 0.0, 1.0
    te[12:16] = te[12:16] * 0.3 + 0.35          # fine intervals partly outside the envelope range (clamped indices)
    w[16:18, 5:20] = 0.0                         # empty fine bins (0 / (0 + eps))
    we[18:20] = 0.0                              # empty envelope: loss = w^2 / (w + eps)
    t[20:22, 10:14] = t[20:22, 10:11]           # repeated fine edges (zero-width intervals)
    te[22:24, 3:7] = te[22:24, 3:4]             # repeated envelope edges
    out["syn_t"], out["syn_w"], out["syn_te"], out["syn_we"] = t.numpy(), w.numpy(), te.numpy(), we.numpy()
    out["syn_outer"] = H.lossfun_outer(t, w, te, we).numpy()
    out["syn_distortion"] = H.lossfun_distortion(t, w).numpy()
    te1, we1 = hist_rand(n, 1)                   # level-0 style single-bin envelope
    out["syn_te1"], out["syn_we1"] = te1.numpy(), we1.numpy()
    out["syn_outer1"] = H.lossfun_outer(t, w, te1, we1).numpy()
    # the widest histogram the kernels accept in one call at C2-like size
    t2, w2 = hist_rand(6, 128)
    te2, we2 = hist_rand(6, 383)
    out["big_t"], out["big_w"], out["big_te"], out["big_we"] = t2.numpy(), w2.numpy(), te2.numpy(), we2.numpy()
    out["big_outer"] = H.lossfun_outer(t2, w2, te2, we2).numpy()
    out["big_distortion"] = H.lossfun_distortion(t2, w2).numpy()

np.savez_compressed(os.path.join(HERE, "s1_losses.npz"), **out)
print({k: (v.shape, str(v.dtype)) for k, v in out.items()})
