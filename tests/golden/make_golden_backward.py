"""Generate tests/golden/s1_backward.npz: gradients of the stage-1 training objective computed by autograd through the
UNMODIFIED reference (S1 src/model/mipnerf360/model.py: MipNeRF360.forward :331-461 in training mode and the objective
of LitMipNeRF360.training_step :488-512 with interlevel_loss :609-618 and distortion_loss :620-625).

The fixture pins the backward contract for the kernels of the next round: which tensors receive gradient (sample
positions and contracted Gaussians are detached in the reference, so gradient reaches the MLP parameters only through
density / rgb -> composite weights -> losses), the loss value, and per parameter the gradient norm and its first entries.
Authoring container only (needs /root/reference):   python tests/golden/make_golden_backward.py
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
from hosnerf_b200 import synth  # noqa: E402

rh.install_stubs()
torch.set_num_threads(8)
N, SEED_RAND, TRAIN_FRAC = 8, 123, 0.4
out = {}
with rh.stage(rh.S1):
    import src.model.mipnerf360.helper as H
    import src.model.mipnerf360.model as M

    saved = M.NeRFMLP.__init__.__defaults__
    M.NeRFMLP.__init__.__defaults__ = (8, 256)            # gin: NeRFMLP.netwidth = 256 (the C2 network)
    try:
        net = M.MipNeRF360("/nonexistent", num_prop_samples=64, num_nerf_samples=32, num_levels=3, opaque_background=True)
    finally:
        M.NeRFMLP.__init__.__defaults__ = saved
    synth.fill_params_(net, 0)
    b = synth.make_bkg_batch(N, seed=2)
    target = torch.rand(N, 3, generator=torch.Generator().manual_seed(6))
    torch.manual_seed(SEED_RAND)
    rands = [torch.rand(N, 1) for _ in range(3)]
    torch.manual_seed(SEED_RAND)
    rendered, hist = net(b, TRAIN_FRAC, True, True, 0.1, 1e6)
    rgb = rendered[-1]["rgb"]
    mse = H.img2mse(rgb, target)
    Lit = M.LitMipNeRF360
    inter, dist = Lit.interlevel_loss(None, hist), Lit.distortion_loss(None, hist)
    loss = torch.sqrt(mse + 0.001 ** 2) * 1.0 + inter * 1.0 + dist * 0.01
    loss.backward()
    out["loss"], out["mse"], out["interlevel"], out["distortion"] = (np.float32(loss.item()), np.float32(mse.item()),
                                                                     np.float32(inter.item()), np.float32(dist.item()))
    out["target"] = target.numpy()
    for i, r in enumerate(rands):
        out[f"rand{i}"] = r.numpy()
    for k, v in b.items():
        out[f"in_{k}"] = v.numpy()
    names = []
    for name, p in net.named_parameters():
        names.append(name)
        if p.grad is None:
            out[f"gnone__{name}"] = np.bool_(True)
            continue
        g = p.grad.detach().reshape(-1)
        out[f"gnorm__{name}"] = np.float64(g.double().norm().item())
        out[f"ghead__{name}"] = g[:16].numpy().copy()
        out[f"gabsmax__{name}"] = np.float32(g.abs().max().item())
    out["param_names"] = np.array(names)
np.savez_compressed(os.path.join(HERE, "s1_backward.npz"), **out)
print({k: (getattr(v, "shape", ()), str(getattr(v, "dtype", type(v)))) for k, v in list(out.items())[:12]}, len(out))
