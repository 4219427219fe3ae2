"""Generate tests/golden/human_s3_flow.npz and human_s2_flow.npz: the train-mode return dict of the UNMODIFIED reference
``Network.forward`` with ``time > 0.005`` - the flow side path (previous-frame pose -> forward motion bases -> forward LBS of
ALL canonical points -> forward non-rigid MLP, S3 network.py:474-502 and 609-631) next to the cycle outputs.
Authoring container only (needs /root/reference):   python tests/golden/make_golden_flow.py
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
KEYS = ("human_rgb", "human_density", "newsmpl_pts", "pts_mask", "z_vals", "rays_d", "deform_pts_final", "observe_pts",
        "deform_pts_prev_final", "rgb", "alpha", "depth", "weights")
for tag, sdir in (("s3", rh.S3), ("s2", rh.S2)):
    with rh.stage(sdir):
        import core.nets.human_nerf.network as N
        cfg = rh.human_cfg(sdir)
        cfg.perturb = 0.0
        net = N.Network(cfg)
        synth.fill_params_(net, 0)
        with torch.no_grad():
            net.cnl_mlp.output_linear[0].bias[3] += 3.0      # see synth.boost_human_density_
        b = synth.make_human_batch(24, time=0.5, is_train=True)
        with torch.no_grad():
            out = net(**b)
        arrays = {k: out[k].numpy() for k in KEYS if k in out and out[k] is not None}
        arrays["out_keys"] = np.array(sorted(k for k in out if k != "bgcolor"))
        for k, v in b.items():
            if isinstance(v, torch.Tensor) and k != "motion_weights_priors":
                arrays[f"in_{k}"] = v.numpy()
        np.savez_compressed(os.path.join(HERE, f"human_{tag}_flow.npz"), **arrays)
        print(tag, {k: v.shape for k, v in arrays.items() if not k.startswith("in_")})

# ---- the flow term of the stage-3 objective on random inputs: LitMipNeRF360.flow_func (S3 model.py:1680-1688), unbound
with rh.stage(rh.S3):
    import src.model.mipnerf360.model as M3
    g = torch.Generator().manual_seed(17)
    n_fg, S = 37, 128
    ray_grid = torch.cat([torch.rand(n_fg, 2, generator=g) * 500, torch.randn(n_fg, 2, generator=g) * 3,
                          (torch.rand(n_fg, 1, generator=g) > 0.3).float()], dim=-1)
    cam = torch.eye(4)
    cam[:3, :3] = torch.linalg.qr(torch.randn(3, 3, generator=g))[0]
    cam[:3, 3] = torch.tensor([0.1, -0.2, 3.0])
    K = torch.tensor([[1500.0, 0.0, 250.0], [0.0, 1500.0, 250.0], [0.0, 0.0, 1.0]])
    w = torch.rand(n_fg, S, generator=g) * 0.05
    pts = torch.randn(n_fg, S, 3, generator=g) * 0.3
    val = M3.LitMipNeRF360.flow_func(None, ray_grid, cam, K, w, pts)
    np.savez_compressed(os.path.join(HERE, "flow_loss.npz"), ray_grid=ray_grid.numpy(), cam=cam.numpy(), K=K.numpy(), w=w.numpy(),
                        pts=pts.numpy(), loss=np.float32(val.item()))
    print("flow loss", float(val))
