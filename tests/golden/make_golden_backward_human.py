"""Generate tests/golden/human_s2_backward.npz: gradients of the surrogate objective mean(rgb) + cycle term through the
UNMODIFIED stage-2 human-object network (2nd_State_Conditional_Human-Object/core/nets/human_nerf/network.py:574-698,
training mode, time = 0 so the flow side path is off) - BASELINE.json's C4 uses the same mean(rgb) surrogate.  Stored:
loss, which parameters receive gradient, per-parameter gradient norm and leading entries (the backward contract of the
human branch for the next round).  Authoring container only:   python tests/golden/make_golden_backward_human.py
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
out = {}
with rh.stage(rh.S2):
    import core.nets.human_nerf.network as N
    cfg = rh.human_cfg(rh.S2)
    cfg.perturb = 0.0
    net = N.Network(cfg)
    synth.fill_params_(net, 0)
    with torch.no_grad():
        net.cnl_mlp.output_linear[0].bias[3] += 3.0          # see synth.boost_human_density_
    b = synth.make_human_batch(12)
    b["is_train"] = True
    res = net(**b)
    cyc = torch.mean(torch.sum((res["observe_pts"] - res["deform_pts_final"]) ** 2, 1) / 2.0)     # S3 model.py:1705-1707
    loss = res["rgb"].mean() + 0.1 * cyc
    loss.backward()
    out["loss"], out["cycle"] = np.float32(loss.item()), np.float32(cyc.item())
    out["rgb"] = res["rgb"].detach().numpy()
    out["n_cycle_pts"] = np.int64(res["observe_pts"].shape[0])
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
np.savez_compressed(os.path.join(HERE, "human_s2_backward.npz"), **out)
print(out["loss"], out["cycle"], out["n_cycle_pts"], len(names), sum(1 for k in out if k.startswith("gnone__")))
