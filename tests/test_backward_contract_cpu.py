"""The backward contract for the next round's kernels, pinned on CPU: autograd through the oracle (forward restatement +
differentiable loss terms) reproduces the gradients that autograd through the UNMODIFIED reference produced
(tests/golden/s1_backward.npz, generator tests/golden/make_golden_backward.py): loss value, which parameters receive
gradient, per-parameter gradient norm and leading entries."""
import os

import numpy as np
import torch

from hosnerf_b200 import MipNeRF360, synth
from oracle import losses_ref as L
from oracle import mip360_ref as R

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "s1_backward.npz"))


def test_oracle_autograd_matches_reference_gradients():
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    net = MipNeRF360("/nonexistent", num_prop_samples=64, num_nerf_samples=32, num_levels=3, nerf_netwidth=256,
                     opaque_background=True)
    synth.fill_params_(net, 0)
    params = {k: v.detach().clone().requires_grad_(v.is_floating_point() and k in dict(net.named_parameters()))
              for k, v in net.state_dict().items()}
    batch = {k[3:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("in_")}
    rands = [torch.from_numpy(G[f"rand{i}"]) for i in range(3)]
    rend, hist = R.mip360_forward(params, batch, 0.4, True, 0.1, 1e6, num_prop_samples=64, num_nerf_samples=32,
                                  num_levels=3, rands=rands)
    terms = L.stage1_objective_t(hist, rend[-1]["rgb"], torch.from_numpy(G["target"]))
    assert np.allclose(float(terms["loss"].detach()), G["loss"], rtol=2e-5)
    assert np.allclose(float(terms["rgbloss"].detach()), G["mse"], rtol=2e-5)
    assert np.allclose(float(terms["interlevel"].detach()), G["interlevel"], rtol=2e-4, atol=1e-9)
    assert np.allclose(float(terms["distortion"].detach()), G["distortion"], rtol=2e-5)
    terms["loss"].backward()
    names = [str(n) for n in G["param_names"]]
    assert sorted(names) == sorted(k for k, _ in net.named_parameters())
    checked = 0
    for name in names:
        g = params[name].grad
        if f"gnone__{name}" in G.files:
            assert g is None or float(g.abs().max()) == 0.0, name
            continue
        assert g is not None, name
        ref_norm, ref_head = float(G[f"gnorm__{name}"]), G[f"ghead__{name}"]
        got = g.reshape(-1)
        # CPU fp32 accumulation order differs between the two op graphs only in the last bits
        assert abs(float(got.double().norm()) - ref_norm) <= 2e-4 * ref_norm + 1e-12, (name, float(got.double().norm()), ref_norm)
        assert np.allclose(got[:16].numpy(), ref_head, rtol=2e-3, atol=2e-4 * float(G[f"gabsmax__{name}"]) + 1e-12), name
        checked += 1
    assert checked > 40


def test_human_oracle_autograd_matches_reference_gradients():
    """Stage-2 human-object network, training mode (time = 0): surrogate objective mean(rgb) + 0.1 * cycle term."""
    from hosnerf_b200 import Network, default_cfg
    from oracle import human_ref as HR
    H = np.load(os.path.join(os.path.dirname(__file__), "golden", "human_s2_backward.npz"))
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    net = Network(default_cfg(), stage2=True)
    synth.fill_params_(net, 0)
    synth.boost_human_density_(net)
    pnames = {k for k, _ in net.named_parameters()}
    params = {k: v.detach().clone().requires_grad_(k in pnames and v.is_floating_point()) for k, v in net.state_dict().items()}
    b = synth.make_human_batch(12)
    res = HR.network_forward(params, b, stage2=True)
    assert res["observe_pts"].shape[0] == int(H["n_cycle_pts"])
    cyc = torch.mean(torch.sum((res["observe_pts"] - res["deform_pts_final"]) ** 2, 1) / 2.0)
    loss = res["rgb"].mean() + 0.1 * cyc
    assert np.allclose(float(loss.detach()), H["loss"], rtol=2e-5)
    assert np.allclose(float(cyc.detach()), H["cycle"], rtol=2e-4)
    loss.backward()
    names = [str(n) for n in H["param_names"]]
    assert sorted(names) == sorted(pnames)
    checked = 0
    for name in names:
        g = params[name].grad
        assert f"gnone__{name}" not in H.files and g is not None, name        # every parameter of the branch trains
        ref_norm, ref_head = float(H[f"gnorm__{name}"]), H[f"ghead__{name}"]
        got = g.reshape(-1)
        assert abs(float(got.double().norm()) - ref_norm) <= 5e-4 * ref_norm + 1e-12, (name, float(got.double().norm()), ref_norm)
        assert np.allclose(got[:16].numpy(), ref_head, rtol=5e-3, atol=5e-4 * float(H[f"gabsmax__{name}"]) + 1e-12), name
        checked += 1
    assert checked == 74
