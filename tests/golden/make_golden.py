"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

Every fixture stores the inputs, the outputs of the reference function/module,
and the ``fill_params_`` seed of the weights (weights themselves are regenerated
by name, see hosnerf_b200/synth.py).  tests/test_oracle_golden.py pins oracle/
to these fixtures; the GPU parity tests pin the CUDA path to them too.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types
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


def npify(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            out[k] = v.detach().cpu().numpy()
        elif isinstance(v, (int, float, bool, np.ndarray)):
            out[k] = np.asarray(v)
        elif v is None:
            continue
        else:
            raise TypeError((k, type(v)))
    return out


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **npify(arrays))
    print(f"wrote {name}.npz  ({os.path.getsize(path) / 1024:.1f} KiB)")


# ---------------------------------------------------------------------------
def helpers_s1():
    with rh.stage(rh.S1):
        import src.model.mipnerf360.helper as H
        g = torch.Generator().manual_seed(11)
        out = {}
        out["basis"] = H.generate_basis("icosahedron", 2)

        # max_dilate_weights ------------------------------------------------
        n, s = 6, 64
        t = torch.sort(torch.rand(n, s + 1, generator=g), -1).values
        t[:, 0], t[:, -1] = 0.0, 1.0
        t[2, 10] = t[2, 11]                       # a zero-width interval
        w = torch.rand(n, s, generator=g)
        w = w / w.sum(-1, keepdim=True)
        w[3, 5:9] = 0.0
        dil = 0.0025 + 0.5 / 64
        td, wd = H.max_dilate_weights(t, w, dil, (0.0, 1.0), True)
        out.update(md_t=t, md_w=w, md_dilation=dil, md_t_out=td, md_w_out=wd)

        # sample_intervals, deterministic + randomized ------------------------
        ts, ws = td[..., 1:-1], wd[..., 1:-1]
        logits = torch.where(ts[..., 1:] > ts[..., :-1], 1.0 * torch.log(ws), torch.full_like(ws, -torch.inf))
        out.update(si_t=ts, si_logits=logits)
        out["si_out_det"] = H.sample_intervals(False, ts, logits, 32, True, (0.0, 1.0))
        out["si_centers_det"] = H.sample(False, ts, logits, 32, True, True)
        torch.manual_seed(5)
        rand1 = torch.rand(n, 1)
        torch.manual_seed(5)
        out["si_out_rand_single"] = H.sample_intervals(True, ts, logits, 32, True, (0.0, 1.0))
        out["si_rand_single"] = rand1
        torch.manual_seed(6)
        randm = torch.rand(n, 32)
        torch.manual_seed(6)
        out["si_out_rand_multi"] = H.sample_intervals(True, ts, logits, 32, False, (0.0, 1.0))
        out["si_rand_multi"] = randm
        # first level: a single unit interval
        t01 = torch.tensor([[0.0, 1.0]]).repeat(n, 1)
        out["si_out_level0"] = H.sample_intervals(False, t01, torch.zeros(n, 1), 64, True, (0.0, 1.0))

        # gaussians / contract / ipe ----------------------------------------
        b = synth.make_bkg_batch(5, seed=21)
        _, s_to_t = H.construct_ray_warps(0.1, 1e6)
        sd = out["si_out_det"][:5, :17]
        tdist = s_to_t(sd)
        means, covs = H.cast_rays(tdist, b["rays_o"], b["rays_d"], b["radii"], "cone", diag=False)
        cm, cc = H.contract(means, covs, is_train=False)
        lm, lv = H.lift_and_diagonalize(cm, cc, out["basis"])
        feat = H.integrated_pos_enc(lm, lv, 0, 12)
        out.update(g_sdist=sd, g_tdist=tdist, g_rays_o=b["rays_o"], g_rays_d=b["rays_d"], g_radii=b["radii"],
                   g_means=means, g_covs=covs, g_cmeans=cm, g_ccovs=cc, g_lmean=lm, g_lvar=lv, g_feat=feat)
        out["g_direnc"] = H.pos_enc(b["viewdirs"], 0, 4, True)
        out["g_viewdirs"] = b["viewdirs"]

        # composite ---------------------------------------------------------
        dens = torch.rand(5, 16, generator=g) * 3.0
        rgbs = torch.rand(5, 16, 3, generator=g)
        for opaque in (False, True):
            wts, alpha, trans = H.compute_alpha_weights(dens, tdist, b["rays_d"], opaque_background=opaque)
            rend = H.volumetric_rendering(rgbs, wts, tdist, 1.0, 1e6, False)["rgb"]
            tag = "op" if opaque else "tr"
            out.update({f"c_w_{tag}": wts, f"c_alpha_{tag}": alpha, f"c_trans_{tag}": trans, f"c_rgb_{tag}": rend})
        out.update(c_density=dens, c_rgbs=rgbs)
        save("s1_helpers", **out)


# ---------------------------------------------------------------------------
def _history_arrays(prefix, renderings, history):
    out = {}
    for i, h in enumerate(history):
        for k, v in h.items():
            out[f"{prefix}L{i}_{k}"] = v
    for i, r in enumerate(renderings):
        out[f"{prefix}R{i}_rgb"] = r["rgb"]
    return out


def forward_s1():
    with rh.stage(rh.S1):
        import src.model.mipnerf360.model as M
        # default Backpack.gin shape: 64/64/32, NeRFMLP 1024 wide -------------------
        net = M.MipNeRF360("/nonexistent", opaque_background=True)
        synth.fill_params_(net, 0)
        b = synth.make_bkg_batch(16, seed=1)
        with torch.no_grad():
            r, h = net(b, 1.0, False, False, 0.1, 1e6)
        out = _history_arrays("", r, h)
        out.update({f"in_{k}": v for k, v in b.items()})
        save("s1_forward_default", **out)

        # randomized=True (train-style jitter), train_frac 0.3 -> anneal != 1 ----
        torch.manual_seed(77)
        rands = [torch.rand(16, 1) for _ in range(3)]
        torch.manual_seed(77)
        with torch.no_grad():
            r, h = net(b, 0.3, True, False, 0.1, 1e6)
        out = _history_arrays("", r, h)
        out.update({f"in_{k}": v for k, v in b.items()})
        out.update({f"rand{i}": x for i, x in enumerate(rands)})
        save("s1_forward_default_rand", **out)

        # C2 shape: gin bindings num_levels=2, 128/128, NeRFMLP.netwidth=256 -------
        # (gin overrides keyword defaults; emulate `NeRFMLP.netwidth = 256` the same way)
        saved = M.NeRFMLP.__init__.__defaults__
        M.NeRFMLP.__init__.__defaults__ = (8, 256)
        try:
            net = M.MipNeRF360("/nonexistent", num_prop_samples=128, num_nerf_samples=128,
                               num_levels=2, opaque_background=True)
        finally:
            M.NeRFMLP.__init__.__defaults__ = saved
        synth.fill_params_(net, 0)
        b = synth.make_bkg_batch(12, seed=1)
        with torch.no_grad():
            r, h = net(b, 1.0, False, False, 0.1, 1e6)
        out = _history_arrays("", r, h)
        out.update({f"in_{k}": v for k, v in b.items()})
        save("s1_forward_c2", **out)

        # state-conditional: 2 transitions -> 3 embeddings, time in the middle state
        with tempfile.TemporaryDirectory() as td:
            with open(os.path.join(td, "transitions_times.json"), "w") as f:
                json.dump({"a": {"time": 0.25}, "b": {"time": 0.6}}, f)
            net = M.MipNeRF360(td, opaque_background=True)
        synth.fill_params_(net, 0)
        b = synth.make_bkg_batch(8, seed=3, time=0.4)
        with torch.no_grad():
            r, h = net(b, 1.0, False, False, 0.1, 1e6)
        out = _history_arrays("", r, h)
        out.update({f"in_{k}": v for k, v in b.items()})
        out["transitions_times"] = np.array([0.25, 0.6], np.float32)
        save("s1_forward_states", **out)


# ---------------------------------------------------------------------------
HUMAN_KEYS = ("human_rgb", "human_density", "newsmpl_pts", "pts_mask", "z_vals", "rays_d",
              "deform_pts_final", "observe_pts", "rgb", "alpha", "depth", "weights")


def _human_inputs(b):
    # the 27x32^3 prior volume is regenerated by synth.make_skeleton (3.5 MB otherwise)
    return {f"in_{k}": v for k, v in b.items()
            if isinstance(v, torch.Tensor) and k != "motion_weights_priors"}


def forward_human():
    for tag, sdir in (("s3", rh.S3), ("s2", rh.S2)):
        with rh.stage(sdir):
            import core.nets.human_nerf.network as N
            cfg = rh.human_cfg(sdir)
            cfg.perturb = 0.0
            net = N.Network(cfg)
            synth.fill_params_(net, 0)
            with torch.no_grad():
                net.cnl_mlp.output_linear[0].bias[3] += 3.0      # see synth.boost_human_density_
            b = synth.make_human_batch(40)
            with torch.no_grad():
                out = net(**b)
            arrays = {k: out[k] for k in HUMAN_KEYS if k in out}
            arrays.update(_human_inputs(b))
            save(f"human_{tag}_eval", **arrays)

            if tag == "s3":
                # stratified jitter (cfg.perturb > 0), early iteration: hann window partly
                # open, pose decoder off, non-rigid condition zeroed
                cfg.perturb = 1.0
                b = synth.make_human_batch(24, iter_val=150000.0)
                torch.manual_seed(9)
                rand = torch.rand(24, cfg.N_samples)
                torch.manual_seed(9)
                with torch.no_grad():
                    out = net(**b)
                arrays = {k: out[k] for k in HUMAN_KEYS if k in out}
                arrays.update(_human_inputs(b))
                arrays["rand"] = rand
                save("human_s3_jitter", **arrays)
                cfg.perturb = 0.0
                b = synth.make_human_batch(24, iter_val=5000.0)
                with torch.no_grad():
                    out = net(**b)
                arrays = {k: out[k] for k in HUMAN_KEYS if k in out}
                arrays.update(_human_inputs(b))
                save("human_s3_early", **arrays)

                # individual pieces
                Rb, Tb, Rf, Tf = net.motion_basis_computer(b["dst_Rs"][None], b["dst_Ts"][None], b["cnl_gtfms"][None])
                with torch.no_grad():
                    vol = net.mweight_vol_decoder(motion_weights_priors=b["motion_weights_priors"][None])[0]
                g = torch.Generator().manual_seed(4)
                pts = torch.rand(300, 3, generator=g) * 1.6 - 0.8
                mv = N.Network._sample_motion_fields(pts[None], Rb[0], Tb[0], vol, b["cnl_bbox_min_xyz"],
                                                     b["cnl_bbox_scale_xyz"], ["x_skel", "fg_likelihood_mask"])
                mvf = N.Network._sample_motion_fields_forward(pts, Rf[0], Tf[0], vol, b["cnl_bbox_min_xyz"],
                                                              b["cnl_bbox_scale_xyz"], ["x_deform", "fg_likelihood_mask_forward"])
                save("human_lbs", pts=pts, Rb=Rb[0], Tb=Tb[0], Rf=Rf[0], Tf=Tf[0],
                     vol_center=vol[:, 12:20, 12:20, 12:20].contiguous(), vol_sum=vol.sum(),
                     x_skel=mv["x_skel"][0], mask=mv["fg_likelihood_mask"][0],
                     x_deform=mvf["x_deform"], mask_fwd=mvf["fg_likelihood_mask_forward"],
                     **_human_inputs(b))


# ---------------------------------------------------------------------------
def composite_s3():
    """Run the reference's own training_step composite (S3 model.py:1501-1596) on CPU:
    Tensor.cuda patched to identity, ``self`` duck-typed."""
    with rh.stage(rh.S3):
        import src.model.mipnerf360.model as M
        import core.nets.human_nerf.network as N
        cfg = rh.human_cfg(rh.S3)
        cfg.perturb = 0.0
        n = 48
        saved = M.NeRFMLP.__init__.__defaults__
        M.NeRFMLP.__init__.__defaults__ = (8, 256)
        try:
            bkg = M.MipNeRF360("/nonexistent", opaque_background=True)
        finally:
            M.NeRFMLP.__init__.__defaults__ = saved
        synth.fill_params_(bkg, 0)
        human = N.Network(cfg)
        synth.fill_params_(human, 0)
        with torch.no_grad():
            human.cnl_mlp.output_linear[0].bias[3] += 3.0

        hb = synth.make_human_batch(n, time=0.0, is_train=True)
        Mw = synth.random_rigid()
        # background rays = the human rays mapped into the scale-world frame
        ro, rd = hb["rays"][0], hb["rays"][1]
        ro_w = (Mw[:3, :3] @ ro.T).T + Mw[:3, 3]
        rd_w = (Mw[:3, :3] @ rd.T).T
        batch = dict(hb)
        batch.update(rays_o_bkg=ro_w, rays_d_bkg=rd_w, viewdirs_bkg=rd_w / rd_w.norm(dim=-1, keepdim=True),
                     radii=torch.full((n, 1), 1e-3), newsmpl_to_scale_world=Mw,
                     patch_masks=torch.zeros(1), target_patches=torch.zeros(1), patch_div_indices=torch.zeros(1))
        captured = {}

        def get_loss(net_output, idx_fg, human_weights_onlyfg, **kw):
            captured.update(rgb=net_output["rgb"], idx_fg=idx_fg, human_w=human_weights_onlyfg,
                            human_rgb=net_output["human_rgb"], human_density=net_output["human_density"],
                            pts_mask=net_output["pts_mask"], newsmpl_pts=net_output["newsmpl_pts"])
            z = torch.zeros(())
            return z, {"mse": z, "lpips": z, "cycle": 0.0}

        duck = types.SimpleNamespace(
            trainer=types.SimpleNamespace(global_step=7), model=bkg, human=human, near_bkg=0.1,
            far_bkg=1e6, cfg=cfg, get_loss=get_loss, log=lambda *a, **k: None, progress=lambda: False)
        hist = {}
        orig_fwd = bkg.forward

        def spy(*a, **k):
            torch.manual_seed(123)            # pins the single_jitter draw of every level
            r = orig_fwd(*a, **k)
            hist["h"] = r[1][-1]
            return r
        bkg.forward = spy
        old_cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            with torch.no_grad():
                M.LitMipNeRF360.training_step(duck, {k: (v[None] if isinstance(v, torch.Tensor) else [v])
                                                     for k, v in batch.items()}, 0)
        finally:
            torch.Tensor.cuda = old_cuda
        h = hist["h"]
        save("s3_composite", rgb=captured["rgb"], idx_fg=captured["idx_fg"], human_w=captured["human_w"],
             human_rgb=captured["human_rgb"], human_density=captured["human_density"],
             pts_mask=captured["pts_mask"], newsmpl_pts=captured["newsmpl_pts"],
             bkg_rgb=h["rgb"], bkg_density=h["density"], bkg_tdist=h["tdist"], bkg_sdist=h["sdist"],
             M=Mw, rays_o_bkg=ro_w, rays_d_bkg=rd_w)


if __name__ == "__main__":
    torch.set_num_threads(8)
    helpers_s1()
    forward_s1()
    forward_human()
    composite_s3()
