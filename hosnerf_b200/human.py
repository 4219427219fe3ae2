"""Drop-in replacement for the human-object branch
(S3/core/nets/human_nerf/network.py ``Network`` and the component classes it loads
through ``cfg.*.module``; the S2 variant returns the composited ray instead).

Parameter names / shapes match the reference ``state_dict`` one to one
(``cnl_mlp.pts_linears.{0,2,..}``, ``non_rigid_mlp.block_mlps.{0,2,..}``,
``mweight_vol_decoder.decoder.block_conv.{0,2,..}``, ``pose_decoder.*``,
``human_stateembeds.{k}``), so stage-2/3 checkpoints load unchanged.

Per-ray hot path (kernels of libhosnerf_b200.so): stratified samples -> LBS warp
(26 bone transforms + trilinear motion-weight gather) -> Hann-windowed PE ->
non-rigid MLP -> Fourier PE -> canonical MLP (-> S2 composite).
Per-frame prologue kept in torch (out of scope per SURVEY section 2, rows 13-14):
pose refiner, kinematic chain / motion bases, ConvTranspose3d volume decoder
(cached across the chunks of a frame).
"""
from __future__ import annotations

import json
import math
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .mip360 import _DEFAULT_PRECISION, select_state_index  # noqa: F401
from . import mip360 as _m
from .synth import PARENT as SMPL_PARENT


# eval chunks in fp16 mode go through ONE library call (hos_render_human) when the cycle / flow side paths are not asked for;
# False keeps the kernel-by-kernel chain (same kernels, same order, same results - used by the A/B test)
ONE_CALL = True
# inside the one-call path: generate both positional encodings in the MLP kernels' prologue (False: materialise them in HBM first)
FUSE_FOURIER = True


class Cfg(dict):
    """Minimal attribute dict standing in for the reference's yacs CfgNode in tests / bench."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def default_cfg(**over):
    """configs/default.yaml + configs/human_nerf/wild/monocular/adventure.yaml (S3)."""
    cfg = Cfg(
        basedir="/nonexistent", total_bones=26, N_samples=128, perturb=0.0, chunk=8192,
        netchunk_per_gpu=10000, ignore_non_rigid_motions=False, bgcolor=[255.0, 255.0, 255.0],
        canonical_mlp=Cfg(mlp_depth=8, mlp_width=256, multires=10, i_embed=0),
        mweight_volume=Cfg(embedding_size=256, volume_size=32, dst_voxel_size=0.0625),
        non_rigid_motion_mlp=Cfg(condition_code_size=75, mlp_width=128, mlp_depth=6, skips=[4], multires=6,
                                 i_embed=0, kick_in_iter=100000, full_band_iter=200000),
        non_rigid_forward_mlp=Cfg(condition_code_size=75, mlp_width=128, mlp_depth=6, skips=[4], multires=6,
                                  i_embed=0, kick_in_iter=0, full_band_iter=0),
        pose_decoder=Cfg(embedding_size=75, mlp_width=256, mlp_depth=4, kick_in_iter=20000),
    )
    cfg.update(over)
    return cfg


# ----------------------------------------------------------------------------- init helpers
def _xavier_std(m, gain):
    """network_util.py:181-236."""
    if isinstance(m, nn.ConvTranspose3d):
        k = m.kernel_size[0] * m.kernel_size[1] * m.kernel_size[2] // m.stride[0] // m.stride[1] // m.stride[2]
        return gain * math.sqrt(2.0 / ((m.in_channels + m.out_channels) * k))
    if isinstance(m, nn.Linear):
        return gain * math.sqrt(2.0 / (m.in_features + m.out_features))
    return None


@torch.no_grad()
def _init_mod(m, gain=1.0):
    std = _xavier_std(m, gain)
    if std is None:
        return
    m.weight.uniform_(-(std * math.sqrt(3.0)), std * math.sqrt(3.0))
    if m.bias is not None:
        m.bias.zero_()
    if isinstance(m, nn.ConvTranspose3d):      # block-wise constant init of stride-2 deconvs
        base = m.weight[:, :, 0::2, 0::2, 0::2].clone()
        for a in (0, 1):
            for b in (0, 1):
                for c in (0, 1):
                    m.weight[:, :, a::2, b::2, c::2] = base


def _init_seq(seq):
    """network_util.py:292-308: gain chosen from the activation that follows."""
    mods = list(seq)
    for a, b in zip(mods[:-1], mods[1:]):
        if isinstance(b, nn.ReLU):
            _init_mod(a, nn.init.calculate_gain("relu"))
        elif isinstance(b, nn.LeakyReLU):
            _init_mod(a, nn.init.calculate_gain("leaky_relu", b.negative_slope))
        else:
            _init_mod(a)
    _init_mod(mods[-1])


# ----------------------------------------------------------------------------- parameter containers
class CanonicalMLP(nn.Module):
    """mlp_rgb_sigma.py:16-58 (8x256, cat([pos_embed, h]) before layer 5, 4 outputs)."""

    def __init__(self, mlp_depth=8, mlp_width=256, input_ch=3, skips=None, **_):
        super().__init__()
        skips = [4] if skips is None else skips
        self.mlp_depth, self.mlp_width, self.input_ch = mlp_depth, mlp_width, input_ch
        blocks = [nn.Linear(input_ch, mlp_width), nn.ReLU()]
        self.layers_to_cat_input = []
        for i in range(mlp_depth - 1):
            if i in skips:
                self.layers_to_cat_input.append(len(blocks))
                blocks += [nn.Linear(mlp_width + input_ch, mlp_width), nn.ReLU()]
            else:
                blocks += [nn.Linear(mlp_width, mlp_width), nn.ReLU()]
        self.pts_linears = nn.ModuleList(blocks)
        _init_seq(self.pts_linears)
        self.output_linear = nn.Sequential(nn.Linear(mlp_width, 4))
        _init_seq(self.output_linear)

    def linears(self):
        return [m for m in self.pts_linears if isinstance(m, nn.Linear)]

    def skip_layer_indices(self):
        return [i // 2 for i in self.layers_to_cat_input]


class NonRigidMotionMLP(nn.Module):
    """mlp_offset.py:16-70 (6x128, input cat(condition, pos_embed), cat([h, pos_embed]) before layer 4)."""

    def __init__(self, pos_embed_size=3, condition_code_size=69, mlp_width=128, mlp_depth=6, skips=None):
        super().__init__()
        self.skips = [4] if skips is None else list(skips)
        self.pos_embed_size, self.condition_code_size = pos_embed_size, condition_code_size
        self.mlp_width, self.mlp_depth = mlp_width, mlp_depth
        blocks = [nn.Linear(pos_embed_size + condition_code_size, mlp_width), nn.ReLU()]
        self.layers_to_cat_inputs = []
        for i in range(1, mlp_depth):
            if i in self.skips:
                self.layers_to_cat_inputs.append(len(blocks))
                blocks += [nn.Linear(mlp_width + pos_embed_size, mlp_width), nn.ReLU()]
            else:
                blocks += [nn.Linear(mlp_width, mlp_width), nn.ReLU()]
        blocks += [nn.Linear(mlp_width, 3)]
        self.block_mlps = nn.ModuleList(blocks)
        _init_seq(self.block_mlps)
        with torch.no_grad():   # start from ~zero offsets
            self.block_mlps[-1].weight.uniform_(-1e-5, 1e-5)
            self.block_mlps[-1].bias.zero_()

    def linears(self):
        return [m for m in self.block_mlps if isinstance(m, nn.Linear)]

    def skip_layer_indices(self):
        return [i // 2 for i in self.layers_to_cat_inputs]


NonRigidForwardMLP = NonRigidMotionMLP   # mlp_forward_offset.py is the same architecture


class ConvDecoder3D(nn.Module):
    """network_util.py:21-59."""

    def __init__(self, embedding_size=256, volume_size=128, voxel_channels=4):
        super().__init__()
        self.block_mlp = nn.Sequential(nn.Linear(embedding_size, 1024), nn.LeakyReLU(0.2))
        convs = []
        cin, cout = 1024, 512
        for _ in range(int(np.log2(volume_size)) - 1):
            convs += [nn.ConvTranspose3d(cin, cout, 4, 2, 1), nn.LeakyReLU(0.2)]
            if cin == cout:
                cout = cin // 2
            else:
                cin = cout
        convs.append(nn.ConvTranspose3d(cin, voxel_channels, 4, 2, 1))
        self.block_conv = nn.Sequential(*convs)
        for m in (self.block_mlp, self.block_conv):
            _init_seq(m)

    def forward(self, embedding):
        return self.block_conv(self.block_mlp(embedding).view(-1, 1024, 1, 1, 1))


class MotionWeightVolumeDecoder(nn.Module):
    """deconv_vol_decoder.py:17-42: softmax(decoder(const) + log prior) over 27 channels."""

    def __init__(self, embedding_size=256, volume_size=32, total_bones=24):
        super().__init__()
        self.total_bones, self.volume_size = total_bones, volume_size
        self.const_embedding = nn.Parameter(torch.randn(embedding_size), requires_grad=True)
        self.decoder = ConvDecoder3D(embedding_size=embedding_size, volume_size=volume_size,
                                     voxel_channels=total_bones + 1)

    def forward(self, motion_weights_priors, **_):
        return F.softmax(self.decoder(self.const_embedding[None, ...]) + torch.log(motion_weights_priors), dim=1)


class BodyPoseRefiner(nn.Module):
    """pose_decoders/mlp_delta_body_pose.py:14-73."""

    def __init__(self, total_bones=23, embedding_size=69, mlp_width=256, mlp_depth=4, **_):
        super().__init__()
        blocks = [nn.Linear(embedding_size, mlp_width), nn.ReLU()]
        for _i in range(0, mlp_depth - 2):
            blocks += [nn.Linear(mlp_width, mlp_width), nn.ReLU()]
        self.total_bones = total_bones - 1

        def branch():
            b = [nn.Linear(mlp_width, mlp_width), nn.ReLU()]
            for _i in range(3, mlp_depth - 1):
                b += [nn.Linear(mlp_width, mlp_width), nn.ReLU()]
            b += [nn.Linear(mlp_width, 3 * self.total_bones)]
            return nn.Sequential(*b)
        self.block_mlps = nn.Sequential(*blocks)
        _init_seq(self.block_mlps)
        self.block_mlps_dstR = branch()
        _init_seq(self.block_mlps_dstR)
        self.block_mlps_dstT = branch()
        _init_seq(self.block_mlps_dstT)
        with torch.no_grad():
            for seq in (self.block_mlps_dstR, self.block_mlps_dstT):
                seq[-1].weight.uniform_(-1e-5, 1e-5)
                seq[-1].bias.zero_()

    @staticmethod
    def _rodrigues(rvec):   # network_util.py:66-92
        theta = torch.sqrt(1e-5 + torch.sum(rvec ** 2, dim=1))
        r = rvec / theta[:, None]
        c, s = torch.cos(theta), torch.sin(theta)
        x, y, z = r[:, 0], r[:, 1], r[:, 2]
        return torch.stack((x * x + (1. - x * x) * c, x * y * (1. - c) - z * s, x * z * (1. - c) + y * s,
                            x * y * (1. - c) + z * s, y * y + (1. - y * y) * c, y * z * (1. - c) - x * s,
                            x * z * (1. - c) - y * s, y * z * (1. - c) + x * s, z * z + (1. - z * z) * c),
                           dim=1).view(-1, 3, 3)

    def forward(self, pose_input):
        h = self.block_mlps(pose_input)
        Rs = self._rodrigues(self.block_mlps_dstR(h).view(-1, 3)).view(-1, self.total_bones, 3, 3)
        Ts = self.block_mlps_dstT(h).view(-1, self.total_bones, 3)
        return {"Rs": Rs, "Ts": Ts}


class MotionBasisComputer(nn.Module):
    """network_util.py:106-174: kinematic chain -> backward / forward bone maps.  26 bones of
    4x4 algebra: evaluated on the host (a chain of 25 dependent 4x4 products is ~100 GPU
    launches otherwise) and shipped to the device as two [26,3,3] / [26,3] tensors."""

    def __init__(self, total_bones=24):
        super().__init__()
        self.total_bones = total_bones

    def forward(self, dst_Rs, dst_Ts, cnl_gtfms):
        dev = cnl_gtfms.device
        Rs, Ts, cg = dst_Rs.detach().cpu(), dst_Ts.detach().cpu(), cnl_gtfms.detach().cpu()
        b, nb = Rs.shape[:2]
        local = torch.zeros(b, nb, 4, 4, dtype=Rs.dtype)
        local[:, :, :3, :3] = Rs
        local[:, :, :3, 3] = Ts
        local[:, :, 3, 3] = 1.0
        glob = torch.zeros_like(cg)
        glob[:, 0] = local[:, 0]
        for i in range(1, nb):
            glob[:, i] = torch.matmul(glob[:, SMPL_PARENT[i]].clone(), local[:, i])
        glob = glob.view(-1, 4, 4)
        cgf = cg.view(-1, 4, 4)
        back = torch.matmul(cgf, torch.inverse(glob)).view(-1, nb, 4, 4)
        fwd = torch.matmul(glob, torch.inverse(cgf)).view(-1, nb, 4, 4)
        return (back[:, :, :3, :3].contiguous().to(dev), back[:, :, :3, 3].contiguous().to(dev),
                fwd[:, :, :3, :3].contiguous().to(dev), fwd[:, :, :3, 3].contiguous().to(dev))


def hann_window_weights(n_freqs, iter_val, kick_in_iter, full_band_iter):
    """hannw_fourier.py:33-44 (host, float32 torch like the reference)."""
    kick = torch.tensor(kick_in_iter, dtype=torch.float32)
    it = iter_val.detach().cpu().float() if isinstance(iter_val, torch.Tensor) else torch.tensor(float(iter_val))
    t = torch.clamp(it - kick, min=0.)
    alpha = n_freqs * t / (full_band_iter - kick)
    w = [(1. - torch.cos(np.pi * torch.clamp(alpha - k, min=0., max=1.))) / 2. for k in range(n_freqs)]
    return torch.stack([x.reshape(()) for x in w]).float()


# The second plug-in seam of the reference: ``cfg.<component>.module`` names the python file a component class is loaded from
# (S3/core/nets/human_nerf/component_factory.py:12-40).  This package ships one kernel-backed implementation per component:
# a config that names the reference's default module gets it; a config that swaps a component for something else must not
# silently get the built-in one.
_BUILTIN_MODULES = {
    "embedder": "core.nets.human_nerf.embedders.fourier",
    "non_rigid_embedder": "core.nets.human_nerf.embedders.hannw_fourier",
    "canonical_mlp": "core.nets.human_nerf.canonical_mlps.mlp_rgb_sigma",
    "mweight_volume": "core.nets.human_nerf.mweight_vol_decoders.deconv_vol_decoder",
    "non_rigid_motion_mlp": "core.nets.human_nerf.non_rigid_motion_mlps.mlp_offset",
    "non_rigid_forward_mlp": "core.nets.human_nerf.non_rigid_motion_mlps.mlp_forward_offset",
    "pose_decoder": "core.nets.human_nerf.pose_decoders.mlp_delta_body_pose",
}


def check_component_modules(cfg):
    """Raise if ``cfg.<component>.module`` asks for a component implementation other than the reference default."""
    for comp, default in _BUILTIN_MODULES.items():
        node = cfg.get(comp) if hasattr(cfg, "get") else getattr(cfg, comp, None)
        mod = None
        if node is not None:
            mod = node.get("module") if hasattr(node, "get") else getattr(node, "module", None)
        if mod is not None and str(mod) != default:
            raise NotImplementedError(
                f"hosnerf_b200.Network: cfg.{comp}.module = {mod!r} selects a component this package has no kernels for "
                f"(built in: {default!r}, S3 component_factory.py:12-40)")


# ----------------------------------------------------------------------------- Network
class Network(nn.Module):
    """S3 network.py:27-698.  ``stage2=True`` gives the S2 return dict (rgb/alpha/depth/weights)."""

    def __init__(self, cfg, stage2: bool = False, precision=None):
        super().__init__()
        check_component_modules(cfg)
        self.cfg = cfg
        self.stage2 = stage2
        self.precision = precision
        # True: the training forward keeps every shape static and reads nothing back to the host (device-side bone chain, dense
        # cycle side path with ``cycle_mask``, device jitter) so a whole step can be captured in a CUDA graph (train.GraphedStep)
        self.static_shapes = False
        nb = cfg.total_bones
        self.motion_basis_computer = MotionBasisComputer(total_bones=nb)
        self.mweight_vol_decoder = MotionWeightVolumeDecoder(
            embedding_size=cfg.mweight_volume.embedding_size, volume_size=cfg.mweight_volume.volume_size,
            total_bones=nb)
        nr = cfg.non_rigid_motion_mlp
        self.nr_freqs = nr.multires
        nr_embed = 6 * nr.multires if nr.i_embed != -1 else 3
        self.non_rigid_mlp = NonRigidMotionMLP(pos_embed_size=nr_embed, condition_code_size=nr.condition_code_size,
                                               mlp_width=nr.mlp_width, mlp_depth=nr.mlp_depth, skips=nr.skips)
        nf = cfg.non_rigid_forward_mlp
        self.non_rigid_forward_mlp = NonRigidForwardMLP(
            pos_embed_size=nr_embed, condition_code_size=nf.condition_code_size, mlp_width=nf.mlp_width,
            mlp_depth=nf.mlp_depth, skips=nf.skips)
        cm = cfg.canonical_mlp
        self.cnl_freqs = cm.multires
        cnl_embed = 3 + 6 * cm.multires if cm.i_embed != -1 else 3
        self.embedding_size = 64
        tt_path = os.path.join(cfg.basedir, "transitions_times.json")
        if os.path.exists(tt_path):
            with open(tt_path, "r") as f:
                infos = json.load(f)
            self.transitions_times = np.stack([np.array(infos[k]["time"], dtype=np.float32) for k in infos], 0)
            n_states = self.transitions_times.shape[0] + 1
        else:
            self.transitions_times, n_states = None, 1
        self.human_stateembeds = nn.ParameterList(
            [nn.Parameter(torch.randn(self.embedding_size), requires_grad=True) for _ in range(n_states)])
        self.cnl_mlp = CanonicalMLP(input_ch=cnl_embed + self.embedding_size, mlp_depth=cm.mlp_depth,
                                    mlp_width=cm.mlp_width, skips=[4])
        pd = cfg.pose_decoder
        self.pose_decoder = BodyPoseRefiner(total_bones=nb, embedding_size=pd.embedding_size,
                                            mlp_width=pd.mlp_width, mlp_depth=pd.mlp_depth)
        self._cache = {}
        if cm.i_embed == -1 or nr.i_embed == -1:
            raise NotImplementedError("hosnerf_b200: identity embedders (i_embed=-1) are not built")

    # ------------------------------------------------------------------ weight preparation
    def _versions(self, mods):
        return tuple(p._version for m in mods for p in m.parameters())

    def _nr_folded(self, mlp: NonRigidMotionMLP, cond):
        """Fold the per-frame condition code into the first-layer bias: W0[:, :75] @ cond."""
        lins = mlp.linears()
        c = mlp.condition_code_size
        skips = mlp.skip_layer_indices()
        out = []
        for i, m in enumerate(lins[:-1]):
            W, b = m.weight.detach(), m.bias.detach()
            if i == 0:
                out.append((W[:, c:].contiguous(), (b + W[:, :c] @ cond.reshape(-1)).contiguous(), False))
            else:
                out.append((W.contiguous(), b.contiguous(), i in skips))      # skip: native [h | pe]
        head = (lins[-1].weight.detach().contiguous(), lins[-1].bias.detach().contiguous())
        return out, head

    def _cnl_folded(self, state_idx):
        key = ("cnl32", state_idx, self._versions([self.cnl_mlp]), self.human_stateembeds[state_idx]._version)
        if self._cache.get("cnl32_key") == key:
            return self._cache["cnl32"]
        e = self.human_stateembeds[state_idx].detach()
        lins = self.cnl_mlp.linears()
        skips = self.cnl_mlp.skip_layer_indices()
        pe = self.cnl_mlp.input_ch - self.embedding_size
        out = []
        for i, m in enumerate(lins):
            W, b = m.weight.detach(), m.bias.detach()
            if i == 0:
                out.append((W[:, :pe].contiguous(), (b + W[:, pe:] @ e).contiguous(), False))
            elif i in skips:    # native [pe | emb | h] -> kernel order [h | pe], emb folded
                nin = self.cnl_mlp.input_ch
                Wc = torch.cat([W[:, nin:], W[:, :pe]], dim=1).contiguous()
                out.append((Wc, (b + W[:, pe:nin] @ e).contiguous(), True))
            else:
                out.append((W.contiguous(), b.contiguous(), False))
        head = (self.cnl_mlp.output_linear[0].weight.detach().contiguous(),
                self.cnl_mlp.output_linear[0].bias.detach().contiguous())
        self._cache["cnl32_key"], self._cache["cnl32"] = key, (out, head)
        return out, head

    def _fused_nr(self, which: str, mlp: NonRigidMotionMLP, cond):
        """tcgen05 program of a non-rigid MLP; weights are uploaded once per parameter version,
        only the condition-folded first-layer bias is refreshed per call (per frame)."""
        key = (which, self._versions([mlp]))
        serial = (id(cond), cond._version)                 # the frame cache keeps `cond` alive, so the id is not recycled
        if self._cache.get(which + "_key") == key and self._cache.get(which + "_frame") == serial:
            return self._cache[which]                      # same weights, same condition code: folded bias already uploaded
        layers32, head32 = self._nr_folded(mlp, cond)
        if self._cache.get(which + "_key") != key:
            pe, w = mlp.pos_embed_size, mlp.mlp_width
            layers = []
            for i, (_, _, skip) in enumerate(layers32):
                layers.append(dict(out_dim=w, in_h=0 if i == 0 else w, in_x=pe if (i == 0 or skip) else 0,
                                   x_first=0, relu=1, rowbias=0, head=-1))
            layers[-1]["head"] = 0
            fm = ops.FusedMLP(pe, layers, [dict(out_dim=3, post=3, shift=0.0, out_slot=0)])
            for i, (W, b, _) in enumerate(layers32):
                fm.set_layer(i, W, b)
            fm.set_head(0, *head32)
            self._cache[which + "_key"], self._cache[which] = key, fm
        fm = self._cache[which]
        fm.set_bias(0, layers32[0][1])
        self._cache[which + "_frame"], self._cache[which + "_cond"] = serial, cond
        return fm

    def _fused_cnl(self, state_idx):
        key = ("cnl16", state_idx, self._versions([self.cnl_mlp]), self.human_stateembeds[state_idx]._version,
               self.stage2)
        if self._cache.get("cnl16_key") == key:
            return self._cache["cnl16"]
        layers32, head32 = self._cnl_folded(state_idx)
        pe = self.cnl_mlp.input_ch - self.embedding_size
        w = self.cnl_mlp.mlp_width
        layers = []
        for i, (_, _, skip) in enumerate(layers32):
            layers.append(dict(out_dim=w, in_h=0 if i == 0 else w, in_x=pe if (i == 0 or skip) else 0,
                               x_first=0, relu=1, rowbias=0, head=-1))
        layers[-1]["head"] = 0
        fm = ops.FusedMLP(pe, layers, [dict(out_dim=4, post=0 if self.stage2 else 4, shift=0.0, out_slot=0)])
        for i, (W, b, _) in enumerate(layers32):
            fm.set_layer(i, W, b)
        fm.set_head(0, *head32)
        self._cache["cnl16_key"], self._cache["cnl16"] = key, fm
        return fm

    def _volume(self, priors):
        key = (self._versions([self.mweight_vol_decoder]), priors.data_ptr(), priors._version, tuple(priors.shape))
        if self._cache.get("vol_key") != key:
            with torch.no_grad():
                vol = self.mweight_vol_decoder(motion_weights_priors=priors)[0].contiguous()
            self._cache["vol_key"], self._cache["vol"] = key, vol
        return self._cache["vol"]

    # ------------------------------------------------------------------ MLP evaluation
    def _eval_non_rigid(self, which, mlp, x, cond, hann_w, precision):
        """x [P,3] -> x + offset(x)   (network.py:165-172 / 486-495 / 521-530)."""
        if precision == "fp16":
            fm = self._fused_nr(which, mlp, cond)
            pe = ops.fourier_embed(x, self.nr_freqs, False, hann_w, out="tiled")
            return fm.forward(pe, x.shape[0], add=x)[0]
        layers32, head32 = self._nr_folded(mlp, cond)
        pe = ops.fourier_embed(x, self.nr_freqs, False, hann_w)
        h = pe
        for W, b, skip in layers32:
            h = ops.linear_f32(h, W, b, act=1, x2=pe if skip else None)
        return ops.head_f32(h, *head32, post=3, add=x)

    def _eval_canonical(self, xyz, state_idx, precision):
        """xyz [P,3] -> [P,4]: S3 (sigmoid rgb, relu sigma) or S2 raw."""
        post = 0 if self.stage2 else 4
        if precision == "fp16":
            fm = self._fused_cnl(state_idx)
            pe = ops.fourier_embed(xyz, self.cnl_freqs, True, None, out="tiled")
            return fm.forward(pe, xyz.shape[0])[0]
        layers32, head32 = self._cnl_folded(state_idx)
        pe = ops.fourier_embed(xyz, self.cnl_freqs, True, None)
        h = pe
        for W, b, skip in layers32:
            h = ops.linear_f32(h, W, b, act=1, x2=pe if skip else None)
        return ops.head_f32(h, *head32, post=post)

    # ------------------------------------------------------------------ forward
    def _correct_pose(self, dst_Rs, dst_Ts, posevec):
        """Pose refinement (network.py:590-605) on the HOST: a 4-layer MLP on one 75-vector and 25
        3x3 products per frame - a handful of microseconds on the CPU versus ~20 kernel launches."""
        key = self._versions([self.pose_decoder])
        if self._cache.get("pose_key") != key:
            self._cache["pose_key"] = key
            self._cache["pose_cpu"] = {k: v.detach().cpu() for k, v in self.pose_decoder.state_dict().items()}
        sd = self._cache["pose_cpu"]

        def seq(prefix, x, n_relu_last):
            idxs = sorted({int(k.split(".")[1]) for k in sd if k.startswith(prefix + ".") and k.endswith("weight")})
            for j, i in enumerate(idxs):
                x = F.linear(x, sd[f"{prefix}.{i}.weight"], sd[f"{prefix}.{i}.bias"])
                if n_relu_last or j < len(idxs) - 1:
                    x = F.relu(x)
            return x
        h = seq("block_mlps", posevec, True)
        nb = self.cfg.total_bones - 1
        dR = BodyPoseRefiner._rodrigues(seq("block_mlps_dstR", h, False).view(-1, 3)).view(-1, nb, 3, 3)
        dT = seq("block_mlps_dstT", h, False).view(-1, nb, 3)
        Rn = torch.matmul(dst_Rs[:, 1:].reshape(-1, 3, 3), dR.reshape(-1, 3, 3)).reshape(-1, nb, 3, 3)
        Rs = torch.cat([dst_Rs[:, 0:1], Rn], dim=1)
        Ts = torch.cat([dst_Ts[:, 0:1], dst_Ts[:, 1:] + dT], dim=1)
        return Rs, Ts

    # ------------------------------------------------------------------ one C call per chunk
    def _render_chunk_fused(self, fr, rays_o, rays_d, near, far, t_lin, jitter, bgcolor):
        """``hos_render_human``: samples -> LBS -> Hann PE -> non-rigid MLP -> Fourier PE -> canonical MLP (-> S2 composite) of one
        ray chunk behind one library call, intermediates in a cached workspace.  Same kernels and order as the chain in
        ``forward``; what is saved is the per-kernel Python / allocator work."""
        import ctypes
        from . import _lib
        cfg = self.cfg
        n, S, dev = rays_o.shape[0], cfg.N_samples, rays_o.device
        hc = _lib.HumanConfig()
        nr = None if cfg.ignore_non_rigid_motions else self._fused_nr("nr", self.non_rigid_mlp, fr["cond"])
        cn = self._fused_cnl(fr["state_idx"])
        hc.nr_mlp, hc.cnl_mlp = (None if nr is None else nr._h), cn._h
        hc.n_samples, hc.t_lin, hc.jitter = S, t_lin.data_ptr(), (None if jitter is None else jitter.data_ptr())
        R, T, vol = fr["Rb"][0].contiguous(), fr["Tb"][0].contiguous(), fr["vol"]
        hc.R, hc.T, hc.vol, hc.bones, hc.grid = R.data_ptr(), T.data_ptr(), vol.data_ptr(), R.shape[0], vol.shape[-1]
        bmin = (ctypes.c_float * 3)(*fr["bbox_min"])
        bsc = (ctypes.c_float * 3)(*fr["bbox_scale"])
        hc.bbox_min_host, hc.bbox_scale_host = bmin, bsc
        hc.nr_freqs, hc.hann_w, hc.cnl_freqs, hc.stage2 = self.nr_freqs, fr["hann_w"].data_ptr(), self.cnl_freqs, int(self.stage2)
        hann_host = None
        if FUSE_FOURIER:           # encodings generated inside the MLP kernels (hos_mlp_forward_fourier)
            if "hann_host" not in fr:
                fr["hann_host"] = [float(v) for v in fr["hann_w"].detach().cpu().tolist()]
            hann_host = (ctypes.c_float * len(fr["hann_host"]))(*fr["hann_host"])
            hc.hann_w_host = hann_host
        bg = None
        if bgcolor is not None:
            bg = (ctypes.c_float * 3)(*[float(x) for x in bgcolor.detach().cpu().reshape(-1).tolist()])
            hc.bgcolor_host = bg
        key = (n, S, str(dev), self.nr_freqs, self.cnl_freqs)
        if self._cache.get("hws_key") != key:
            nbytes = ctypes.c_size_t(0)
            _lib.call("hos_render_human_workspace", ctypes.byref(hc), n, ctypes.byref(nbytes))
            self._cache["hws_key"], self._cache["hws"] = key, torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        ws = self._cache["hws"]
        f32 = dict(device=dev, dtype=torch.float32)
        pts, z = torch.empty(n, S, 3, **f32), torch.empty(n, S, **f32)
        ret = {}
        if self.stage2:
            rgb, acc, w, depth = torch.empty(n, 3, **f32), torch.empty(n, **f32), torch.empty(n, S, **f32), torch.empty(n, **f32)
            outs = (rgb.data_ptr(), acc.data_ptr(), w.data_ptr(), depth.data_ptr(), None, None)
            ret.update(rgb=rgb, alpha=acc, depth=depth, weights=w)
        else:
            raw, mask = torch.empty(n, S, 4, **f32), torch.empty(n, S, **f32)
            outs = (None, None, None, None, raw.data_ptr(), mask.data_ptr())
            ret.update(human_rgb=raw[..., :3], human_density=raw[..., 3], newsmpl_pts=pts, pts_mask=mask, z_vals=z, rays_d=rays_d)
        _lib.call("hos_render_human", ctypes.byref(hc), rays_o.data_ptr(), rays_d.data_ptr(), near.data_ptr(), far.data_ptr(), n,
                  ws.data_ptr(), ws.numel(), *outs, pts.data_ptr(), z.data_ptr(), ops._stream())
        if self.stage2:
            _lib.LAUNCHES += 0       # hos_render_human counts its 7 launches (6 without the composite) in _KERNELS_PER_CALL
        ret["deform_pts_final"] = pts[0, 0, :][None, :]
        ret["observe_pts"] = pts[0, 0, :][None, :]
        return ret

    # ------------------------------------------------------------------ training forward (autograd)
    @staticmethod
    def _motion_bases_autograd(Rs, Ts, cnl_gtfms):
        """MotionBasisComputer (network_util.py:106-174) with a graph: the 26-bone chain runs on the host (differentiable
        device <-> host copies), so the pose decoder receives its gradient through the bone maps."""
        dev = Rs.device
        Rs, Ts, cg = Rs.cpu(), Ts.cpu(), cnl_gtfms.detach().cpu()
        nb = Rs.shape[1]
        local = torch.cat([torch.cat([Rs, Ts[..., None]], dim=-1),
                           torch.tensor([0., 0., 0., 1.]).expand(1, nb, 1, 4)], dim=-2)            # [1, nb, 4, 4]
        glob = [local[:, 0]]
        for i in range(1, nb):
            glob.append(torch.matmul(glob[SMPL_PARENT[i]], local[:, i]))
        glob = torch.stack(glob, dim=1).view(-1, 4, 4)
        cgf = cg.view(-1, 4, 4)
        back = torch.matmul(cgf, torch.inverse(glob)).view(-1, nb, 4, 4)
        fwd = torch.matmul(glob, torch.inverse(cgf)).view(-1, nb, 4, 4)
        return (back[0, :, :3, :3].contiguous().to(dev), back[0, :, :3, 3].contiguous().to(dev),
                fwd[0, :, :3, :3].contiguous().to(dev), fwd[0, :, :3, 3].contiguous().to(dev))

    @staticmethod
    def _affine_inverse(M):
        """Inverse of [B, 4, 4] affine maps ([A t; 0 1]) from cross products - no LU, no pivots read on the host."""
        A, t = M[:, :3, :3], M[:, :3, 3]
        a, b, c = A[:, :, 0], A[:, :, 1], A[:, :, 2]
        r0, r1, r2 = torch.linalg.cross(b, c), torch.linalg.cross(c, a), torch.linalg.cross(a, b)
        det = (a * r0).sum(-1)
        Ai = torch.stack([r0, r1, r2], dim=1) / det[:, None, None]
        ti = -torch.matmul(Ai, t[:, :, None])
        return torch.cat([torch.cat([Ai, ti], dim=-1), M[:, 3:4, :]], dim=-2)

    @classmethod
    def _motion_bases_device(cls, Rs, Ts, cnl_gtfms):
        """MotionBasisComputer (network_util.py:106-174) with a graph, entirely on the device and free of host reads (the
        ``static_shapes`` training mode: CUDA-graph capturable).  Same chain as ``_motion_bases_autograd``; the two inverses
        use the closed form for affine maps."""
        nb = Rs.shape[1]
        cg = cnl_gtfms.detach().to(Rs.dtype)
        local = torch.cat([torch.cat([Rs, Ts[..., None]], dim=-1), cg[:, :, 3:4, :]], dim=-2)      # last row [0 0 0 1] of the inputs
        glob = [local[:, 0]]
        for i in range(1, nb):
            glob.append(torch.matmul(glob[SMPL_PARENT[i]], local[:, i]))
        glob = torch.stack(glob, dim=1).view(-1, 4, 4)
        cgf = cg.view(-1, 4, 4)
        back = torch.matmul(cgf, cls._affine_inverse(glob)).view(-1, nb, 4, 4)
        fwd = torch.matmul(glob, cls._affine_inverse(cgf)).view(-1, nb, 4, 4)
        return (back[0, :, :3, :3].contiguous(), back[0, :, :3, 3].contiguous(),
                fwd[0, :, :3, :3].contiguous(), fwd[0, :, :3, 3].contiguous())

    def _train_const(self, key, make):
        """Per-module constants of the training forward that start on the host (Hann window, linspace, bounding box): made
        once per key, so a step that has been run before touches no host memory (and can be captured in a CUDA graph)."""
        d = self._cache.setdefault("train_const", {})
        if key not in d:
            if len(d) > 64:
                d.clear()
            d[key] = make()
        return d[key]

    def _refined_pose(self, Rs, Ts, posevec, it):
        """network.py:590-605 on the device, with a graph."""
        if it < self.cfg.pose_decoder.get("kick_in_iter", 0):
            return Rs, Ts
        out = self.pose_decoder(posevec)
        nb = self.cfg.total_bones - 1
        Rn = torch.matmul(Rs[:, 1:].reshape(-1, 3, 3), out["Rs"].reshape(-1, 3, 3)).reshape(-1, nb, 3, 3)
        return torch.cat([Rs[:, 0:1], Rn], dim=1), torch.cat([Ts[:, 0:1], Ts[:, 1:] + out["Ts"]], dim=1)

    def _forward_train(self, rays, dst_Rs, dst_Ts, cnl_gtfms, motion_weights_priors, dst_posevec, near, far, iter_val, rand,
                       **kwargs):
        """``Network.forward`` under autograd (training, S3 network.py:574-698 / S2): every parameter receives its gradient.
        The per-point work runs on the library's kernels through autograd Functions (``train.LbsWarpFn`` / ``LbsForwardFn``:
        LBS forward + backward kernels; ``train.MlpFn``: tcgen05 layer GEMMs with dgrad / wgrad); encodings, the <= 4-wide
        activations and the S2 composite are small elementwise torch ops; the per-frame prologue (pose decoder, kinematic
        chain, volume decoder) is torch with a graph."""
        from . import train
        cfg = self.cfg
        dev = rays.device
        time = kwargs.get("time", 0.0)
        flow = float(time) > 0.005
        it = float(iter_val.reshape(-1)[0]) if isinstance(iter_val, torch.Tensor) else float(iter_val)
        posevec = dst_posevec[None, ...].float()
        Rs, Ts = self._refined_pose(dst_Rs[None, ...].float(), dst_Ts[None, ...].float(), posevec, it)
        static = bool(getattr(self, "static_shapes", False))       # CUDA-graph capturable variant, see train.GraphedStep
        bases = self._motion_bases_device if static else self._motion_bases_autograd
        Rb, Tb, Rf, Tf = bases(Rs, Ts, cnl_gtfms[None, ...])
        vol = self.mweight_vol_decoder(motion_weights_priors=motion_weights_priors[None, ...])[0]
        kick = cfg.non_rigid_motion_mlp.kick_in_iter
        hann_w = self._train_const(("hann", it, str(dev)), lambda: hann_window_weights(
            self.nr_freqs, it, kick, cfg.non_rigid_motion_mlp.full_band_iter).to(dev))
        cond = torch.zeros_like(posevec) if it < kick else posevec
        state_idx = select_state_index(len(self.human_stateembeds), time, self.transitions_times)
        bmin_t, bscale_t = kwargs["cnl_bbox_min_xyz"], kwargs["cnl_bbox_scale_xyz"]
        bbox_min, bbox_scale, _, _ = self._train_const(
            ("bbox", id(bmin_t), bmin_t._version, id(bscale_t), bscale_t._version),
            lambda: (bmin_t.detach().cpu().reshape(-1).tolist(), bscale_t.detach().cpu().reshape(-1).tolist(), bmin_t, bscale_t))
        rays_o, rays_d = rays
        rays_shape = rays_d.shape
        rays_o = torch.reshape(rays_o, [-1, 3]).float().contiguous()
        rays_d = torch.reshape(rays_d, [-1, 3]).float().contiguous()
        n, S = rays_o.shape[0], cfg.N_samples
        t_lin = self._train_const(("t_lin", S, str(dev)), lambda: torch.linspace(0., 1., steps=S).to(dev))
        jitter = None
        if cfg.perturb > 0.:
            if rand is None and static:
                jitter = torch.rand(n, S, device=dev)           # device generator: its state advances with every graph replay
            else:
                jitter = (torch.rand(n, S) if rand is None else rand).to(dev, torch.float32).contiguous()
        z, pts = ops.human_samples(rays_o, rays_d, near.reshape(-1).float().contiguous(), far.reshape(-1).float().contiguous(),
                                   t_lin, jitter)
        flat = pts.view(-1, 3)

        def hann_pe(x):
            f = 2.0 ** torch.arange(self.nr_freqs, device=dev, dtype=torch.float32)
            ang = x[:, None, :] * f[None, :, None]                                        # [P, F, 3]
            return (torch.stack([torch.sin(ang), torch.cos(ang)], dim=2) * hann_w[None, :, None, None]).reshape(x.shape[0], -1)

        def fourier_pe(x):
            f = 2.0 ** torch.arange(self.cnl_freqs, device=dev, dtype=torch.float32)
            ang = x[:, None, :] * f[None, :, None]
            return torch.cat([x, torch.stack([torch.sin(ang), torch.cos(ang)], dim=2).reshape(x.shape[0], -1)], dim=-1)

        def nr(mlp, x, c):
            return x if cfg.ignore_non_rigid_motions else train.non_rigid_mlp_train(mlp, hann_pe(x), c, x)

        x_skel, mask = train.LbsWarpFn.apply(flat, Rb, Tb, vol, bbox_min, bbox_scale)
        cnl = nr(self.non_rigid_mlp, x_skel, cond)
        raw = train.canonical_mlp_train(self, fourier_pe(cnl), self.human_stateembeds[state_idx]).view(n, S, 4)
        mask2 = mask.view(n, S)
        ret = {}
        sel = mask.detach() > 0.005
        if static:                   # dense cycle side path: every point goes through it, ``cycle_mask`` marks the reference's subset
            xd = train.LbsForwardFn.apply(cnl, Rf, Tf, vol, bbox_min, bbox_scale)
            ret["deform_pts_final"], ret["observe_pts"], ret["cycle_mask"] = nr(self.non_rigid_forward_mlp, xd, cond), flat, sel
        elif bool(sel.any()):        # cycle side path (network.py:505-536)
            xd = train.LbsForwardFn.apply(cnl[sel], Rf, Tf, vol, bbox_min, bbox_scale)
            ret["deform_pts_final"], ret["observe_pts"] = nr(self.non_rigid_forward_mlp, xd, cond), flat[sel]
        else:
            ret["deform_pts_final"] = ret["observe_pts"] = pts[0, 0, :][None, :]
        if flow:                     # previous frame (network.py:474-502, 609-637)
            pv = kwargs["dst_posevec_prev"][None, ...].float()
            Rp, Tp = self._refined_pose(kwargs["dst_Rs_prev"][None, ...].float(), kwargs["dst_Ts_prev"][None, ...].float(), pv, it)
            _, _, Rfp, Tfp = self._motion_bases_autograd(Rp, Tp, cnl_gtfms[None, ...])
            xp = train.LbsForwardFn.apply(cnl, Rfp, Tfp, vol, bbox_min, bbox_scale)
            ret["deform_pts_prev_final"] = nr(self.non_rigid_forward_mlp, xp, torch.zeros_like(pv) if it < kick else pv).view(n, S, 3)
        bgcolor = kwargs.get("bgcolor")
        if self.stage2:              # S2 network.py:273-299
            dists = torch.cat([z[..., 1:] - z[..., :-1], torch.full_like(z[..., :1], 1e10)], dim=-1) * torch.norm(rays_d[..., None, :], dim=-1)
            rgb = torch.sigmoid(raw[..., :3])
            alpha = (1.0 - torch.exp(-F.relu(raw[..., 3]) * dists)) * mask2
            if static:       # exp(cumsum(log)): torch's cumprod backward reads a flag back to the host (not capturable); factors >= 1e-10
                lg = torch.log(1. - alpha + 1e-10)
                T = torch.exp(torch.cumsum(lg, dim=-1) - lg)
            else:
                T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1. - alpha + 1e-10], dim=-1), dim=-1)[:, :-1]
            w = alpha * T
            acc = torch.sum(w, -1)
            ret.update(rgb=torch.sum(w[..., None] * rgb, -2) + (1. - acc[..., None]) * bgcolor.to(dev)[None, :] / 255.,
                       alpha=acc, depth=torch.sum(w * z, -1), weights=w)
        else:
            ret.update(human_rgb=torch.sigmoid(raw[..., :3]), human_density=F.relu(raw[..., 3]), newsmpl_pts=pts, pts_mask=mask2)
            if not flow:
                ret.update(z_vals=z, rays_d=rays_d)
        for k in ret:
            if k not in ("deform_pts_prev_final", "deform_pts_final", "observe_pts", "cycle_mask"):
                ret[k] = torch.reshape(ret[k], list(rays_shape[:-1]) + list(ret[k].shape[1:]))
        ret["bgcolor"] = bgcolor
        return ret

    def forward(self, rays, dst_Rs, dst_Ts, cnl_gtfms, motion_weights_priors, dst_posevec=None,
                near=None, far=None, iter_val=1e7, rand=None, **kwargs):
        if not rays.is_cuda:
            raise RuntimeError("hosnerf_b200.Network: inputs must be CUDA tensors (no CPU fallback)")
        is_train = bool(kwargs.get("is_train", False))
        time = kwargs.get("time", 0.0)
        if is_train and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._forward_train(rays, dst_Rs, dst_Ts, cnl_gtfms, motion_weights_priors, dst_posevec, near, far,
                                       iter_val, rand, **kwargs)
        flow = is_train and float(time) > 0.005           # flow side path: previous-frame forward warp (network.py:474-502)
        precision = self.precision or _m._DEFAULT_PRECISION
        cfg = self.cfg
        with torch.no_grad():
            # ---- per-frame prologue, cached: the reference re-evaluates pose refinement, kinematic chain and the
            # motion-weight volume decoder for every ray chunk of a frame (S3 model.py:745); the inputs are the same
            # tensors for all chunks of a frame, so the results are kept while (tensor identity, in-place version) of
            # every input and the parameter versions of the modules involved are unchanged.  The cache holds the input
            # tensors, so an address cannot be reused by a different tensor while its entry is alive.
            def tag(x):
                return (id(x), x._version) if isinstance(x, torch.Tensor) else ("v", x)
            frame_inputs = [dst_Rs, dst_Ts, cnl_gtfms, motion_weights_priors, dst_posevec, iter_val, time,
                            kwargs["cnl_bbox_min_xyz"], kwargs["cnl_bbox_scale_xyz"]]
            if flow:
                frame_inputs += [kwargs["dst_Rs_prev"], kwargs["dst_Ts_prev"], kwargs["dst_posevec_prev"]]
            fkey = (tuple(tag(x) for x in frame_inputs), str(rays.device),
                    self._versions([self.pose_decoder, self.mweight_vol_decoder]), len(self.human_stateembeds))
            fr = self._cache.get("frame")
            if fr is None or fr["key"] != fkey:
                it = float(iter_val.reshape(-1)[0]) if isinstance(iter_val, torch.Tensor) else float(iter_val)
                # the prologue runs on the host (26 bones of 4x4 algebra), see MotionBasisComputer
                Rs_h, Ts_h = dst_Rs[None, ...].detach().cpu(), dst_Ts[None, ...].detach().cpu()
                posevec = dst_posevec[None, ...]
                if it >= cfg.pose_decoder.get("kick_in_iter", 0):
                    Rs_h, Ts_h = self._correct_pose(Rs_h, Ts_h, posevec.detach().cpu())
                hann_w = hann_window_weights(self.nr_freqs, it, cfg.non_rigid_motion_mlp.kick_in_iter,
                                             cfg.non_rigid_motion_mlp.full_band_iter).to(rays.device)
                cond = torch.zeros_like(posevec) * posevec if it < cfg.non_rigid_motion_mlp.kick_in_iter else posevec
                Rb, Tb, Rf, Tf = self.motion_basis_computer(Rs_h, Ts_h, cnl_gtfms[None, ...])
                prev = None
                if flow:        # previous frame: refined pose -> forward motion bases, its own condition code (network.py:609-637)
                    Rp, Tp = kwargs["dst_Rs_prev"][None, ...].detach().cpu(), kwargs["dst_Ts_prev"][None, ...].detach().cpu()
                    pv = kwargs["dst_posevec_prev"][None, ...]
                    if it >= cfg.pose_decoder.get("kick_in_iter", 0):
                        Rp, Tp = self._correct_pose(Rp, Tp, pv.detach().cpu())
                    _, _, Rfp, Tfp = self.motion_basis_computer(Rp, Tp, cnl_gtfms[None, ...])
                    prev = dict(Rf=Rfp, Tf=Tfp,
                                cond=torch.zeros_like(pv) * pv if it < cfg.non_rigid_motion_mlp.kick_in_iter else pv)
                fr = dict(key=fkey, refs=frame_inputs, it=it, hann_w=hann_w, cond=cond, Rb=Rb, Tb=Tb, Rf=Rf, Tf=Tf, prev=prev,
                          vol=self._volume(motion_weights_priors[None, ...]),
                          state_idx=select_state_index(len(self.human_stateembeds), time, self.transitions_times),
                          bbox_min=kwargs["cnl_bbox_min_xyz"].detach().cpu().reshape(-1).tolist(),
                          bbox_scale=kwargs["cnl_bbox_scale_xyz"].detach().cpu().reshape(-1).tolist(), serial=self._cache.get("frame_serial", 0) + 1)
                self._cache["frame"], self._cache["frame_serial"] = fr, fr["serial"]
            hann_w, cond, Rb, Tb, Rf, Tf, vol, state_idx = (fr["hann_w"], fr["cond"], fr["Rb"], fr["Tb"], fr["Rf"], fr["Tf"],
                                                          fr["vol"], fr["state_idx"])

            rays_o, rays_d = rays
            rays_shape = rays_d.shape
            rays_o = torch.reshape(rays_o, [-1, 3]).float().contiguous()
            rays_d = torch.reshape(rays_d, [-1, 3]).float().contiguous()
            near = near.reshape(-1).float().contiguous()
            far = far.reshape(-1).float().contiguous()
            S = cfg.N_samples
            key = ("tlin", S, str(rays.device))
            if self._cache.get("tlin_key") != key:
                self._cache["tlin_key"], self._cache["tlin"] = key, torch.linspace(0., 1., steps=S).to(rays.device)
            t_lin = self._cache["tlin"]
            bbox_min, bbox_scale = fr["bbox_min"], fr["bbox_scale"]
            bgcolor = kwargs.get("bgcolor")
            n = rays_o.shape[0]
            jitter = None
            if cfg.perturb > 0.:
                # reference: torch.rand(z_vals.shape) on the host per chunk (network.py:421)
                jitter = (torch.rand(n, S) if rand is None else rand).to(rays.device, torch.float32).contiguous()

            outs = {}
            if n == 0:
                raise RuntimeError("hosnerf_b200.Network: empty ray batch (the reference fails on it too: torch.cat of no chunks)")
            one_call = ONE_CALL and precision == "fp16" and not flow and not kwargs.get("cycle_outputs", True)
            for c0 in range(0, n, cfg.chunk):
                c1 = min(n, c0 + cfg.chunk)
                if one_call:
                    ret = self._render_chunk_fused(fr, rays_o[c0:c1], rays_d[c0:c1], near[c0:c1], far[c0:c1], t_lin,
                                                   None if jitter is None else jitter[c0:c1], bgcolor)
                    for k, v in ret.items():
                        outs.setdefault(k, []).append(v)
                    continue
                z, pts = ops.human_samples(rays_o[c0:c1], rays_d[c0:c1], near[c0:c1], far[c0:c1], t_lin,
                                           None if jitter is None else jitter[c0:c1])
                x_skel, mask = ops.lbs_warp(pts, Rb[0], Tb[0], vol, bbox_min, bbox_scale)
                if cfg.ignore_non_rigid_motions:
                    cnl = x_skel
                else:
                    cnl = self._eval_non_rigid("nr", self.non_rigid_mlp, x_skel, cond, hann_w, precision)
                raw = self._eval_canonical(cnl, state_idx, precision).view(c1 - c0, S, 4)
                mask = mask.view(c1 - c0, S)
                ret = {}
                if self.stage2:
                    rgb, acc, w, depth = ops.composite_nerf(raw, mask, z, rays_d[c0:c1], bgcolor, activate=True)
                    ret.update(rgb=rgb, alpha=acc, depth=depth, weights=w)
                else:
                    ret.update(human_rgb=raw[..., :3], human_density=raw[..., 3], newsmpl_pts=pts, pts_mask=mask)
                    if not flow:                        # the reference's train-mode dict drops them (network.py:538-547)
                        ret.update(z_vals=z, rays_d=rays_d[c0:c1])
                if flow:
                    pf = fr["prev"]
                    xp, _ = ops.lbs_forward(cnl.reshape(-1, 3).contiguous(), pf["Rf"][0], pf["Tf"][0], vol, bbox_min, bbox_scale)
                    if not cfg.ignore_non_rigid_motions:
                        xp = self._eval_non_rigid("nrf_prev", self.non_rigid_forward_mlp, xp, pf["cond"], hann_w, precision)
                    ret["deform_pts_prev_final"] = xp.view(c1 - c0, S, 3)
                # cycle-consistency side path (network.py:505-536): forward warp of the points the motion
                # field considers foreground.  Evaluated on every call like the reference does (its outputs are
                # only read by the training loss, so render loops may pass cycle_outputs=False - an extension).
                ret["deform_pts_final"] = pts[0, 0, :][None, :]
                ret["observe_pts"] = pts[0, 0, :][None, :]
                if kwargs.get("cycle_outputs", True):
                    sel = mask.reshape(-1) > 0.005
                    if bool(sel.any()):
                        observe = pts.reshape(-1, 3)[sel].contiguous()
                        xd, _ = ops.lbs_forward(cnl.reshape(-1, 3)[sel].contiguous(), Rf[0], Tf[0], vol, bbox_min, bbox_scale)
                        if not cfg.ignore_non_rigid_motions:
                            xd = self._eval_non_rigid("nrf", self.non_rigid_forward_mlp, xd, cond, hann_w, precision)
                        ret["deform_pts_final"], ret["observe_pts"] = xd, observe
                ret["_x_skel"], ret["_cnl_pts"] = x_skel, cnl          # stage-wise parity hooks (all chunks, like every other output)
                for k, v in ret.items():
                    outs.setdefault(k, []).append(v)
            all_ret = {k: torch.cat(v, 0) for k, v in outs.items()}
            for k in all_ret:
                if k not in ("deform_pts_prev_final", "deform_pts_final", "observe_pts", "_x_skel", "_cnl_pts"):
                    all_ret[k] = torch.reshape(all_ret[k], list(rays_shape[:-1]) + list(all_ret[k].shape[1:]))
            all_ret["bgcolor"] = bgcolor
            return all_ret
