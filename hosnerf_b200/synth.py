"""Synthetic inputs and deterministic weights for the HOSNeRF per-ray hot path.

Nothing here is on the product path: these are the seeded generators that the
golden-vector script (run against the real reference), the parity tests, the
smoke test and ``bench.py`` all share, so that every party sees the *same*
rays, skeleton and network weights.

Weights are filled *by parameter name* (``fill_params_``) instead of relying on
constructor RNG order, so the reference modules and the drop-in modules of this
package receive bit-identical parameters as long as their ``state_dict`` keys
and shapes agree - which is itself part of the drop-in contract
("checkpoint / resume").
"""
from __future__ import annotations

import math
import zlib

import numpy as np
import torch

# Kinematic tree of the 26-joint skeleton (24 SMPL joints + 2 object joints);
# same topology as the reference (S3/core/utils/body_util.py:43-46).
PARENT = {
    1: 0, 2: 0, 3: 0, 4: 1, 5: 2, 6: 3, 7: 4, 8: 5, 9: 6, 10: 7,
    11: 8, 12: 9, 13: 9, 14: 9, 15: 12, 16: 13, 17: 14, 18: 16, 19: 17,
    20: 18, 21: 19, 22: 20, 23: 21, 24: 23, 25: 22,
}
TOTAL_BONES = 26


# --------------------------------------------------------------------------
# deterministic parameters
# --------------------------------------------------------------------------
def _name_seed(name: str, seed: int) -> int:
    return (zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF


# Output layers that the reference initialises to ~0 (offsets / pose corrections start at
# identity, mlp_offset.py:45-50, mlp_delta_body_pose.py:52-60) and that stay small after
# training: a uniformly Kaiming-scaled fill would make the synthetic human warp metre-sized
# instead of centimetre-sized, which no trained model does.
SMALL_OUTPUT_LAYERS = {
    "non_rigid_mlp.block_mlps.12.": 0.02,
    "non_rigid_forward_mlp.block_mlps.12.": 0.02,
    "pose_decoder.block_mlps_dstR.2.": 0.1,
    "pose_decoder.block_mlps_dstT.2.": 0.02,
}


@torch.no_grad()
def fill_params_(module: torch.nn.Module, seed: int = 0, skip_prefixes=()) -> None:
    """Overwrite every parameter of ``module`` with values that depend only on
    (parameter name, shape, seed).

    * >=2-D weights: U(-b, b) with b = sqrt(6 / fan_in)  (Kaiming-uniform scale,
      keeps activations O(1) through the ReLU stacks)
    * 1-D tensors named ``*bias``: U(-0.05, 0.05)
    * other 1-D tensors (state embeddings, const embeddings): N(0, 1)
    """
    for name, p in module.named_parameters():
        if any(name.startswith(pre) for pre in skip_prefixes):
            continue
        g = torch.Generator().manual_seed(_name_seed(name, seed))
        if p.dim() >= 2:
            fan_in = p.shape[1] * int(np.prod(p.shape[2:])) if p.dim() > 2 else p.shape[1]
            b = math.sqrt(6.0 / fan_in)
            v = (torch.rand(p.shape, generator=g) * 2 - 1) * b
        elif name.endswith("bias"):
            v = (torch.rand(p.shape, generator=g) * 2 - 1) * 0.05
        else:
            v = torch.randn(p.shape, generator=g)
        for key, sc in SMALL_OUTPUT_LAYERS.items():
            if key in name + ".":
                v = v * sc
        p.copy_(v.to(p.dtype))


@torch.no_grad()
def boost_human_density_(net, amount: float = 3.0) -> None:
    """``fill_params_`` leaves the canonical MLP's sigma logit mostly negative (relu -> 0 density
    everywhere); shift its bias so the human branch has real opacity to composite."""
    net.cnl_mlp.output_linear[0].bias[3] += amount


# --------------------------------------------------------------------------
# background (mip-NeRF-360 branch) rays  -- SURVEY 8d, C1/C2
# --------------------------------------------------------------------------
def make_bkg_batch(n_rays: int, seed: int = 1, time: float = 0.0, s3_times: bool = False):
    """rays_o ~ 0.3 N(0,I), rays_d ~ N(0,I) (un-normalised), viewdirs = d/|d|,
    radii = 1e-3.  ``times`` is [N] (stage 1 reads ``times[0:1]``,
    S1/src/model/mipnerf360/model.py:335) or a 0-d tensor (stage 3, S3 :422)."""
    g = torch.Generator().manual_seed(seed)
    rays_o = 0.3 * torch.randn(n_rays, 3, generator=g)
    rays_d = torch.randn(n_rays, 3, generator=g)
    viewdirs = rays_d / rays_d.norm(dim=-1, keepdim=True)
    radii = torch.full((n_rays, 1), 1e-3)
    times = torch.tensor(time) if s3_times else torch.full((n_rays,), time)
    return {"rays_o": rays_o, "rays_d": rays_d, "viewdirs": viewdirs,
            "radii": radii, "times": times}


# --------------------------------------------------------------------------
# human branch inputs -- SURVEY 8d, C3
# --------------------------------------------------------------------------
def _rodrigues(rvec: np.ndarray) -> np.ndarray:
    theta = float(np.linalg.norm(rvec))
    r = (rvec / (theta + 1e-5)).reshape(3, 1)
    K = np.array([[0, -r[2, 0], r[1, 0]], [r[2, 0], 0, -r[0, 0]], [-r[1, 0], r[0, 0], 0]])
    return math.cos(theta) * np.eye(3) + math.sin(theta) * K + (1 - math.cos(theta)) * (r @ r.T)


def _rt(R, t):
    G = np.eye(4, dtype=np.float32)
    G[:3, :3] = R
    G[:3, 3] = t
    return G


def make_skeleton(seed: int = 0, pose_scale: float = 0.2, grid: int = 32, bbox_offset: float = 0.6):
    """Random 26-joint tree + pose -> the per-frame tensors ``Network.forward``
    consumes (S3/core/nets/human_nerf/network.py:574-581 and the kwargs listed
    in SURVEY 8b): dst_Rs, dst_Ts, cnl_gtfms, motion_weights_priors, dst_posevec,
    cnl_bbox_min_xyz, cnl_bbox_scale_xyz.  Everything float32 torch on CPU."""
    rs = np.random.RandomState(seed)
    joints = np.zeros((TOTAL_BONES, 3), np.float32)
    for i in range(1, TOTAL_BONES):
        joints[i] = joints[PARENT[i]] + rs.uniform(-0.15, 0.15, 3)
    poses = (pose_scale * rs.randn(TOTAL_BONES * 3)).astype(np.float32)

    Rs = np.zeros((TOTAL_BONES, 3, 3), np.float32)
    Ts = np.zeros((TOTAL_BONES, 3), np.float32)
    ang = poses.reshape(-1, 3)
    Rs[0] = _rodrigues(ang[0])
    Ts[0] = joints[0]
    for i in range(1, TOTAL_BONES):
        Rs[i] = _rodrigues(ang[i])
        Ts[i] = joints[i] - joints[PARENT[i]]

    gt = np.zeros((TOTAL_BONES, 4, 4), np.float32)
    gt[0] = _rt(np.eye(3), joints[0])
    for i in range(1, TOTAL_BONES):
        gt[i] = gt[PARENT[i]] @ _rt(np.eye(3), joints[i] - joints[PARENT[i]])

    bmin = joints.min(0) - bbox_offset
    bmax = joints.max(0) + bbox_offset

    # Gaussian blobs per joint (anisotropy is irrelevant for a synthetic prior;
    # what matters is a normalised 27 x G^3 volume with a background channel).
    lin = [np.linspace(bmin[a], bmax[a], grid, dtype=np.float32) for a in range(3)]
    zz, yy, xx = np.meshgrid(lin[2], lin[1], lin[0], indexing="ij")
    vols = []
    for j in range(TOTAL_BONES):
        c = joints[j] if j == 0 else 0.5 * (joints[j] + joints[PARENT[j]])
        d2 = (xx - c[0]) ** 2 + (yy - c[1]) ** 2 + (zz - c[2]) ** 2
        vols.append(np.exp(-d2 / (2 * 0.08 ** 2)).astype(np.float32))
    vols = np.stack(vols, 0)
    bg = 1.0 - np.clip(vols.sum(0, keepdims=True), 0.0, 1.0)
    vols = np.concatenate([vols, bg], 0)
    vols = vols / np.clip(vols.sum(0, keepdims=True), 1e-3, None)
    vols = np.clip(vols, 1e-12, None).astype(np.float32)  # log() is taken downstream

    return {
        "dst_Rs": torch.from_numpy(Rs),
        "dst_Ts": torch.from_numpy(Ts),
        "cnl_gtfms": torch.from_numpy(gt),
        "motion_weights_priors": torch.from_numpy(vols),
        "dst_posevec": torch.from_numpy(poses[3:] + 1e-2),
        "cnl_bbox_min_xyz": torch.from_numpy(bmin.astype(np.float32)),
        "cnl_bbox_scale_xyz": torch.from_numpy((2.0 / (bmax - bmin)).astype(np.float32)),
        "joints": torch.from_numpy(joints),
    }


def make_human_rays(n_rays: int, seed: int = 2):
    """o = (0,0,-3), d = (U(-.3,.3), U(-.3,.3), 1), near 2, far 4."""
    g = torch.Generator().manual_seed(seed)
    rays_o = torch.tensor([0.0, 0.0, -3.0]).repeat(n_rays, 1)
    xy = (torch.rand(n_rays, 2, generator=g) * 2 - 1) * 0.3
    rays_d = torch.cat([xy, torch.ones(n_rays, 1)], -1)
    near = torch.full((n_rays, 1), 2.0)
    far = torch.full((n_rays, 1), 4.0)
    return torch.stack([rays_o, rays_d], 0), near, far


def make_human_batch(n_rays: int, seed: int = 0, ray_seed: int = 2, time: float = 0.0,
                     is_train: bool = False, iter_val: float = 1e7):
    sk = make_skeleton(seed)
    rays, near, far = make_human_rays(n_rays, ray_seed)
    batch = {
        "rays": rays, "near": near, "far": far,
        "dst_Rs": sk["dst_Rs"], "dst_Ts": sk["dst_Ts"], "cnl_gtfms": sk["cnl_gtfms"],
        "motion_weights_priors": sk["motion_weights_priors"], "dst_posevec": sk["dst_posevec"],
        "cnl_bbox_min_xyz": sk["cnl_bbox_min_xyz"], "cnl_bbox_scale_xyz": sk["cnl_bbox_scale_xyz"],
        "bgcolor": torch.tensor([255.0, 255.0, 255.0]),
        "iter_val": torch.full((1,), iter_val),
        "time": torch.tensor(time), "is_train": is_train,
    }
    if is_train and time > 0.005:
        prev = make_skeleton(seed + 100)
        batch["dst_Rs_prev"] = prev["dst_Rs"]
        batch["dst_Ts_prev"] = prev["dst_Ts"]
        batch["dst_posevec_prev"] = prev["dst_posevec"]
    return batch


def random_rigid(seed: int = 3, scale: float = 1.3) -> torch.Tensor:
    """newsmpl_to_scale_world: random similarity 4x4 (S3/.../model.py:1524)."""
    rs = np.random.RandomState(seed)
    R = _rodrigues(rs.randn(3) * 0.4)
    G = np.eye(4, dtype=np.float32)
    G[:3, :3] = scale * R
    G[:3, 3] = rs.uniform(-0.2, 0.2, 3)
    return torch.from_numpy(G)
