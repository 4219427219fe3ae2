"""Drop-in replacements for the reference's per-frame ray generators (S3 core/utils/camera_util.py:154-265):
same names, argument order and return tuples, but the rays are produced by CUDA kernels and returned as torch
CUDA tensors instead of numpy arrays (2 M rays per 1080p frame, numpy on the host in the reference)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, ops

_D = C.c_double


def _host(a, n):
    v = np.asarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, dtype=np.float64).reshape(-1)
    assert v.size == n, (v.size, n)
    return (_D * n)(*v.tolist())


def _rays(H, W, K, R, T, bkg, dtype, device):
    if dtype not in (torch.float32, torch.float64):
        raise ValueError("dtype must be torch.float32 or torch.float64")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("hosnerf_b200.camera: rays are generated on a CUDA device (this package has no CPU path)")
    kinv = np.linalg.inv(np.asarray(K.detach().cpu().numpy() if isinstance(K, torch.Tensor) else K, dtype=np.float64))
    o = torch.empty(H, W, 3, device=dev, dtype=dtype)
    d = torch.empty(H, W, 3, device=dev, dtype=dtype)
    v = torch.empty(H, W, 3, device=dev, dtype=dtype) if bkg else None
    r = torch.empty(H, W, 1, device=dev, dtype=dtype) if bkg else None
    with torch.cuda.device(dev):
        _lib.call("hos_rays_from_krt", int(H), int(W), _host(kinv, 9), _host(R, 9), _host(T, 3), int(bkg),
                  int(dtype == torch.float64), o.data_ptr(), d.data_ptr(), None if v is None else v.data_ptr(),
                  None if r is None else r.data_ptr(), ops._stream())
    return (o, d, v, r) if bkg else (o, d)


def get_rays_from_KRT(H, W, K, R, T, dtype=torch.float32, device="cuda"):
    """camera_util.py:154-181 -> (rays_o [H, W, 3], rays_d [H, W, 3])."""
    return _rays(H, W, K, R, T, False, dtype, device)


def get_rays_from_KRT_bkg(H, W, K, R, T, dtype=torch.float32, device="cuda"):
    """camera_util.py:183-216 -> (rays_o, rays_d, viewdirs [H, W, 3], radii [H, W, 1])."""
    return _rays(H, W, K, R, T, True, dtype, device)


def rays_intersect_3d_bbox(bounds, ray_o, ray_d):
    """camera_util.py:219-265 -> (near [N_valid], far [N_valid], mask_at_box [N]); like the reference, components of
    ``ray_d`` below 1e-5 in magnitude are overwritten with 1e-5 in place."""
    if isinstance(bounds, dict):
        bounds = np.stack([np.asarray(bounds["min_xyz"]), np.asarray(bounds["max_xyz"])], axis=0)
    b = np.asarray(bounds.detach().cpu().numpy() if isinstance(bounds, torch.Tensor) else bounds, dtype=np.float64)
    assert b.shape == (2, 3)
    for t, nm in ((ray_o, "ray_o"), (ray_d, "ray_d")):
        if not isinstance(t, torch.Tensor) or not t.is_cuda or not t.is_contiguous() or t.dtype not in (torch.float32, torch.float64):
            raise RuntimeError(f"hosnerf_b200.camera: `{nm}` must be a contiguous float32/float64 CUDA tensor")
    assert ray_o.dtype == ray_d.dtype and ray_o.shape == ray_d.shape and ray_o.shape[-1] == 3
    n = ray_o.numel() // 3
    dev = ray_o.device
    near = torch.empty(n, device=dev, dtype=torch.float32)
    far = torch.empty(n, device=dev, dtype=torch.float32)
    mask = torch.empty(n, device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        _lib.call_unless_empty(n, "hos_rays_intersect_bbox", _host(b[0], 3), _host(b[1], 3), ray_o.data_ptr(), ray_d.data_ptr(), n,
                               int(ray_o.dtype == torch.float64), near.data_ptr(), far.data_ptr(), mask.data_ptr(),
                               ops._stream())
    m = mask.bool()
    return near[m], far[m], m


def batchified_get_rays(intrinsics, extrinsics, image_sizes, use_pixel_centers, get_radii, ndc_coord=False, ndc_coeffs=None,
                        multlosses=None, device="cuda"):
    """Stage-1 dataset ray generator (S1 src/data/ray_utils.py:34-139), non-NDC branch (the 360 scenes): for every image
    pixel-centre directions through the intrinsics, rotated by the camera-to-world extrinsic, origins = camera centres,
    ``viewdirs`` = normalised directions - and, like the reference (``viewdirs = rays_d; viewdirs /= norm`` aliases the array),
    ``rays_d`` comes back normalised too - radii from the distance between vertically neighbouring un-normalised directions
    (x 2 / sqrt(12)), per-image loss multipliers expanded per ray.  Runs once per dataset, as float32 torch ops on the device
    (the reference pins NumPy 1.23, whose value-based casting keeps this arithmetic in float32); returns CUDA tensors
    (rays_o, rays_d, viewdirs [N,3], radii [N,1] or None, multloss_expand [N,1] or None)."""
    if ndc_coord:
        raise NotImplementedError("hosnerf_b200.camera.batchified_get_rays: the NDC branch (forward-facing scenes) is not built")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("hosnerf_b200.camera: rays are generated on a CUDA device (this package has no CPU path)")
    center = 0.5 if use_pixel_centers else 0.0
    f32 = dict(device=dev, dtype=torch.float32)
    ros, rds, radii, mults = [], [], [], []
    for idx, ((h, w), K, E) in enumerate(zip(image_sizes, intrinsics, extrinsics)):
        K = torch.as_tensor(np.asarray(K), **f32)
        E = torch.as_tensor(np.asarray(E), **f32)
        i = (torch.arange(w, **f32) + center)[None, :].expand(h, w)
        j = (torch.arange(h, **f32) + center)[:, None].expand(h, w)
        dirs = torch.stack([(i - K[0, 2]) / K[0, 0], (j - K[1, 2]) / K[1, 1], torch.ones_like(i)], -1)       # [h, w, 3]
        d = torch.einsum("hwc,rc->hwr", dirs, E[:3, :3])
        ros.append(E[:3, 3].expand(h * w, 3))
        if get_radii:
            dx = torch.sqrt(torch.sum((d[:-1] - d[1:]) ** 2, -1))
            dx = torch.cat([dx, dx[-2:-1]], 0)
            radii.append((dx[..., None] * 2 / np.sqrt(12)).reshape(-1))
        rds.append(d.reshape(-1, 3))
        if multlosses is not None:
            mults.append(torch.full((h * w,), float(multlosses[idx]), **f32))
    rays_o = torch.cat(ros).contiguous()
    rays_d = torch.cat(rds)
    rays_d = (rays_d / torch.linalg.norm(rays_d, dim=-1, keepdim=True)).contiguous()
    return (rays_o, rays_d, rays_d, torch.cat(radii)[:, None] if get_radii else None,
            torch.cat(mults)[:, None] if multlosses is not None else None)
