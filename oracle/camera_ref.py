"""ORACLE (test infrastructure, not a product path): numpy restatement of the reference's ray generators,
SURVEY 8f item 2.  Only tests/ may import this.  Pinned to outputs of the unmodified reference by
tests/golden/camera_rays.npz (generator: tests/golden/make_golden_camera.py).

  rays_from_krt / rays_from_krt_bkg  <-  S3 core/utils/camera_util.py:154-216
  rays_intersect_bbox                 <-  S3 core/utils/camera_util.py:219-265
"""
import numpy as np


def rays_from_krt(H, W, K, R, T):
    """camera_util.py:154-181: pixel (i, j) -> K^-1 [i, j, 1] -> world; origin = -R^T T (float64 like numpy's)."""
    K, R, T = np.asarray(K, np.float64), np.asarray(R, np.float64), np.asarray(T, np.float64).ravel()
    origin = -(R.T @ T)
    jj, ii = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    pix = np.stack([ii, jj, np.ones_like(ii)], axis=-1).astype(np.float64)          # [H, W, 3]
    cam = pix @ np.linalg.inv(K).T
    world = (cam - T) @ R
    d = world - origin
    return np.broadcast_to(origin, d.shape).copy(), d


def rays_from_krt_bkg(H, W, K, R, T):
    """camera_util.py:183-216: the same rays plus unit view directions and the mip-NeRF pixel radii
    (distance to the next image row, last row copied from dx[-2:-1], times 2 / sqrt(12))."""
    o, d = rays_from_krt(H, W, K, R, T)
    viewdirs = d / np.linalg.norm(d, axis=-1, keepdims=True)
    dx = np.sqrt(((d[:-1] - d[1:]) ** 2).sum(-1))              # [H - 1, W]
    dx = np.concatenate([dx, dx[-2:-1]], axis=0)               # note: row H - 3, as in the reference
    return o, d, viewdirs, dx[..., None] * 2 / np.sqrt(12)


def rays_intersect_bbox(bounds, ray_o, ray_d):
    """camera_util.py:219-265.  Returns (near[N_valid], far[N_valid], mask[N], clamped ray_d) - the reference writes the
    |d| < 1e-5 clamp back into the caller's array, here it is returned."""
    if isinstance(bounds, dict):
        bounds = np.stack([bounds["min_xyz"], bounds["max_xyz"]], axis=0)
    b = np.asarray(bounds, np.float64) + np.array([-0.01, 0.01])[:, None]
    d = np.array(ray_d, copy=True)
    d[np.abs(d) < 1e-5] = 1e-5
    o = np.asarray(ray_o)
    t = ((b[None] - o[:, None]) / d[:, None]).reshape(-1, 6)                    # plane order: min xyz, max xyz
    p = t[..., None] * d[:, None] + o[:, None]                                  # [N, 6, 3]
    eps = 1e-6
    lo, hi = b[0] - eps, b[1] + eps
    inside = np.all((p >= lo) & (p <= hi), axis=-1)                             # [N, 6]
    mask = inside.sum(-1) == 2
    hits = p[mask][inside[mask]].reshape(-1, 2, 3)
    om, dm = o[mask], d[mask]
    norm = np.linalg.norm(dm, axis=1)                                           # in the rays' own precision
    d0 = np.linalg.norm(hits[:, 0] - om, axis=1) / norm
    d1 = np.linalg.norm(hits[:, 1] - om, axis=1) / norm
    return np.minimum(d0, d1), np.maximum(d0, d1), mask, d
