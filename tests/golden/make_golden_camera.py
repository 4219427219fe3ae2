"""Generate tests/golden/camera_rays.npz by running the UNMODIFIED reference ray generators
(S3 core/utils/camera_util.py:154-265).  Authoring container only (needs /root/reference):

    python tests/golden/make_golden_camera.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/3rd_Complete_HOSNeRF/core/utils/camera_util.py"
if "cv2" not in sys.modules:
    try:
        import cv2  # noqa: F401
    except Exception:                       # only used by functions outside this path
        sys.modules["cv2"] = types.ModuleType("cv2")
spec = importlib.util.spec_from_file_location("ref_camera_util", REF)
cu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cu)

rng = np.random.default_rng(7)


def rodrigues(v):
    th = np.linalg.norm(v)
    k = v / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


H, W = 37, 53
K = np.array([[61.5, 0.0, 26.1], [0.0, 60.2, 18.7], [0.0, 0.0, 1.0]])
R = rodrigues(np.array([0.3, -0.5, 0.2]))
T = np.array([0.1, -0.2, 3.1])
out = {"H": H, "W": W, "K": K, "R": R, "T": T}
ro, rd = cu.get_rays_from_KRT(H, W, K, R, T)
out["krt_rays_o"], out["krt_rays_d"] = np.ascontiguousarray(ro), rd
ro, rd, vd, rad = cu.get_rays_from_KRT_bkg(H, W, K, R, T)
out["bkg_rays_o"], out["bkg_rays_d"], out["bkg_viewdirs"], out["bkg_radii"] = np.ascontiguousarray(ro), rd, vd, rad

# bbox intersection: rays from the camera above against a box around the origin, plus axis-parallel and grazing rays
bounds = np.array([[-0.6, -0.9, -0.5], [0.7, 0.8, 0.45]])
o = np.ascontiguousarray(out["krt_rays_o"].reshape(-1, 3)).astype(np.float32)
d = out["krt_rays_d"].reshape(-1, 3).astype(np.float32)
extra_o = rng.uniform(-2, 2, size=(400, 3)).astype(np.float32)
extra_d = rng.normal(size=(400, 3)).astype(np.float32)
extra_d[:40, 0] = 0.0          # exercises the |d| < 1e-5 clamp
extra_d[40:80, 1] = 1e-6
extra_o[80:120] = rng.uniform(-0.4, 0.4, size=(40, 3)).astype(np.float32)      # origins inside the box
o = np.concatenate([o, extra_o], 0)
d = np.concatenate([d, extra_d], 0)
out["box_bounds"], out["box_rays_o"], out["box_rays_d"] = bounds, o.copy(), d.copy()
near, far, mask = cu.rays_intersect_3d_bbox(bounds, o.copy(), d.copy())       # the reference clamps ray_d in place
out["box_near"], out["box_far"], out["box_mask"] = near, far, mask
np.savez_compressed(os.path.join(HERE, "camera_rays.npz"), **out)
print({k: (v.shape, v.dtype) if isinstance(v, np.ndarray) else v for k, v in out.items()})
