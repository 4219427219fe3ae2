"""Generate tests/golden/batchified_rays.npz with the UNMODIFIED reference `batchified_get_rays`
(1st_State-Conditional_Scene/src/data/ray_utils.py:34-139), non-NDC branch.  Authoring container only.
NumPy here is 2.x (float32 array op float64 scalar promotes to float64 before the final astype(float32)); the reference pins
1.23.5 (stays float32): the two differ by at most an ulp, which is the tolerance of the test."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_ray_utils", "/root/reference/1st_State-Conditional_Scene/src/data/ray_utils.py")
ru = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ru)
rng = np.random.default_rng(3)
sizes = [(5, 7), (4, 6), (6, 3)]
intr, extr = [], []
for h, w in sizes:
    K = np.array([[300.0 + rng.uniform(-20, 20), 0, w / 2 + rng.uniform(-0.5, 0.5)], [0, 310.0 + rng.uniform(-20, 20), h / 2 + rng.uniform(-0.5, 0.5)], [0, 0, 1]])
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    E = np.eye(4)
    E[:3, :3] = q
    E[:3, 3] = rng.normal(size=3)
    intr.append(K)
    extr.append(E)
ro, rd, vd, radii, ml = ru.batchified_get_rays(intr, extr, sizes, True, True, False, None, [1.0, 0.5, 2.0])
np.savez_compressed(os.path.join(HERE, "batchified_rays.npz"), sizes=np.array(sizes), intr=np.stack(intr), extr=np.stack(extr),
                    rays_o=ro, rays_d=rd, viewdirs=vd, radii=radii, multloss=ml)
print(ro.shape, rd.shape, radii.shape, ml.shape, ro.dtype, radii.dtype)
