"""GPU parity of the on-device ray generators (hosnerf_b200/camera.py -> hos_rays_from_krt / hos_rays_intersect_bbox)
against the reference's own outputs (tests/golden/camera_rays.npz) and, at full frame size, against the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from hosnerf_b200 import camera  # noqa: E402
from oracle import camera_ref as C  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "camera_rays.npz"))


def test_rays_from_krt_golden():
    H, W = int(G["H"]), int(G["W"])
    o, d = camera.get_rays_from_KRT(H, W, G["K"], G["R"], G["T"], dtype=torch.float64)
    assert np.allclose(o.cpu().numpy(), G["krt_rays_o"], rtol=0, atol=1e-12)
    assert np.allclose(d.cpu().numpy(), G["krt_rays_d"], rtol=0, atol=1e-12)
    o32, d32, v32, r32 = camera.get_rays_from_KRT_bkg(H, W, G["K"], G["R"], G["T"])           # float32, what the data loaders keep
    assert o32.dtype == torch.float32 and r32.shape == (H, W, 1)
    assert np.array_equal(d32.cpu().numpy(), G["bkg_rays_d"].astype(np.float32))              # correctly rounded from float64
    assert np.allclose(v32.cpu().numpy(), G["bkg_viewdirs"].astype(np.float32), rtol=0, atol=1e-7)
    assert np.allclose(r32.cpu().numpy(), G["bkg_radii"].astype(np.float32), rtol=2e-7, atol=0)


def test_rays_from_krt_bkg_full_frame_vs_oracle():
    H, W = 1080, 1920                                     # C5: 2,073,600 rays
    K = np.array([[1500.0, 0.0, 960.0], [0.0, 1500.0, 540.0], [0.0, 0.0, 1.0]])
    th = 0.4
    R = np.array([[np.cos(th), 0.0, np.sin(th)], [0.0, 1.0, 0.0], [-np.sin(th), 0.0, np.cos(th)]])
    T = np.array([0.3, -0.1, 4.0])
    o, d, v, r = camera.get_rays_from_KRT_bkg(H, W, K, R, T, dtype=torch.float64)
    ro, rd, rv, rr = C.rays_from_krt_bkg(H, W, K, R, T)
    assert np.allclose(o.cpu().numpy(), ro, rtol=0, atol=1e-12)
    assert np.allclose(d.cpu().numpy(), rd, rtol=0, atol=1e-9)
    assert np.allclose(v.cpu().numpy(), rv, rtol=0, atol=1e-12)
    assert np.allclose(r.cpu().numpy(), rr, rtol=1e-9, atol=0)
    assert torch.allclose(v.norm(dim=-1), torch.ones(H, W, device=v.device, dtype=v.dtype), atol=1e-12)
    assert torch.equal(r[-1], r[-3])                      # the reference's dx[-2:-1] quirk


def test_rays_intersect_bbox_golden():
    o = torch.from_numpy(G["box_rays_o"]).cuda()
    d = torch.from_numpy(G["box_rays_d"]).cuda()
    near, far, mask = camera.rays_intersect_3d_bbox(G["box_bounds"], o, d)
    assert np.array_equal(mask.cpu().numpy(), G["box_mask"])
    assert np.allclose(near.cpu().numpy(), G["box_near"], rtol=2e-7, atol=1e-7)
    assert np.allclose(far.cpu().numpy(), G["box_far"], rtol=2e-7, atol=1e-7)
    assert (d.abs() >= 1e-5).all()                        # clamp written back, as the reference does to its argument
    assert (near <= far).all()


def test_rays_intersect_bbox_large_vs_oracle_and_edges():
    rng = np.random.default_rng(3)
    n = 300_000
    o = rng.uniform(-3, 3, size=(n, 3))
    d = rng.normal(size=(n, 3))
    bounds = {"min_xyz": np.array([-1.0, -0.5, -0.7]), "max_xyz": np.array([0.8, 0.9, 0.6])}
    near, far, mask = camera.rays_intersect_3d_bbox(bounds, torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda())
    rn, rf, rm, _ = C.rays_intersect_bbox(bounds, o, d)
    assert np.array_equal(mask.cpu().numpy(), rm)
    assert np.allclose(near.cpu().numpy(), rn, rtol=2e-7, atol=1e-7) and np.allclose(far.cpu().numpy(), rf, rtol=2e-7, atol=1e-7)
    # empty input
    e = torch.empty(0, 3, device="cuda")
    n0, f0, m0 = camera.rays_intersect_3d_bbox(bounds, e, e.clone())
    assert n0.numel() == 0 and m0.numel() == 0
    with pytest.raises(RuntimeError, match="CUDA"):
        camera.rays_intersect_3d_bbox(bounds, torch.zeros(4, 3), torch.ones(4, 3))


def test_batchified_get_rays_golden():
    """Stage-1 dataset ray generator (S1 src/data/ray_utils.py:34-139) against the reference's own output on three small cameras:
    origins exact, directions / radii within 2 float32 ulps (NumPy-version dependent promotion in the reference, see the
    generator), rays_d returned normalised like the reference's aliased array."""
    import numpy as np
    from hosnerf_b200 import camera
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "batchified_rays.npz"))
    sizes = [tuple(int(v) for v in s) for s in z["sizes"]]
    ro, rd, vd, radii, ml = camera.batchified_get_rays(list(z["intr"]), list(z["extr"]), sizes, True, True, False, None, [1.0, 0.5, 2.0])
    assert torch.equal(ro.cpu(), torch.from_numpy(z["rays_o"]))
    assert rd.data_ptr() == vd.data_ptr()
    assert float((rd.cpu() - torch.from_numpy(z["rays_d"])).abs().max()) < 3e-7
    assert float(((radii.cpu() - torch.from_numpy(z["radii"]).float()).abs() / torch.from_numpy(z["radii"]).float()).max()) < 2e-5
    assert torch.equal(ml.cpu(), torch.from_numpy(z["multloss"]).float())
