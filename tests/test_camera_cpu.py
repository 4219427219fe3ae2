"""Oracle of the ray generators (oracle/camera_ref.py) against the reference's own outputs (tests/golden/camera_rays.npz)."""
import os

import numpy as np

from oracle import camera_ref as C

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "camera_rays.npz"))


def test_rays_from_krt_matches_reference():
    o, d = C.rays_from_krt(int(G["H"]), int(G["W"]), G["K"], G["R"], G["T"])
    assert np.allclose(o, G["krt_rays_o"], rtol=0, atol=1e-13)
    assert np.allclose(d, G["krt_rays_d"], rtol=0, atol=1e-13)


def test_rays_from_krt_bkg_matches_reference():
    o, d, v, r = C.rays_from_krt_bkg(int(G["H"]), int(G["W"]), G["K"], G["R"], G["T"])
    assert np.allclose(d, G["bkg_rays_d"], rtol=0, atol=1e-13)
    assert np.allclose(v, G["bkg_viewdirs"], rtol=0, atol=1e-13)
    assert np.allclose(r, G["bkg_radii"], rtol=1e-12, atol=0)
    assert np.array_equal(r[-1], r[-3])          # the reference's dx[-2:-1] quirk


def test_rays_intersect_bbox_matches_reference():
    near, far, mask, d = C.rays_intersect_bbox(G["box_bounds"], G["box_rays_o"], G["box_rays_d"])
    assert np.array_equal(mask, G["box_mask"])
    assert near.shape == G["box_near"].shape
    assert np.allclose(near, G["box_near"], rtol=1e-12, atol=0)
    assert np.allclose(far, G["box_far"], rtol=1e-12, atol=0)
    assert mask.sum() > 100 and (~mask).sum() > 100
    assert (np.abs(d) >= np.float32(1e-5)).all()
