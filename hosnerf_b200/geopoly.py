"""Geodesic-polyhedron basis for the integrated positional encoding.

Produces the [3, 21] ``pos_basis_t`` buffer of the reference MLPs
(S1/src/model/mipnerf360/helper.py:420-494, ``generate_basis("icosahedron", 2)``).
The column ORDER is part of the checkpoint contract (layer-0 weight columns are
tied to it), so the construction walks the faces / barycentric lattice in the
same order as the reference; tests/test_oracle_golden.py compares the result
bit-for-bit with the buffer of the real reference module.
"""
from __future__ import annotations

import numpy as np
import torch

_PHI = (np.sqrt(5.0) + 1.0) / 2.0

_ICO_VERTS = np.array(
    [(-1, 0, _PHI), (1, 0, _PHI), (-1, 0, -_PHI), (1, 0, -_PHI),
     (0, _PHI, 1), (0, _PHI, -1), (0, -_PHI, 1), (0, -_PHI, -1),
     (_PHI, 1, 0), (-_PHI, 1, 0), (_PHI, -1, 0), (-_PHI, -1, 0)]) / np.sqrt(_PHI + 2.0)

_ICO_FACES = np.array(
    [(0, 4, 1), (0, 9, 4), (9, 5, 4), (4, 5, 8), (4, 8, 1), (8, 10, 1), (8, 3, 10), (5, 3, 8), (5, 2, 3),
     (2, 7, 3), (7, 10, 3), (7, 6, 10), (7, 11, 6), (11, 0, 6), (0, 1, 6), (6, 1, 10), (9, 0, 11),
     (9, 11, 2), (9, 2, 5), (7, 2, 11)])

_OCT_VERTS = np.array([(0, 0, -1), (0, 0, 1), (0, -1, 0), (0, 1, 0), (-1, 0, 0), (1, 0, 0)], dtype=np.float64)


def _pairwise_sq(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Squared distances between rows of a and rows of b (Gram-matrix form, clamped at 0)."""
    na = np.sum(a.T ** 2, 0)
    nb = np.sum(b.T ** 2, 0)
    return np.maximum(0, na[:, None] + nb[None, :] - 2 * a @ b.T)


def _tessellate(verts: np.ndarray, faces: np.ndarray, v: int, tol: float) -> np.ndarray:
    lattice = np.array([(i, j, v - i - j) for i in range(v + 1) for j in range(v + 1 - i)]) / v
    pts = []
    for f in faces:
        p = np.matmul(lattice, verts[f, :])
        p /= np.sqrt(np.sum(p ** 2, 1, keepdims=True))
        pts.append(p)
    pts = np.concatenate(pts, 0)
    d2 = _pairwise_sq(pts, pts)
    first_seen = np.array([np.min(np.argwhere(row <= tol)) for row in d2])
    return pts[np.unique(first_seen), :]


def generate_basis(base_shape: str = "icosahedron", angular_tesselation: int = 2,
                   remove_symmetries: bool = True, eps: float = 1e-4) -> torch.Tensor:
    if not isinstance(angular_tesselation, int) or angular_tesselation < 1:
        raise ValueError(f"angular_tesselation {angular_tesselation} must be an integer >= 1")
    if base_shape == "icosahedron":
        verts = _tessellate(_ICO_VERTS, _ICO_FACES, angular_tesselation, eps)
    elif base_shape == "octahedron":
        import itertools
        corners = np.array(list(itertools.product([-1, 1], repeat=3)))
        pairs = np.argwhere(_pairwise_sq(corners, _OCT_VERTS) == 2)
        faces = np.sort(np.reshape(pairs[:, 1], [3, -1]).T, 1)
        verts = _tessellate(_OCT_VERTS, faces, angular_tesselation, eps)
    else:
        raise ValueError(f"base_shape {base_shape} not supported")
    if remove_symmetries:
        mirrored = _pairwise_sq(verts, -verts) < eps
        verts = verts[np.any(np.triu(mirrored), 1), :]
    return torch.from_numpy(verts[:, ::-1].copy().T).to(dtype=torch.float32).contiguous()
