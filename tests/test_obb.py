"""utils/obb.py — the oriented culling box (stand-in for trimesh's bounding_box_oriented, multiply.py:208-214).
trimesh is absent, so these are the properties such a box must have, not a comparison with the library."""
import numpy as np
import pytest

from multiply_b200.utils import obb
from multiply_b200 import scene as S


def _rot(rng):
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def _inside(p, c, h, rot, tol=1e-9):
    return np.all(np.abs((p - c) @ rot.T) <= h + tol, axis=1)


def test_rotated_cuboid_is_recovered():
    rng = np.random.default_rng(0)
    ext = np.array([1.7, 0.6, 0.25])
    q = _rot(rng)
    t = np.array([0.3, -0.2, 0.5])
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float) * ext / 2
    inner = (rng.random((500, 3)) - 0.5) * ext
    pts = np.concatenate([corners, inner]) @ q.T + t
    c, h, rot = obb.oriented_bounds(pts)
    assert np.allclose(rot @ rot.T, np.eye(3), atol=1e-12) and np.linalg.det(rot) > 0
    assert np.allclose(c, t, atol=1e-9)
    assert np.allclose(np.sort(2 * h), np.sort(ext), atol=1e-9)
    # every box axis is (up to sign) one of the cuboid's axes
    assert np.allclose(np.sort(np.abs(rot @ q).max(1)), 1.0, atol=1e-9)


def test_body_box_properties():
    sc = S.make_scene(P=2, S=16, seed=42)
    for person in sc["persons"]:
        v = person["verts_p"].double().numpy()
        c, h, rot = obb.oriented_bounds(v)
        assert np.allclose(rot @ rot.T, np.eye(3), atol=1e-12)
        assert _inside(v, c, h, rot).all()
        local = (v - c) @ rot.T
        assert np.allclose(local.max(0), h, atol=1e-9) and np.allclose(local.min(0), -h, atol=1e-9)   # tight
        vol = np.prod(2 * h)
        aabb = np.prod(v.max(0) - v.min(0))
        w, e = np.linalg.eigh(np.cov((v - v.mean(0)).T))
        pl = (v - v.mean(0)) @ e
        pca = np.prod(pl.max(0) - pl.min(0))
        assert vol <= aabb * (1 + 1e-9) and vol <= pca * (1 + 1e-9)
        c2, h2, rot2 = obb.culling_box(v, 1.2)
        assert np.allclose(c2, c) and np.allclose(h2, 1.2 * h) and np.allclose(rot2, rot)


def test_rigid_motion_equivariance():
    """Moving the body moves the box with it (same extents): the candidate set is the hull's own face normals."""
    rng = np.random.default_rng(3)
    sc = S.make_scene(P=1, S=16, seed=5)
    v = sc["persons"][0]["verts_p"].double().numpy()
    c, h, rot = obb.oriented_bounds(v, angle_digits=6)
    q = _rot(rng)
    t = rng.standard_normal(3)
    c2, h2, rot2 = obb.oriented_bounds(v @ q.T + t, angle_digits=6)
    assert np.allclose(np.sort(h2), np.sort(h), rtol=1e-6)
    assert np.allclose(c2, q @ c + t, atol=1e-6)
