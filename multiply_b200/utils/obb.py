"""Oriented bounding box of a posed body — the host-side stand-in for trimesh's ``Trimesh.bounding_box_oriented``
that the reference culls rays against (/root/reference/code/lib/model/multiply.py:208-214):

    oriented_box = smpl_mesh.bounding_box_oriented.copy()
    Box(oriented_box.primitive.extents * 1.2, oriented_box.transform)

trimesh is a third-party dependency that is absent from the reference tree and from this image, so this is a
restatement of its PUBLISHED algorithm (``trimesh.bounds.oriented_bounds`` / ``oriented_bounds_2D``, trimesh 3.x / 4.x),
and parity with the library is UNPINNED: there is nothing to run it against here.  What the algorithm does:

  1. convex hull of the vertices (qhull; scipy.spatial.ConvexHull here);
  2. candidate box normals = the hull's face normals, folded onto one hemisphere, converted to spherical angles and
     de-duplicated after rounding to one decimal (``angle_digits=1``): the first face of each angle cell is kept;
  3. for every candidate: project the hull vertices along it, height = extent along the normal, footprint = the
     minimum-area rectangle of the projected points (rotating calipers: one candidate per edge of the 2-D hull);
  4. keep the candidate with the smallest volume.

The result is therefore not the exact minimum-volume box (candidates are restricted to hull face normals, thinned at
0.1 rad), just as trimesh's is not.  ``tests/test_obb.py`` checks the properties any such box must have (contains the
points, tight in its own axes, no larger than the axis-aligned and PCA boxes, exact on a rotated cuboid).
The default culling of the mirror stays the device-side axis-aligned box (no host round trip); ``Multiply(...,
culling="obb")`` selects this one at the price of one device->host copy of the posed vertices per person and frame —
which is what the reference pays as well (``smpl_verts[0].detach().cpu().numpy()``, multiply.py:208).
"""
import numpy as np


def _min_area_rectangle(pts2):
    """trimesh.bounds.oriented_bounds_2D: (u, v, lo, hi) — in-plane unit axes and the bounds of the points along them for
    the hull edge whose aligned rectangle has the smallest area."""
    from scipy.spatial import ConvexHull
    try:
        hull = ConvexHull(pts2, qhull_options="QbB")
        edges = hull.points[hull.simplices]                      # [n,2,2]
        hp = hull.points[hull.vertices]
    except Exception:                                            # degenerate footprint (collinear points)
        d = pts2.max(0) - pts2.min(0)
        u = np.array([1.0, 0.0]) if d[0] >= d[1] else np.array([0.0, 1.0])
        edges = np.stack([np.zeros(2), u])[None]
        hp = pts2
    ev = edges[:, 1] - edges[:, 0]
    ev = ev / np.maximum(np.linalg.norm(ev, axis=1, keepdims=True), 1e-300)
    pv = np.fliplr(ev) * np.array([-1.0, 1.0])
    x = ev @ hp.T
    y = pv @ hp.T
    lo = np.stack([x.min(1), y.min(1)], 1)
    hi = np.stack([x.max(1), y.max(1)], 1)
    area = np.prod(hi - lo, axis=1)
    i = int(area.argmin())
    return ev[i], pv[i], lo[i], hi[i]


def _frame_of(n):
    """Two unit vectors spanning the plane orthogonal to the unit vector n (the in-plane rotation is irrelevant: the
    rectangle search is rotation invariant)."""
    a = np.array([1.0, 0.0, 0.0]) if abs(n[0]) < 0.9 else np.array([0.0, 1.0, 0.0])
    e0 = np.cross(n, a)
    e0 /= np.linalg.norm(e0)
    e1 = np.cross(n, e0)
    return e0, e1


def oriented_bounds(points, angle_digits=1):
    """(center[3], half_extent[3], rot[3,3]) of the oriented box of ``points`` [N,3]: rows of ``rot`` are the box axes in
    world coordinates, a point p is inside iff |rot @ (p - center)| <= half_extent.  float64."""
    from scipy.spatial import ConvexHull
    pts = np.asarray(points, dtype=np.float64)
    hull = ConvexHull(pts)
    verts = pts[hull.vertices]
    normals = hull.equations[:, :3].copy()
    normals /= np.linalg.norm(normals, axis=1, keepdims=True)
    # trimesh.util.vector_hemisphere: fold onto z > 0 (ties: y > 0, then x > 0)
    neg = (normals[:, 2] < 0) | ((normals[:, 2] == 0) & (normals[:, 1] < 0)) | \
          ((normals[:, 2] == 0) & (normals[:, 1] == 0) & (normals[:, 0] < 0))
    normals[neg] *= -1.0
    # trimesh.util.vector_to_spherical + grouping.unique_rows(digits=angle_digits): first face of every angle cell
    sph = np.stack([np.arctan2(normals[:, 1], normals[:, 0]), np.arccos(np.clip(normals[:, 2], -1.0, 1.0))], 1)
    key = np.round(sph * 10 ** angle_digits).astype(np.int64)
    _, first = np.unique(key, axis=0, return_index=True)
    best = None
    for i in np.sort(first):
        n = normals[i]
        e0, e1 = _frame_of(n)
        h = verts @ n
        u, v, lo, hi = _min_area_rectangle(np.stack([verts @ e0, verts @ e1], 1))
        vol = float(np.prod(hi - lo) * (h.max() - h.min()))
        if best is None or vol < best[0]:
            ax0 = u[0] * e0 + u[1] * e1
            ax1 = v[0] * e0 + v[1] * e1
            best = (vol, np.stack([ax0, ax1, n]), np.array([lo[0], lo[1], h.min()]), np.array([hi[0], hi[1], h.max()]))
    _, rot, lo, hi = best
    if np.linalg.det(rot) < 0:                                   # keep a proper rotation
        rot = rot * np.array([[1.0], [1.0], [-1.0]])
        lo, hi = np.array([lo[0], lo[1], -hi[2]]), np.array([hi[0], hi[1], -lo[2]])
    center = rot.T @ ((lo + hi) * 0.5)
    return center, (hi - lo) * 0.5, rot


def culling_box(verts_posed, inflate=1.2):
    """The reference's culling box (multiply.py:208-214): the oriented box with its extents scaled by ``inflate`` about
    its own centre."""
    c, h, rot = oriented_bounds(verts_posed)
    return c, h * inflate, rot
