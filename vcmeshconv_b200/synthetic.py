"""Synthetic mesh templates (the reference ships no meshes).

  sphere(n)     convex hull of n unit-normalised N(0,1)^3 points -> closed
                genus-0 triangle mesh with n vertices, 2n-4 faces, 3n-6 edges.
                n=6890 has exactly SMPL's counts (6890 v / 13776 f), n=5023 is
                FLAME-sized, n=150000 is the high-resolution case.
  torus(a, b)   regular a x b triangulated torus (every vertex has degree 6).
  tets(n)       Delaunay tetrahedra of n uniform points, each tet emitted as 4
                triangles: irregular, non-manifold connectivity with a long
                degree tail (config C5).

All generators are deterministic in (size, seed).
"""
import numpy as np


def sphere(n_vert, seed=0):
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(seed)
    pts = rng.standard_normal((n_vert, 3))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    hull = ConvexHull(pts)
    faces = np.ascontiguousarray(hull.simplices, dtype=np.int32)
    # consistent outward orientation (ConvexHull does not guarantee one)
    a, b, c = pts[faces[:, 0]], pts[faces[:, 1]], pts[faces[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(b - a, c - b), a + b + c) < 0
    faces[flip] = faces[flip][:, ::-1]
    # ConvexHull's facet order is implementation-defined; sort for determinism
    faces = faces[np.lexsort((faces[:, 2], faces[:, 1], faces[:, 0]))]
    return pts.astype(np.float32), np.ascontiguousarray(faces)


def torus(n_major=65, n_minor=106, r_major=1.0, r_minor=0.35):
    u = np.arange(n_major) * (2 * np.pi / n_major)
    v = np.arange(n_minor) * (2 * np.pi / n_minor)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    pts = np.stack([(r_major + r_minor * np.cos(vv)) * np.cos(uu),
                    (r_major + r_minor * np.cos(vv)) * np.sin(uu),
                    r_minor * np.sin(vv)], -1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(n_major), np.arange(n_minor), indexing="ij")
    i1, j1 = (i + 1) % n_major, (j + 1) % n_minor
    vid = lambda a, b: a * n_minor + b
    f0 = np.stack([vid(i, j), vid(i1, j), vid(i1, j1)], -1).reshape(-1, 3)
    f1 = np.stack([vid(i, j), vid(i1, j1), vid(i, j1)], -1).reshape(-1, 3)
    return pts.astype(np.float32), np.ascontiguousarray(np.concatenate([f0, f1]), dtype=np.int32)


def tets(n_vert, seed=0):
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = rng.random((n_vert, 3))
    t = Delaunay(pts).simplices.astype(np.int32)
    t = t[np.lexsort((t[:, 3], t[:, 2], t[:, 1], t[:, 0]))]
    faces = np.concatenate([t[:, [0, 1, 2]], t[:, [0, 1, 3]], t[:, [0, 2, 3]], t[:, [1, 2, 3]]])
    return pts.astype(np.float32), np.ascontiguousarray(faces, dtype=np.int32)


def tets_capped(n_vert, max_count=32, seed=0):
    """Config C5: the tetrahedral template with the neighbour count (self included, as the connectivity tables count
    it) capped at `max_count`. Delaunay tets are taken in sorted order and a tet is kept only if none of its four
    vertices would get more than max_count - 1 distinct neighbours: an irregular, non-manifold tet complex (holes
    where hub vertices were saturated) whose radius-1 table has N_max = max_count exactly when the cap binds."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = rng.random((n_vert, 3))
    t = Delaunay(pts).simplices.astype(np.int32)
    t = t[np.lexsort((t[:, 3], t[:, 2], t[:, 1], t[:, 0]))]
    adj = [set() for _ in range(n_vert)]
    keep = np.zeros(len(t), dtype=bool)
    cap = max_count - 1
    for k, tet in enumerate(t.tolist()):
        ok = True
        for v in tet:
            a = adj[v]
            if len(a) + sum(1 for w in tet if w != v and w not in a) > cap:
                ok = False
                break
        if ok:
            keep[k] = True
            for v in tet:
                adj[v].update(w for w in tet if w != v)
    t = t[keep]
    faces = np.concatenate([t[:, [0, 1, 2]], t[:, [0, 1, 3]], t[:, [0, 2, 3]], t[:, [1, 2, 3]]])
    return pts.astype(np.float32), np.ascontiguousarray(faces, dtype=np.int32)


def write_obj(path, pts, faces, green=()):
    """OBJ with per-vertex colour, the input format of the reference tool
    (GraphSampling/meshLoader.h:221-330: 'v x y z r g b', 'f a b c', 1-based).
    Vertices listed in `green` get colour exactly (0,1,0) = must stay a centre."""
    green = set(int(g) for g in green)
    with open(path, "w") as f:
        for i, p in enumerate(pts):
            col = "0 1 0" if i in green else "0.7 0.7 0.7"
            f.write("v %.6f %.6f %.6f %s\n" % (p[0], p[1], p[2], col))
        for t in faces:
            f.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))


def face_normals(pts, faces):
    """Unit face normals as the reference dataset computes them
    (GraphAutoEncoder/graphVAE_train.py:115-120)."""
    v0, v1, v2 = pts[..., faces[:, 0], :], pts[..., faces[:, 1], :], pts[..., faces[:, 2], :]
    n = np.cross(v1 - v0, v2 - v1)
    return n / (np.sqrt((n ** 2).sum(-1, keepdims=True)) + 1e-6)


# Layer presets: (stride, pool_radius, unpool_radius) per level.
def alternating_levels(n_levels):
    """The reference's set_7k_mesh_layers pattern (GraphSampling/main.cpp:17-32):
    stride 1,2,1,2,... with both radii equal to the stride."""
    return [(1, 1, 1) if l % 2 == 0 else (2, 2, 2) for l in range(n_levels)]
