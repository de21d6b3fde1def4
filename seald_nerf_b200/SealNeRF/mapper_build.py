"""Mapper CONSTRUCTION geometry (host side, numpy / scipy): what the reference gets from trimesh, scikit-spatial, pytorch3d and
open3d when it turns a GUI edit config into the tensors the proxy mapping runs on (SealNeRF/seal_utils.py:156-242 bbox, :289-413
brush, :464-520 anchor, :595-636 helpers).  None of those packages is needed here:

  plane_best_fit      skspatial Plane.best_fit: centroid + direction of least variance (SVD)
  oriented_box        trimesh PointCloud.bounding_box_oriented: minimum-volume box over the convex hull's face directions, each with
                      the minimum-area rectangle of the projected hull (rotating over its edges)
  fit_curve_mesh      get_trimesh_fit (:599-632): k-nearest-neighbour prism mesh of a stroke, simplified by vertex clustering
                      (open3d simplify_vertex_clustering, 'Average' contraction)
  surface_points_mask mesh_surface_points_mask (:720-733): points with a 1e-4 neighbour outside the mesh
  uv_sphere_vertices  trimesh.creation.uv_sphere vertices (32 x 32)

This runs once per edit, on a few thousand points; the per-sample runtime is csrc/seal.cu.  Box corner order everywhere:
(-,-,-), (-,-,+), (-,+,-), (-,+,+), (+,-,-) ... in the box's own frame (`BOX_FACES` indexes it).
"""
import numpy as np

BOX_FACES = np.array([[0, 1, 3], [0, 3, 2], [4, 7, 5], [4, 6, 7], [0, 5, 1], [0, 4, 5], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4],
                      [1, 5, 7], [1, 7, 3]])
DEFAULT_TEST_DIR = np.array([0.4395064455, 0.617598629942, 0.652231566745])


def plane_best_fit(points):
    """-> (point on plane = centroid, unit normal)."""
    p = np.asarray(points, np.float64).reshape(-1, 3)
    c = p.mean(0)
    u, _, _ = np.linalg.svd((p - c).T)
    return c, u[:, 2]


def project_points(normal, point, targets):
    """Orthogonal projection onto the plane (seal_utils.py:736-744)."""
    n = np.asarray(normal, np.float64)
    t = np.asarray(targets, np.float64)
    return t - ((t - point) @ n)[:, None] / (n @ n) * n


def _min_area_rect(pts2):
    """Minimum-area enclosing rectangle of 2-D points: (area, axis u, axis v, lo[2], hi[2]); one side lies on a hull edge."""
    from scipy.spatial import ConvexHull
    try:
        h = pts2[ConvexHull(pts2).vertices]
    except Exception:  # collinear
        h = pts2
    best = None
    n = len(h)
    for i in range(n):
        e = h[(i + 1) % n] - h[i]
        L = np.linalg.norm(e)
        if L < 1e-14:
            continue
        u = e / L
        v = np.array([-u[1], u[0]])
        a, b = pts2 @ u, pts2 @ v
        area = (a.max() - a.min()) * (b.max() - b.min())
        if best is None or area < best[0] - 1e-15:
            best = (area, u, v, np.array([a.min(), b.min()]), np.array([a.max(), b.max()]))
    if best is None:
        u, v = np.array([1.0, 0.0]), np.array([0.0, 1.0])
        best = (0.0, u, v, pts2.min(0), pts2.max(0))
    return best


def oriented_box(points):
    """Corners [8,3] of the minimum-volume oriented bounding box of a point cloud."""
    from scipy.spatial import ConvexHull
    p = np.asarray(points, np.float64).reshape(-1, 3)
    normals = []
    try:
        hull = ConvexHull(p)
        hv = p[hull.vertices]
        for eq in hull.equations:
            n = eq[:3] / np.linalg.norm(eq[:3])
            if n[np.argmax(np.abs(n))] < 0:
                n = -n
            if not any(abs(n @ m) > 1 - 1e-10 for m in normals):
                normals.append(n)
    except Exception:  # flat / degenerate cloud: the plane normal and the two in-plane principal directions
        hv = p
        c = p.mean(0)
        u, _, _ = np.linalg.svd((p - c).T)
        normals = [u[:, 2], u[:, 1], u[:, 0]]
    best = None
    for n in normals:
        a = np.eye(3)[np.argmin(np.abs(n))]
        u = np.cross(n, a)
        u /= np.linalg.norm(u)
        v = np.cross(n, u)
        h = hv @ n
        area, r0, r1, lo2, hi2 = _min_area_rect(np.stack([hv @ u, hv @ v], 1))
        vol = area * (h.max() - h.min())
        if best is None or vol < best[0] - 1e-15:
            ax = np.stack([n, r0[0] * u + r0[1] * v, r1[0] * u + r1[1] * v])  # box axes (rows), orthonormal
            best = (vol, ax, np.array([h.min(), lo2[0], lo2[1]]), np.array([h.max(), hi2[0], hi2[1]]))
    _, ax, lo, hi = best
    if np.linalg.det(ax) < 0:
        ax[2] = -ax[2]
        lo[2], hi[2] = -hi[2], -lo[2]
    return np.array([np.array([i, j, k]) @ ax for i in (lo[0], hi[0]) for j in (lo[1], hi[1]) for k in (lo[2], hi[2])])


def box_triangles(corners):
    return np.asarray(corners, np.float64)[BOX_FACES]


def moller_trumbore_any(ray_o, ray_d, tris, eps=1e-8):
    """Does ray i hit any triangle at t >= 0 (seal_utils.py:639-672)."""
    E1, E2 = tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]
    N = np.cross(E1, E2)
    invdet = 1.0 / -(ray_d @ N.T + eps)
    A0 = ray_o[:, None] - tris[None, :, 0]
    DA0 = np.cross(A0, np.broadcast_to(ray_d[:, None], A0.shape))
    u = np.einsum("mnd,nd->mn", DA0, E2) * invdet
    v = -np.einsum("mnd,nd->mn", DA0, E1) * invdet
    t = np.einsum("mnd,nd->mn", A0, N) * invdet
    return ((t >= 0) & (u >= 0) & (v >= 0) & (u + v <= 1)).any(1)


def points_in_mesh(points, tris, test_dir=None, chunk=8192):
    """Hit along +d AND along -d (seal_utils.py:675-693)."""
    p = np.asarray(points, np.float64).reshape(-1, 3)
    tris = np.asarray(tris, np.float64)
    d = np.asarray(DEFAULT_TEST_DIR if test_dir is None else test_dir, np.float64).reshape(1, 3)
    out = np.zeros(len(p), bool)
    for s in range(0, len(p), chunk):
        q = p[s:s + chunk]
        dd = np.repeat(d, len(q), 0)
        out[s:s + chunk] = moller_trumbore_any(q, dd, tris) & moller_trumbore_any(q, -dd, tris)
    return out


def surface_points_mask(tris, points, offset=1e-4):
    """Points with at least one of their six axis neighbours at `offset` outside the mesh (seal_utils.py:720-733)."""
    p = np.asarray(points, np.float64).reshape(-1, 3)
    offs = np.array([[0, 0, 1], [0, 0, -1], [0, 1, 0], [0, -1, 0], [1, 0, 0], [-1, 0, 0]], np.float64) * offset
    mask = np.zeros(len(p), bool)
    for o in offs:
        mask |= ~points_in_mesh(p + o, tris)
    return mask


def fit_curve_mesh(points, normal, growth=(-0.3, 1.0), simplify_voxel=16, K=10):
    """Triangles [F,3,3] of a 'curve' stroke: every point is joined to pairs of its K nearest neighbours on a lower sheet
    (points + normal * growth[0]) and an upper sheet (growth[1]) plus side walls, then vertices are clustered on a grid of
    max-extent / simplify_voxel cells (cluster = mean of its vertices) and degenerate triangles dropped."""
    from scipy.spatial import cKDTree
    p = np.asarray(points, np.float64).reshape(-1, 3)
    N = len(p)
    K = min(K, N)
    idx = cKDTree(p).query(p, K)[1].reshape(N, K)
    faces = []
    for i in range(N):
        for j in range(1, K):
            for k in range(j + 1, K):
                x, y, z = i, idx[i][j], idx[i][k]
                faces += [[x, y, z], [x + N, y + N, z + N], [x, y, x + N], [x + N, y, y + N]]
    verts = np.concatenate([p + normal * growth[0], p + normal * growth[1]])
    faces = np.asarray(faces, np.int64).reshape(-1, 3)
    lo = verts.min(0)
    cell = max((verts.max(0) - lo).max() / simplify_voxel, 1e-12)
    key = np.floor((verts - lo) / cell).astype(np.int64)
    _, inv = np.unique(key, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    cnt = np.bincount(inv)
    cv = np.stack([np.bincount(inv, verts[:, d]) / cnt for d in range(3)], 1)
    f = inv[faces]
    f = f[(f[:, 0] != f[:, 1]) & (f[:, 1] != f[:, 2]) & (f[:, 0] != f[:, 2])]
    f = np.unique(np.sort(f, 1), axis=0)
    return cv[f]


def uv_sphere_vertices(radius=1.0, count=(32, 32)):
    """Vertices of a latitude / longitude sphere (poles once)."""
    th = np.linspace(0, np.pi, count[0] + 1)[1:-1]
    ph = np.linspace(0, 2 * np.pi, count[1], endpoint=False)
    T, P = np.meshgrid(th, ph, indexing="ij")
    v = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], -1).reshape(-1, 3)
    return np.concatenate([[[0, 0, 1.0]], v, [[0, 0, -1.0]]]) * radius
