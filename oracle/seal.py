"""numpy restatement of the Seal proxy-mapping RUNTIME (SealNeRF/seal_utils.py, SealNeRF/color_utils.py).
TEST INFRASTRUCTURE ONLY.

A mapper is a plain dict of numpy arrays — exactly the tensors the reference keeps in `SealMapper.map_data`,
`.map_triangles` and `.map_test_dir` after construction (construction itself needs trimesh / pytorch3d / scikit-spatial
and is outside the hot path):

    type            'bbox' | 'brush' | 'anchor'
    map_bound       [B,2,3] or [2,3]            seal_utils.py:137
    map_triangles   [F,3,3]                     seal_utils.py:218-219,359-360,498-499
    map_test_dir    [1,3] or None               seal_utils.py:366 (brush), None elsewhere
    bbox  : transform [4,4], rotation [3,3], scale [3], center [3] (+ optional empty_bound [2,3], map_source [3])   :221-240
    brush : normal_expand [3], center [3], border_points [P,3], attenuation_distance, attenuation_mode              :368-383
    anchor: v_anchor [3], v_offset [3], v_h [3], len_h, radius, scale [3]                                            :501-514
    colour: hsv [3] / rgb [3] + rgb_light_offset / image [H,W,3] + image_mask [H,W] + v_image_{norm,o,w,h}           :233-237,384-411

Functions (reference file:line each follows):
    moller_trumbore      seal_utils.py:638-672
    points_in_mesh       seal_utils.py:675-693
    map_mask             seal_utils.py:132-153
    project_points       seal_utils.py:736-744
    map_to_origin        bbox :244-286, brush :415-461, anchor :522-578
    rgb2hsv / hsv2rgb    color_utils.py:31-63
    modify_hsv / modify_rgb / map_color   seal_utils.py:747-777, :48-81

Pinned against the reference's own Python (imported on CPU with its heavy third-party imports stubbed) by
tests/golden/make_seal_golden.py -> tests/golden/seal.npz.
"""
import numpy as np

DEFAULT_TEST_DIR = np.array([[0.4395064455, 0.617598629942, 0.652231566745]], np.float32)  # seal_utils.py:686-688


def moller_trumbore(ray_o, ray_d, tris, eps=1e-8):
    """[m,3], [m,3], [F,3,3] -> bool [m]: does ray i hit ANY triangle (t >= 0)."""
    ray_o = np.asarray(ray_o, np.float32)
    ray_d = np.asarray(ray_d, np.float32)
    tris = np.asarray(tris, np.float32)
    E1 = tris[:, 1] - tris[:, 0]
    E2 = tris[:, 2] - tris[:, 0]
    N = np.cross(E1, E2).astype(np.float32)
    invdet = (np.float32(1.0) / -(np.einsum("md,nd->mn", ray_d, N).astype(np.float32) + np.float32(eps))).astype(np.float32)
    A0 = (ray_o[:, None] - tris[None, :, 0]).astype(np.float32)
    DA0 = np.cross(A0, np.broadcast_to(ray_d[:, None], A0.shape)).astype(np.float32)
    u = np.einsum("mnd,nd->mn", DA0, E2).astype(np.float32) * invdet
    v = -np.einsum("mnd,nd->mn", DA0, E1).astype(np.float32) * invdet
    t = np.einsum("mnd,nd->mn", A0, N).astype(np.float32) * invdet
    hit = (t >= 0.0) & (u >= 0.0) & (v >= 0.0) & ((u + v) <= 1.0)
    return hit.any(1)


def points_in_mesh(points, triangles, rays_d=None, chunk=65536):
    points = np.asarray(points, np.float32)
    if rays_d is None:
        rays_d = DEFAULT_TEST_DIR
    rays_d = np.asarray(rays_d, np.float32).reshape(1, 3)
    out = np.zeros(points.shape[0], bool)
    for s in range(0, points.shape[0], chunk):
        p = points[s:s + chunk]
        d = np.repeat(rays_d, p.shape[0], 0)
        out[s:s + chunk] = moller_trumbore(p, d, triangles) & moller_trumbore(p, -d, triangles)
    return out


def map_mask(mapper, points):
    points = np.asarray(points, np.float32)
    bounds = np.asarray(mapper["map_bound"], np.float32)
    if bounds.ndim == 2:
        bounds = bounds[None]
    nz = (points != 0).all(1)  # `points.all(1)`: rows with a zero coordinate (the march's zero padding) never map
    bound_mask = np.zeros(points.shape[0], bool)
    for i in range(bounds.shape[0]):
        bound_mask |= nz & ((bounds[i, 1] > points) & (points > bounds[i, 0])).all(1)
    if not bound_mask.any():
        return bound_mask
    shape = points_in_mesh(points[bound_mask], mapper["map_triangles"], mapper.get("map_test_dir"))
    out = bound_mask.copy()
    out[bound_mask] = shape
    return out


def project_points(plane_norm, plane_point, target_points):
    plane_norm = np.asarray(plane_norm, np.float32)
    v = (target_points - np.asarray(plane_point, np.float32)).astype(np.float32)
    proj = ((v @ plane_norm)[:, None] / (plane_norm @ plane_norm) * plane_norm).astype(np.float32)
    return (target_points - proj).astype(np.float32)


def map_to_origin(mapper, points, dirs=None):
    """-> (points', dirs', mask) like SealBBoxMapper / SealBrushMapper / SealAnchorMapper.map_to_origin."""
    points = np.asarray(points, np.float32)
    dirs = None if dirs is None else np.asarray(dirs, np.float32)
    kind = mapper["type"]
    mask = map_mask(mapper, points)
    if not mask.any():
        return points, dirs, mask
    if kind == "bbox":
        inner = points[mask]
        T = np.asarray(mapper["transform"], np.float32)
        hom = np.vstack([inner.T, np.ones((1, inner.shape[0]), np.float32)])
        transformed = (T @ hom).T[:, :3]
        c = np.asarray(mapper["center"], np.float32)
        origin = ((transformed - c) * np.asarray(mapper["scale"], np.float32) + c).astype(np.float32)
        pc = points.copy()
        if "map_source" in mapper:
            sb = np.asarray(mapper["empty_bound"], np.float32)
            src = ((sb[1] > points) & (points > sb[0])).all(1)
            pc[src] = np.asarray(mapper["map_source"], np.float32)
        pc[mask] = origin
        dc = None
        if dirs is not None:
            dc = dirs.copy()
            dc[mask] = (np.asarray(mapper["rotation"], np.float32) @ dirs[mask].T).T
        return pc, dc, mask
    if kind == "brush":
        inner = points[mask]
        mode = mapper["attenuation_mode"]
        ne = np.asarray(mapper["normal_expand"], np.float32)
        if mode == "linear":
            proj = project_points(ne, mapper["center"], inner)
            bp = np.asarray(mapper["border_points"], np.float32)
            dist = np.sqrt(((proj[:, None, :].astype(np.float64) - bp[None].astype(np.float64)) ** 2).sum(-1)).min(1).astype(np.float32)
            mapped = inner - ne
            ad = np.float32(mapper["attenuation_distance"])
            filt = ad > dist
            mapped[filt] += (np.abs(ad - dist[filt]) / ad)[:, None] * ne[None]
        elif mode == "dry":
            mapped = inner
        else:
            raise NotImplementedError(mode)  # seal_utils.py:444-449
        pc = points.copy()
        pc[mask] = mapped
        return pc, dirs, mask
    if kind == "anchor":
        v_h = np.asarray(mapper["v_h"], np.float32)
        v_anchor = np.asarray(mapper["v_anchor"], np.float32)
        v_offset = np.asarray(mapper["v_offset"], np.float32)
        len_h, radius = np.float32(mapper["len_h"]), np.float32(mapper["radius"])
        proj = project_points(v_h, v_anchor, points)
        v_pp = proj - points
        dist = np.linalg.norm(v_pp, 2, 1).astype(np.float32)
        pop = (proj - (dist[:, None] / len_h) * v_offset).astype(np.float32)
        pop_anchor = np.linalg.norm(pop - v_anchor, 2, 1).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            cone = (pop_anchor <= radius) & (dist / (radius - pop_anchor) < len_h / radius * np.float32(1.1))
        side = (v_pp @ v_h) > 0
        valid = cone & side
        vd = dist[valid]
        v_map = -((len_h - vd) / 10)[:, None] * v_h[None] / len_h
        mapped = pop[valid] - v_map
        mapped = (mapped - v_anchor) * np.asarray(mapper["scale"], np.float32) + v_anchor
        pc = points.copy()
        pc[valid] = mapped
        return pc, dirs, valid
    raise NotImplementedError(kind)


# ---- colour ---------------------------------------------------------------------------------------------------------
def rgb2hsv(rgb):
    rgb = np.asarray(rgb, np.float32)
    cmax = rgb.max(1)
    idx = rgb.argmax(1)
    cmin = rgb.min(1)
    delta = cmax - cmin
    r, g, b = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        h = np.where(idx == 0, np.mod((g - b) / delta, 6), np.where(idx == 1, (b - r) / delta + 2, (r - g) / delta + 4))
        h = np.where(delta == 0, 0.0, h) / 6.0
        s = np.where(cmax == 0, 0.0, delta / cmax)
    return np.stack([h, s, cmax], 1).astype(np.float32)


def hsv2rgb(hsv):
    hsv = np.asarray(hsv, np.float32)
    h, s, v = hsv[:, 0], hsv[:, 1], hsv[:, 2]
    c = v * s
    x = c * (1.0 - np.abs(np.mod(h * 6.0, 2.0) - 1.0))
    m = v - c
    o = np.zeros_like(c)
    idx = (h * 6.0).astype(np.int64).astype(np.uint8) % 6
    table = [(c, x, o), (x, c, o), (o, c, x), (o, x, c), (x, o, c), (c, o, x)]
    rgb = np.zeros_like(hsv)
    for k, (rr, gg, bb) in enumerate(table):
        sel = idx == k
        rgb[sel, 0], rgb[sel, 1], rgb[sel, 2] = rr[sel], gg[sel], bb[sel]
    return (rgb + m[:, None]).astype(np.float32)


def modify_hsv(rgb, modification):
    if rgb.shape[0] == 0:
        return rgb
    hsv = rgb2hsv(rgb)
    hsv = hsv + np.asarray(modification, np.float32)[None]
    return hsv2rgb(hsv)


def modify_rgb(rgb, modification, light_offset=0.0):
    if rgb.shape[0] == 0:
        return rgb
    hsl = rgb2hsv(rgb)
    mod = rgb2hsv(np.asarray(modification, np.float32).reshape(-1, 3))
    raw_l = hsl[:, 2]
    off = raw_l - raw_l.mean(dtype=np.float32)
    out = np.empty_like(hsl)
    out[:, :2] = np.broadcast_to(mod[:, :2], (hsl.shape[0], 2))
    out[:, 2] = np.clip(mod[:, 2] + off + np.float32(light_offset), 0.0, 1.0)
    return hsv2rgb(out)


def map_color(mapper, points, dirs, colors):
    colors = np.asarray(colors, np.float32)
    if "hsv" in mapper:
        colors = modify_hsv(colors, mapper["hsv"])
    if "rgb" in mapper:
        colors = modify_rgb(colors, mapper["rgb"], mapper.get("rgb_light_offset", 0.0))
    if "image" in mapper:
        image = np.asarray(mapper["image"], np.float32)
        H, W, _ = image.shape
        v_o = np.asarray(mapper["v_image_o"], np.float32)
        v_ow = np.asarray(mapper["v_image_w"], np.float32) - v_o
        v_oh = np.asarray(mapper["v_image_h"], np.float32) - v_o
        proj = project_points(mapper["v_image_norm"], v_o, np.asarray(points, np.float32))
        v_op = proj - v_o
        iw = np.clip(np.floor(v_op @ v_ow / (v_ow @ v_ow) * W), 0, W - 1).astype(np.int64)
        ih = np.clip(np.floor(v_op @ v_oh / (v_oh @ v_oh) * H), 0, H - 1).astype(np.int64)
        m = np.asarray(mapper["image_mask"], np.float32)[ih, iw][:, None]
        mod = modify_rgb(colors, image[ih, iw], mapper.get("rgb_light_offset", 0.0))
        colors = m * mod + (1 - m) * colors
    return colors.astype(np.float32)


# ---- fixture builders (shared by the golden generator and the tests) ---------------------------------------------------
_BOX_FACES = np.array([[0, 1, 3], [0, 3, 2], [4, 7, 5], [4, 6, 7], [0, 5, 1], [0, 4, 5], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4],
                       [1, 5, 7], [1, 7, 3]])


def box_corners(center, half, R=None):
    s = np.array([[i, j, k] for i in (-1, 1) for j in (-1, 1) for k in (-1, 1)], np.float64)
    v = s * np.asarray(half, np.float64)
    if R is not None:
        v = v @ np.asarray(R, np.float64).T
    return v + np.asarray(center, np.float64)


def box_triangles(corners):
    return np.asarray(corners, np.float64)[_BOX_FACES].astype(np.float32)  # [12,3,3]


def rot_y(deg):
    a = np.deg2rad(deg)
    return np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])


def make_bbox_mapper(center=(0.0, 0.05, 0.0), half=(0.15, 0.15, 0.15), translate=(0.2, 0.0, 0.0), rot_deg=30.0, scale=(1.0, 1.0, 1.0),
                     bound_type="to", map_source=None, hsv=None, rgb=None, light_offset=0.0):
    """The tensors SealBBoxMapper.__init__ (seal_utils.py:168-242) derives from a config: OBB `raw` box -> scaled about its
    centre -> rigid transform; map_data holds the INVERSE transform/rotation/scale."""
    frm = box_corners(center, half)
    c = np.asarray(center, np.float64)
    T = np.eye(4)
    T[:3, :3] = rot_y(rot_deg)
    T[:3, 3] = np.asarray(translate, np.float64) + c - T[:3, :3] @ c  # rotate about the box centre, then translate
    to = (frm - c) * np.asarray(scale, np.float64) + c
    to = to @ T[:3, :3].T + T[:3, 3]
    meshes = {"to": [to], "from": [frm], "both": [to, frm]}[bound_type]
    tris = np.concatenate([box_triangles(m) for m in meshes], 0)
    bounds = np.stack([np.stack([m.min(0), m.max(0)]) for m in meshes]).astype(np.float32)
    mp = dict(type="bbox", map_bound=bounds if bound_type == "both" else bounds[0], map_triangles=tris, map_test_dir=None,
              transform=np.linalg.inv(T).astype(np.float32), rotation=np.linalg.inv(T[:3, :3]).astype(np.float32),
              scale=(1.0 / np.asarray(scale, np.float64)).astype(np.float32), center=c.astype(np.float32),
              force_fill_bound=np.stack([np.stack([m.min(0), m.max(0)]) for m in (to, frm)]).astype(np.float32))
    if map_source is not None:
        mp["empty_bound"] = np.stack([frm.min(0), frm.max(0)]).astype(np.float32)
        mp["map_source"] = np.asarray(map_source, np.float32)
    if hsv is not None:
        mp["hsv"] = np.asarray(hsv, np.float32)
    if rgb is not None:
        mp["rgb"] = np.asarray(rgb, np.float32)
        mp["rgb_light_offset"] = float(light_offset)
    return mp


def make_brush_mapper(mode="linear", pressure=0.02, depth=0.6, attenuation=0.02, rgb=None, hsv=None, seed=5, image=False):
    """Two 'line' strokes (seal_utils.py:321-360): each stroke's OBB spans its points pushed +2 and -brushDepth times
    normal_expand along the plane normal; border points lie on the stroke plane."""
    rng = np.random.default_rng(seed)
    n = np.array([1.0, 0.0, 0.0])  # stroke-plane normal (the fixture keeps the stroke boxes axis aligned with it)
    ne = n * pressure
    strokes, tris, bounds, border = [], [], [], []
    for k in range(2):
        c = np.array([0.12, 0.25 - 0.35 * k, 0.02 + 0.05 * k])
        half = np.array([(2 + depth) * pressure / 2, 0.12, 0.05])
        cc = c + n * (2 - depth) * pressure / 2
        corners = box_corners(cc, half)
        strokes.append(corners)
        tris.append(box_triangles(corners))
        bounds.append(np.stack([corners.min(0), corners.max(0)]))
        # border points: rim of the stroke rectangle in the plane through c
        u = np.linspace(-1, 1, 12)
        rim = [np.array([0, a * half[1], s * half[2]]) for a in u for s in (-1, 1)] + [np.array([0, s * half[1], a * half[2]]) for a in u for s in (-1, 1)]
        border.append(c + np.array(rim))
    mp = dict(type="brush", map_bound=np.stack(bounds).astype(np.float32), map_triangles=np.concatenate(tris, 0),
              map_test_dir=ne[None].astype(np.float32), normal_expand=ne.astype(np.float32),
              center=np.array([0.12, -0.1, 0.07], np.float32), border_points=np.concatenate(border, 0).astype(np.float32),
              attenuation_distance=float(attenuation), attenuation_mode=mode, force_fill_bound=np.stack(bounds).astype(np.float32))
    if rgb is not None:
        mp["rgb"] = np.asarray(rgb, np.float32)
        mp["rgb_light_offset"] = 0.0
    if hsv is not None:
        mp["hsv"] = np.asarray(hsv, np.float32)
    if image:
        mp["rgb_light_offset"] = 0.05
        mp["image"] = rng.random((6, 5, 3)).astype(np.float32)
        mp["image_mask"] = (rng.random((6, 5)) > 0.3).astype(np.float32)
        mp["v_image_norm"] = n.astype(np.float32)
        mp["v_image_o"] = np.array([0.12, -0.25, -0.05], np.float32)
        mp["v_image_w"] = np.array([0.12, 0.4, -0.05], np.float32)
        mp["v_image_h"] = np.array([0.12, -0.25, 0.15], np.float32)
    return mp


def make_anchor_mapper(scale=(1.0, 1.0, 1.0)):
    """Tensors of SealAnchorMapper.__init__ (seal_utils.py:475-520) for a plane z = 0.05 anchor pulled along +z/+x."""
    v_anchor = np.array([0.05, 0.0, 0.05])
    v_translation = np.array([0.04, 0.0, 0.12])
    radius = 0.1
    translated = v_anchor + v_translation
    projected = translated.copy()
    projected[2] = 0.05  # projection onto the plane z = 0.05
    v_offset = projected - v_anchor
    v_h = projected - translated
    sph = np.array([[radius * 1.1 * np.cos(a) * np.sin(b), radius * 1.1 * np.sin(a) * np.sin(b), radius * 1.1 * np.cos(b)]
                    for a in np.linspace(0, 2 * np.pi, 16) for b in np.linspace(0, np.pi, 9)]) + v_anchor
    pts = np.vstack([sph, v_anchor + 1.1 * v_translation, sph - 0.1 * v_translation])
    lo, hi = pts.min(0), pts.max(0)
    corners = box_corners((lo + hi) / 2, (hi - lo) / 2)
    return dict(type="anchor", map_bound=np.stack([lo, hi]).astype(np.float32), map_triangles=box_triangles(corners), map_test_dir=None,
                v_anchor=v_anchor.astype(np.float32), v_offset=v_offset.astype(np.float32), v_h=v_h.astype(np.float32),
                len_h=float(np.linalg.norm(v_h)), radius=radius, scale=np.asarray(scale, np.float32), map_source=True,
                force_fill_bound=np.stack([lo, hi]).astype(np.float32))
