"""Drop-in `SealNeRF.seal_utils` RUNTIME: the Seal editing proxy mapping (reference: SealNeRF/seal_utils.py).

Same classes and methods as the reference — `SealMapper.map_mask / map_color`, `SealBBoxMapper`, `SealBrushMapper`,
`SealAnchorMapper` with `map_to_origin(points, dirs) -> (points', dirs', mask)`, `get_seal_mapper(config_path,
config_dict, config_file)` and the `map_data` / `map_triangles` / `map_test_dir` attributes the renderers read
(SealDNeRF/renderer.py:52,157,252,272) — evaluated by the sm_100a kernels of csrc/seal.cu (stand-alone) and
csrc/raymarch.cu (fused into the march, `march_rays_seal`).

Scope: the per-sample runtime and mapper CONSTRUCTION from a GUI edit config (§8 f4; seal_utils.py:168-242,
304-413, 475-520).  The reference builds its boxes / planes / stroke meshes with trimesh, scikit-spatial, pytorch3d and open3d; here
the same geometry comes from `mapper_build.py` (numpy / scipy: minimum-volume oriented box, least-squares plane, k-NN stroke mesh
with vertex clustering), so `get_seal_mapper(config)` works for the bbox, brush ('line' and 'curve') and anchor tools without them.
`from_tensors` builds a mapper from finished tensors (tests, benchmarks).  Nothing is written to `config_path` (the reference
exports from.obj / to.obj there for its GUI).
"""
import ctypes as C
import json
import os

import numpy as np
import torch

from .. import _lib
from .._lib import ptr
from . import mapper_build as mb

_DEFAULT_TEST_DIR = (0.4395064455, 0.617598629942, 0.652231566745)  # seal_utils.py:686-688 (trimesh's magic direction)


class _MapperDesc(C.Structure):
    """seald_seal_mapper of include/seald_b200.h."""
    _fields_ = [("type", C.c_int32), ("n_bounds", C.c_int32), ("n_tris", C.c_int32), ("n_border", C.c_int32),
                ("bounds", C.c_void_p), ("tris", C.c_void_p), ("border", C.c_void_p),
                ("test_dir", C.c_float * 3), ("transform", C.c_float * 12), ("rotation", C.c_float * 9), ("scale", C.c_float * 3),
                ("center", C.c_float * 3), ("has_map_source", C.c_int32), ("empty_bound", C.c_float * 6), ("map_source", C.c_float * 3),
                ("normal_expand", C.c_float * 3), ("attenuation_distance", C.c_float), ("attenuation_mode", C.c_int32),
                ("v_anchor", C.c_float * 3), ("v_offset", C.c_float * 3), ("v_h", C.c_float * 3), ("len_h", C.c_float),
                ("radius", C.c_float)]


class _ColorDesc(C.Structure):
    """seald_seal_color of include/seald_b200.h."""
    _fields_ = [("has_hsv", C.c_int32), ("has_rgb", C.c_int32), ("has_image", C.c_int32), ("hsv", C.c_float * 3), ("rgb", C.c_float * 3),
                ("rgb_light_offset", C.c_float), ("img_h", C.c_int32), ("img_w", C.c_int32), ("image", C.c_void_p),
                ("image_mask", C.c_void_p), ("v_norm", C.c_float * 3), ("v_o", C.c_float * 3), ("v_w", C.c_float * 3), ("v_h", C.c_float * 3)]


def _np(v, shape=None):
    if torch.is_tensor(v):
        v = v.detach().cpu().numpy()
    a = np.asarray(v, np.float64)
    return a if shape is None else a.reshape(shape)


def _fill(dst, values):
    flat = np.asarray(values, np.float64).reshape(-1)
    for i in range(len(dst)):
        dst[i] = float(flat[i])


class SealMapper:
    """Root class of the seal mappers (seal_utils.py:18-153)."""

    TYPE_ID = -1

    def __init__(self, seal_config=None):
        self.config = seal_config or {}
        self.device = torch.device("cpu")
        self.dtype = torch.float32
        self.map_data = {}
        self.map_meshes = None
        self.map_triangles = None
        self.map_test_dir = None
        self._dev_cache = None

    # ---- construction from finished tensors -----------------------------------------------------------------------
    @classmethod
    def from_tensors(cls, map_data, map_triangles, map_test_dir=None):
        """map_data: the reference's `map_data` dict (numpy / lists / tensors); map_triangles [F,3,3]."""
        self = cls.__new__(cls)
        SealMapper.__init__(self, {})
        self.map_data = dict(map_data)
        self.map_triangles = torch.as_tensor(np.asarray(_np(map_triangles), np.float32))
        self.map_test_dir = None if map_test_dir is None else torch.as_tensor(np.asarray(_np(map_test_dir), np.float32)).reshape(1, 3)
        return self

    # ---- device descriptor -------------------------------------------------------------------------------------------
    def map_data_conversion(self, T=None, force=False):
        if T is not None and (self._dev_cache is None or self._dev_cache["device"] != T.device):
            self._build(T.device)

    def _fill_type(self, d, device, keep):
        raise NotImplementedError()

    def _build(self, device):
        if device.type != "cuda":
            raise RuntimeError("seald_b200 Seal mapping needs CUDA tensors (no CPU fallback)")
        md = self.map_data
        bounds = torch.as_tensor(np.asarray(_np(md["map_bound"], (-1, 2, 3)), np.float32)).to(device).contiguous()
        tris = torch.as_tensor(np.asarray(_np(self.map_triangles, (-1, 3, 3)), np.float32)).to(device).contiguous()
        d = _MapperDesc()
        d.type = self.TYPE_ID
        d.n_bounds, d.n_tris = bounds.shape[0], tris.shape[0]
        d.bounds, d.tris = bounds.data_ptr(), tris.data_ptr()
        _fill(d.test_dir, _np(self.map_test_dir) if self.map_test_dir is not None else _DEFAULT_TEST_DIR)
        keep = [bounds, tris]
        self._fill_type(d, device, keep)
        # colour
        c = _ColorDesc()
        c.has_hsv, c.has_rgb, c.has_image = int("hsv" in md), int("rgb" in md), int("image" in md)
        if c.has_hsv:
            _fill(c.hsv, _np(md["hsv"]))
        if c.has_rgb:
            _fill(c.rgb, _np(md["rgb"]))
        c.rgb_light_offset = float(_np(md.get("rgb_light_offset", 0.0)))
        if c.has_image:
            img = torch.as_tensor(np.asarray(_np(md["image"]), np.float32)).to(device).contiguous()
            msk = torch.as_tensor(np.asarray(_np(md["image_mask"]), np.float32)).to(device).contiguous()
            c.img_h, c.img_w = img.shape[0], img.shape[1]
            c.image, c.image_mask = img.data_ptr(), msk.data_ptr()
            _fill(c.v_norm, _np(md["v_image_norm"])); _fill(c.v_o, _np(md["v_image_o"]))
            _fill(c.v_w, _np(md["v_image_w"])); _fill(c.v_h, _np(md["v_image_h"]))
            keep += [img, msk]
        self.device = device
        self._dev_cache = {"device": device, "desc": d, "color": c, "keep": keep,
                           "scratch_i": torch.zeros(1, dtype=torch.int32, device=device),
                           "scratch_f": torch.zeros(4, dtype=torch.float32, device=device)}

    def descriptor(self, device):
        """ctypes `seald_seal_mapper` for `device` (what `seald_march_rays_seal` takes)."""
        if self._dev_cache is None or self._dev_cache["device"] != device:
            self._build(device)
        return self._dev_cache["desc"]

    @property
    def fusable(self):
        """bbox and brush mappers act per sample and can run inside the march kernels; the anchor mapper has a batch-wide
        early exit (seal_utils.py:526-528) and runs as a separate op."""
        return self.TYPE_ID in (0, 1)

    @property
    def has_color_map(self):
        return any(k in self.map_data for k in ("hsv", "rgb", "image"))

    # ---- reference API -----------------------------------------------------------------------------------------------
    def map_to_origin(self, points, dirs=None):
        """points, dirs [N,3] -> (points', dirs', mask bool [N])   (bbox :244-286, brush :415-461, anchor :522-578)."""
        _lib.require_cuda(points, dirs)
        pts = points.detach().float().contiguous().view(-1, 3)
        dr = None if dirs is None else dirs.detach().float().contiguous().view(-1, 3)
        desc = self.descriptor(pts.device)
        M = pts.shape[0]
        p_out = torch.empty_like(pts)
        d_out = None if dr is None else torch.empty_like(dr)
        mask = torch.empty(M, dtype=torch.bool, device=pts.device)
        _lib.call("seald_seal_map_to_origin", C.byref(desc), ptr(pts), ptr(dr), M, None, ptr(p_out), ptr(d_out), ptr(mask),
                  ptr(self._dev_cache["scratch_i"]), _lib.stream())
        return p_out, d_out, mask

    def map_mask(self, points):
        """Inside one of the map AABBs and inside the mesh (seal_utils.py:132-153)."""
        d = self.descriptor(points.device)
        saved = (d.type, d.attenuation_mode)
        d.type, d.attenuation_mode = 1, 1  # evaluated as a 'dry' brush: mask only, points untouched (the struct is copied at launch)
        try:
            return SealMapper.map_to_origin(self, points, None)[2]
        finally:
            d.type, d.attenuation_mode = saved

    def map_color(self, points, dirs, colors):
        """colours of the MAPPED samples -> edited colours (seal_utils.py:48-81).  `points` are the mapped points."""
        if not self.has_color_map or colors.shape[0] == 0:
            return colors
        _lib.require_cuda(points, colors)
        self.descriptor(colors.device)
        out = colors.detach().float().contiguous().clone()
        M = out.shape[0]
        mask = torch.ones(M, dtype=torch.bool, device=out.device)
        pts = points.detach().float().contiguous()
        _lib.call("seald_seal_map_color", C.byref(self._dev_cache["color"]), ptr(pts), ptr(mask), ptr(out), M, None,
                  ptr(self._dev_cache["scratch_f"]), _lib.stream())
        return out

    def map_color_masked_(self, points, mask, colors):
        """In-place `colors[mask] = map_color(points[mask], ., colors[mask])` without the gathers (colors fp32 [M,3])."""
        if not self.has_color_map or colors.shape[0] == 0:
            return colors
        self.descriptor(colors.device)
        assert colors.dtype == torch.float32 and colors.is_contiguous()
        _lib.call("seald_seal_map_color", C.byref(self._dev_cache["color"]), ptr(points), ptr(mask), ptr(colors), colors.shape[0], None,
                  ptr(self._dev_cache["scratch_f"]), _lib.stream())
        return colors


class SealBBoxMapper(SealMapper):
    """Bounding-box tool: rigid transform + per-axis scale of the space inside an oriented box (seal_utils.py:156-286)."""
    TYPE_ID = 0

    def __init__(self, config_path, seal_config):
        super().__init__(seal_config)
        T = np.array(seal_config["transform"], np.float64)
        R = T[:3, :3]
        scale = np.array(seal_config["scale"], np.float64)
        raw = np.array(seal_config["raw"], np.float64).reshape(-1, 3)
        # get_trimesh_box (:595-596): the oriented bounding box of the raw points (8 box corners are taken as the box itself)
        frm = _box_from_corners(raw) if raw.shape[0] == 8 else mb.oriented_box(raw)
        from_center = frm.mean(0)
        to = (frm - from_center) * scale + from_center
        to = to @ R.T + T[:3, 3]
        to_center = to.mean(0)
        bound_type = seal_config.get("boundType", "to")
        fill_bounds = np.stack([np.stack([m.min(0), m.max(0)]) for m in (to, frm)])
        meshes = {"to": [to], "from": [frm], "both": [to, frm]}[bound_type]
        self.map_triangles = torch.as_tensor(np.concatenate([m[_BOX_FACES] for m in meshes], 0).astype(np.float32))
        self.map_data = {
            "force_fill_bound": fill_bounds, "map_bound": fill_bounds if bound_type == "both" else np.stack([meshes[0].min(0), meshes[0].max(0)]),
            "pose_center": (from_center + to_center) / 2, "pose_radius": np.linalg.norm(from_center - to_center, 2) * 10,
            "transform": np.linalg.inv(T), "rotation": np.linalg.inv(R), "scale": 1 / scale, "center": from_center,
        }
        _color_entries(self.map_data, seal_config)
        if seal_config.get("mapSource"):
            self.map_data["empty_bound"] = np.stack([frm.min(0), frm.max(0)])
            self.map_data["map_source"] = seal_config["mapSource"]

    def _fill_type(self, d, device, keep):
        md = self.map_data
        _fill(d.transform, _np(md["transform"], (4, 4))[:3])
        _fill(d.rotation, _np(md["rotation"], (3, 3)))
        _fill(d.scale, _np(md["scale"]))
        _fill(d.center, _np(md["center"]))
        d.has_map_source = int("map_source" in md)
        if d.has_map_source:
            _fill(d.empty_bound, _np(md["empty_bound"], (2, 3)))
            _fill(d.map_source, _np(md["map_source"]))


class SealBrushMapper(SealMapper):
    """Brush tool: raise / lower the surface along the stroke-plane normal (seal_utils.py:289-461)."""
    TYPE_ID = 1

    def __init__(self, config_path, seal_config):
        # seal_utils.py:304-413.  raw: [N,3] stroke points or a list of strokes; normal: which side of the stroke plane is "up";
        # brushType 'line' | 'curve' (per stroke or one for all); brushPressure / brushDepth / attenuationDistance / attenuationMode.
        super().__init__(seal_config)
        strokes = seal_config["raw"]
        if np.asarray(strokes[0]).ndim == 1:
            strokes = [strokes]
        kinds = seal_config["brushType"]
        if isinstance(kinds, str):
            kinds = [kinds] * len(strokes)
        tris, bounds, border = [], [], []
        for pts, kind in zip(strokes, kinds):
            pts = np.asarray(pts, np.float64).reshape(-1, 3)
            point, normal = mb.plane_best_fit(pts)
            if "normal" in seal_config and normal @ np.asarray(seal_config["normal"], np.float64) < 0:
                normal = -normal
            normal_expand = normal * seal_config["brushPressure"]
            projected = mb.project_points(normal, point, pts)
            if kind == "line":   # box around the stroke pushed +2 and -brushDepth pressures along the normal
                t = mb.box_triangles(mb.oriented_box(np.vstack([pts + 2 * normal_expand, pts - seal_config["brushDepth"] * normal_expand])))
            else:                # smooth sheet through the projected stroke points
                t = mb.fit_curve_mesh(projected, normal_expand, (-seal_config["brushDepth"], 2), seal_config.get("simplifyVoxel", 16))
            tris.append(t)
            v = t.reshape(-1, 3)
            bounds.append(np.stack([v.min(0), v.max(0)]))
            border.append(projected[mb.surface_points_mask(t.astype(np.float32).astype(np.float64), projected)])
        self.map_triangles = torch.as_tensor(np.concatenate(tris, 0).astype(np.float32))
        self.map_test_dir = torch.as_tensor(normal_expand[None].astype(np.float32))  # (from the last stroke, like the reference)
        self.map_data = {
            "force_fill_bound": np.stack(bounds), "map_bound": np.stack(bounds), "normal_expand": normal_expand, "center": point,
            "border_points": np.concatenate(border, 0), "attenuation_distance": seal_config["attenuationDistance"],
            "attenuation_mode": seal_config["attenuationMode"],
        }
        _color_entries(self.map_data, seal_config)
        if "imageConfig" in seal_config:
            ic = seal_config["imageConfig"]
            self.map_data["rgb_light_offset"] = seal_config.get("rgbLightOffset", 0)
            image, alpha = _read_image(ic["path"])
            v_o, v_w, v_h = (np.asarray(ic[k], np.float64) for k in ("o", "w", "h"))
            self.map_data.update(image=image, image_mask=alpha, v_image_norm=mb.plane_best_fit([v_o, v_w, v_h])[1], v_image_o=v_o,
                                 v_image_w=v_w, v_image_h=v_h)

    def _fill_type(self, d, device, keep):
        md = self.map_data
        _fill(d.normal_expand, _np(md["normal_expand"]))
        _fill(d.center, _np(md["center"]))
        d.attenuation_distance = float(_np(md["attenuation_distance"]))
        mode = md["attenuation_mode"]
        if mode not in ("linear", "dry"):
            raise NotImplementedError("attenuation mode %r (the reference implements 'linear' and 'dry' only, seal_utils.py:444-449)" % (mode,))
        d.attenuation_mode = 0 if mode == "linear" else 1
        if mode == "linear":
            border = torch.as_tensor(np.asarray(_np(md["border_points"], (-1, 3)), np.float32)).to(device).contiguous()
            d.n_border, d.border = border.shape[0], border.data_ptr()
            keep.append(border)


class SealAnchorMapper(SealMapper):
    """Control-point tool (seal_utils.py:464-578)."""
    TYPE_ID = 2

    def __init__(self, config_path, seal_config):
        # seal_utils.py:475-520.  raw: [N,3] points of the anchor plane; translation [3]; radius; scale [3].
        super().__init__(seal_config)
        raw = np.asarray(seal_config["raw"], np.float64).reshape(-1, 3)
        v_translation = np.asarray(seal_config["translation"], np.float64)
        v_anchor = raw.mean(0)
        radius = seal_config["radius"]
        point, normal = mb.plane_best_fit(raw)
        v_translated = v_anchor + v_translation
        v_projected = mb.project_points(normal, point, v_translated[None])[0]
        v_offset = v_projected - v_anchor
        v_h = v_projected - v_translated
        sphere = mb.uv_sphere_vertices(radius * 1.1) + v_anchor
        box = mb.oriented_box(np.vstack([sphere, v_anchor + 1.1 * v_translation, sphere - 0.1 * v_translation]))
        self.map_triangles = torch.as_tensor(mb.box_triangles(box).astype(np.float32))
        bounds = np.stack([box.min(0), box.max(0)])
        self.map_data = {
            "force_fill_bound": bounds, "map_bound": bounds, "pose_center": box.mean(0), "pose_radius": np.linalg.norm(v_translation, 2) * 10,
            "v_anchor": v_anchor, "v_offset": v_offset, "v_h": v_h, "len_h": np.linalg.norm(v_h, 2), "radius": radius,
            "scale": seal_config["scale"], "map_source": True,  # (the reference's workaround: keeps every local pre-training point)
        }
        _color_entries(self.map_data, seal_config)

    def _fill_type(self, d, device, keep):
        md = self.map_data
        _fill(d.v_anchor, _np(md["v_anchor"])); _fill(d.v_offset, _np(md["v_offset"])); _fill(d.v_h, _np(md["v_h"]))
        _fill(d.scale, _np(md["scale"]))
        d.len_h, d.radius = float(_np(md["len_h"])), float(_np(md["radius"]))


def get_seal_mapper(config_path, config_dict=None, config_file="seal.json"):
    """Factory with the reference's signature (seal_utils.py:581-592)."""
    if config_dict is None:
        with open(os.path.join(config_path, config_file), "r") as f:
            config_dict = json.load(f)
    kind = config_dict["type"]
    if kind == "bbox":
        return SealBBoxMapper(config_path, config_dict)
    if kind == "brush":
        return SealBrushMapper(config_path, config_dict)
    if kind == "anchor":
        return SealAnchorMapper(config_path, config_dict)
    raise NotImplementedError()


def _color_entries(map_data, seal_config):
    if "hsv" in seal_config:
        map_data["hsv"] = seal_config["hsv"]
    if "rgb" in seal_config:
        map_data["rgb"] = seal_config["rgb"]
        map_data["rgb_light_offset"] = seal_config.get("rgbLightOffset", 0)


def _read_image(path):
    """RGB float image in [0,1] + alpha mask (the reference reads it with cv2, seal_utils.py:389-399)."""
    from PIL import Image
    im = Image.open(path)
    has_alpha = im.mode in ("RGBA", "LA") or "transparency" in im.info
    a = np.asarray(im.convert("RGBA"), np.float32) / 255
    return a[:, :, :3].copy(), (a[:, :, 3].astype(np.float64) if has_alpha else np.ones(a.shape[:2]))


# ---- box helpers ---------------------------------------------------------------------------------------------------------
_BOX_FACES = mb.BOX_FACES


def _box_from_corners(raw):
    """Order the 8 corners of an (oriented) box as (-,-,-), (-,-,+), (-,+,-) ... (+,+,+) in its own frame."""
    raw = np.asarray(raw, np.float64).reshape(-1, 3)
    if raw.shape[0] != 8:
        return mb.oriented_box(raw)
    v = raw[1:] - raw[0]
    far = int(np.argmax((v ** 2).sum(1)))
    best = None
    idx = [i for i in range(7) if i != far]
    for a in range(len(idx)):
        for b in range(a + 1, len(idx)):
            for c in range(b + 1, len(idx)):
                e = v[[idx[a], idx[b], idx[c]]]
                err = np.abs(e.sum(0) - v[far]).sum() + abs(e[0] @ e[1]) + abs(e[0] @ e[2]) + abs(e[1] @ e[2])
                if best is None or err < best[0]:
                    best = (err, e)
    e = best[1]
    o = raw[0]
    return np.array([o + i * e[0] + j * e[1] + k * e[2] for i in (0, 1) for j in (0, 1) for k in (0, 1)])
