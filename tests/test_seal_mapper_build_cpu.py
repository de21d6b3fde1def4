"""Mapper construction from GUI edit configs (seald_nerf_b200/SealNeRF/mapper_build.py + the constructors of seal_utils.py;
reference seal_utils.py:168-242, 304-413, 475-520) on CPU: the geometry the reference takes from trimesh / scikit-spatial /
pytorch3d, checked against closed-form cases and against the tensors oracle/seal.py derives for the same edits."""
import numpy as np
import pytest

pytest.importorskip("scipy.spatial")

from oracle import seal as S
from seald_nerf_b200.SealNeRF import mapper_build as mb
from seald_nerf_b200.SealNeRF import seal_utils as su


def _rot(axis, deg):
    a = np.deg2rad(deg)
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K


def _same_box(a, b, tol=1e-9):
    """Two corner sets describe the same box."""
    d = np.linalg.norm(a[:, None] - b[None], axis=-1)
    return d.min(1).max() < tol and d.min(0).max() < tol


def test_oriented_box_recovers_a_rotated_box():
    rng = np.random.default_rng(0)
    R = _rot([1, 2, 3], 37.0)
    half = np.array([0.3, 0.1, 0.05])
    corners = S.box_corners([0.1, -0.2, 0.3], half, R)
    inside = (rng.random((400, 3)) * 2 - 1) * half @ R.T + np.array([0.1, -0.2, 0.3])
    box = mb.oriented_box(np.vstack([corners, inside]))
    assert _same_box(box, corners)
    # the corner order is a consistent (i, j, k) lattice: BOX_FACES triangulates a closed surface around every interior point
    assert mb.points_in_mesh(inside, mb.box_triangles(box)).all()
    assert not mb.points_in_mesh(inside + 1.0, mb.box_triangles(box)).any()
    # a cloud without its corners: the box is no larger than the true one and still contains every point
    box2 = mb.oriented_box(inside)
    e = [np.linalg.norm(box2[4] - box2[0]), np.linalg.norm(box2[2] - box2[0]), np.linalg.norm(box2[1] - box2[0])]
    assert np.prod(e) <= np.prod(2 * half) * (1 + 1e-9)
    assert mb.points_in_mesh(inside * 0.999 + 0.001 * inside.mean(0), mb.box_triangles(box2)).all()


def test_plane_fit_and_projection():
    rng = np.random.default_rng(1)
    n = np.array([1.0, 2.0, -1.0]) / np.sqrt(6)
    u = np.cross(n, [0, 0, 1.0]); u /= np.linalg.norm(u)
    v = np.cross(n, u)
    pts = np.array([0.2, 0.1, 0.0]) + rng.normal(size=(200, 1)) * u + rng.normal(size=(200, 1)) * v + rng.normal(size=(200, 1)) * 1e-6 * n
    c, normal = mb.plane_best_fit(pts)
    assert abs(abs(normal @ n) - 1) < 1e-9 and np.allclose(c, pts.mean(0))
    proj = mb.project_points(normal, c, pts + 0.3 * n)
    assert np.abs((proj - c) @ normal).max() < 1e-12


def test_bbox_mapper_from_a_point_cloud_equals_the_one_from_its_corners():
    rng = np.random.default_rng(2)
    R = S.rot_y(20.0)
    corners = S.box_corners([0.0, 0.15, 0.0], [0.15, 0.1, 0.2], R)
    cloud = np.vstack([corners, (rng.random((100, 3)) * 2 - 1) * [0.15, 0.1, 0.2] @ R.T + [0.0, 0.15, 0.0]])
    T = np.eye(4); T[:3, :3] = S.rot_y(30.0); T[:3, 3] = [0.2, 0.0, 0.05]
    cfg = {"type": "bbox", "transform": T.tolist(), "scale": [1.0, 1.2, 0.8], "boundType": "both", "hsv": [0.1, 0, 0], "mapSource": [0.5, 0.5, 0.5]}
    a = su.get_seal_mapper("", dict(cfg, raw=corners.tolist()))
    b = su.get_seal_mapper("", dict(cfg, raw=cloud.tolist()))
    for k in ("force_fill_bound", "map_bound", "transform", "rotation", "scale", "center", "pose_center", "empty_bound"):
        np.testing.assert_allclose(np.asarray(a.map_data[k], np.float64), np.asarray(b.map_data[k], np.float64), rtol=0, atol=1e-9, err_msg=k)
    q = (rng.random((4000, 3)) * 2 - 1) * 0.6
    assert np.array_equal(S.points_in_mesh(q, a.map_triangles.numpy()), S.points_in_mesh(q, b.map_triangles.numpy()))


def test_brush_line_mapper_matches_the_oracle_tensors():
    """Two line strokes on planes x = const: the stroke boxes, their bounds, the pushed normal and the rim points derived from the
    config equal what oracle/seal.py's make_brush_mapper (the checker of the CUDA mapping) builds for the same edit."""
    ref = S.make_brush_mapper(mode="linear", pressure=0.02, depth=0.6, attenuation=0.02)
    strokes = []
    for k in range(2):
        c = np.array([0.12, 0.25 - 0.35 * k, 0.02 + 0.05 * k])
        u = np.linspace(-1, 1, 12)
        rim = [np.array([0, a * 0.12, s * 0.05]) for a in u for s in (-1, 1)] + [np.array([0, s * 0.12, a * 0.05]) for a in u for s in (-1, 1)]
        inner = [np.array([0, a * 0.06, b * 0.02]) for a in (-1, 0, 1) for b in (-1, 1)]
        strokes.append((c + np.array(rim + inner)).tolist())
    m = su.get_seal_mapper("", {"type": "brush", "raw": strokes, "normal": [1, 0, 0], "brushType": "line", "brushDepth": 0.6, "brushPressure": 0.02,
                                "attenuationDistance": 0.02, "attenuationMode": "linear", "rgb": [1, 0, 0]})
    np.testing.assert_allclose(m.map_data["map_bound"], ref["map_bound"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(m.map_data["normal_expand"], ref["normal_expand"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(m.map_test_dir.numpy(), ref["map_test_dir"], rtol=0, atol=1e-7)
    rng = np.random.default_rng(3)
    q = np.array([0.12, 0.05, 0.05]) + (rng.random((6000, 3)) * 2 - 1) * [0.06, 0.45, 0.12]
    assert np.array_equal(S.points_in_mesh(q, m.map_triangles.numpy(), ref["map_test_dir"]), S.points_in_mesh(q, ref["map_triangles"], ref["map_test_dir"]))
    # border points: exactly the rim of each stroke rectangle (interior stroke points are not on the surface)
    ours = np.asarray(m.map_data["border_points"], np.float64)
    d = np.linalg.norm(ours[:, None] - ref["border_points"][None].astype(np.float64), axis=-1)
    assert ours.shape[0] == ref["border_points"].shape[0] and d.min(1).max() < 1e-6 and d.min(0).max() < 1e-6
    assert m.map_data["rgb"] == [1, 0, 0] and m.map_data["attenuation_mode"] == "linear"


def test_brush_curve_mesh_encloses_the_stroke_sheet():
    rng = np.random.default_rng(4)
    uv = rng.random((1500, 2)) * [0.3, 0.2]                              # a dense GUI stroke
    pts = np.stack([uv[:, 0], uv[:, 1], 0.02 * np.sin(8 * uv[:, 0])], 1)   # a wavy sheet
    m = su.get_seal_mapper("", {"type": "brush", "raw": pts.tolist(), "normal": [0, 0, 1], "brushType": "curve", "brushDepth": 1.0,
                                "brushPressure": 0.05, "attenuationDistance": 0.05, "attenuationMode": "dry", "simplifyVoxel": 16})
    tris = m.map_triangles.numpy()
    assert tris.ndim == 3 and tris.shape[1:] == (3, 3) and tris.shape[0] > 50
    lo, hi = np.asarray(m.map_data["map_bound"])[0]
    assert lo[2] < -0.04 and hi[2] > 0.09  # -depth * pressure below, +2 pressures above the fitted plane
    core = np.stack([rng.random(400) * 0.2 + 0.05, rng.random(400) * 0.1 + 0.05, rng.random(400) * 0.03], 1)
    assert S.points_in_mesh(core, tris, m.map_test_dir.numpy()).mean() > 0.9  # (a triangle soup: small gaps between neighbourhoods remain)


def test_anchor_mapper_vectors_match_the_oracle():
    ref = S.make_anchor_mapper()
    raw = [[0.0, 0.05, 0.05], [0.1, 0.05, 0.05], [0.05, -0.1, 0.05]]   # plane z = 0.05, centroid = the anchor (0.05, 0, 0.05)
    m = su.get_seal_mapper("", {"type": "anchor", "raw": raw, "translation": [0.04, 0.0, 0.12], "radius": 0.1, "scale": [1, 1, 1]})
    for k in ("v_anchor", "v_offset", "v_h"):
        np.testing.assert_allclose(np.asarray(m.map_data[k], np.float64), ref[k], rtol=0, atol=1e-7, err_msg=k)
    assert abs(m.map_data["len_h"] - ref["len_h"]) < 1e-9 and m.map_data["map_source"] is True
    # the affected region (sphere of 1.1 r around the anchor, swept along the translation) lies inside the map box
    sph = mb.uv_sphere_vertices(0.1 * 1.05) + np.array([0.05, 0.0, 0.05])
    assert S.points_in_mesh(sph, m.map_triangles.numpy()).all()
    lo, hi = np.asarray(m.map_data["map_bound"])
    assert (lo <= sph.min(0)).all() and (hi >= sph.max(0)).all()
