"""Generate tests/golden/seal.npz: outputs of the REFERENCE's own Seal runtime (SealNeRF/seal_utils.py, imported from
/root/reference on CPU) on the seeded cases of seal_cases().

    python tests/golden/make_seal_golden.py            (needs /root/reference; run in the build container)

The reference module imports json5 / cv2 / pytorch3d / trimesh / scikit-spatial / open3d at module scope; they are only
used by mapper CONSTRUCTION (outside the hot path), so they are stubbed and the mapper objects are created with
`__new__` + the tensors of oracle.seal.make_*_mapper — the runtime functions under test are the reference's, unmodified.
Only outputs are stored; inputs are regenerated from seeds on both sides.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import seal_cases  # noqa: E402


def import_reference_seal_utils(ref="/root/reference"):
    for name in ["json5", "cv2", "pytorch3d", "pytorch3d.structures", "trimesh", "trimesh.creation", "trimesh.primitives", "skspatial",
                 "skspatial.objects", "open3d"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    sys.modules["pytorch3d.structures"].Meshes = object
    sys.modules["pytorch3d"]._C = None
    sys.modules["trimesh.creation"].uv_sphere = None
    sys.modules["trimesh.primitives"].Box = object
    sys.modules["trimesh"].primitives = sys.modules["trimesh.primitives"]
    sys.modules["trimesh"].Trimesh = object
    sys.modules["skspatial.objects"].Plane = object
    sys.path.insert(0, ref)
    import SealNeRF.seal_utils as su
    return su


def to_reference_mapper(su, mp):
    cls = {"bbox": su.SealBBoxMapper, "brush": su.SealBrushMapper, "anchor": su.SealAnchorMapper}[mp["type"]]
    obj = cls.__new__(cls)
    su.SealMapper.__init__(obj, {})
    skip = ("type", "map_triangles", "map_test_dir")
    obj.map_data = {k: (v if isinstance(v, (str, bool)) else (np.array(v) if isinstance(v, np.ndarray) else v)) for k, v in mp.items()
                    if k not in skip}
    if mp["type"] == "anchor":
        obj.map_data["map_source"] = True
    obj.map_triangles = torch.from_numpy(np.array(mp["map_triangles"]))
    obj.map_test_dir = None if mp.get("map_test_dir") is None else torch.from_numpy(np.array(mp["map_test_dir"]))
    obj.map_data_conversion(force=True)
    return obj


def main():
    su = import_reference_seal_utils()
    out = {}
    for name, (mp, pts, dirs, cols) in seal_cases.cases().items():
        ref = to_reference_mapper(su, mp)
        p, d, m = ref.map_to_origin(torch.from_numpy(pts), torch.from_numpy(dirs))
        out[name + "_points"] = p.numpy().astype(np.float32)
        out[name + "_dirs"] = d.numpy().astype(np.float32)
        out[name + "_mask"] = m.numpy()
        if any(k in mp for k in ("hsv", "rgb", "image")):
            mm = m.numpy()
            c = ref.map_color(p[m], d[m], torch.from_numpy(cols[mm]))
            out[name + "_colors"] = c.numpy().astype(np.float32)
        print(name, "mapped", int(m.sum()), "of", len(pts))
    path = os.path.join(HERE, "seal.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
