"""Golden list of parameter / buffer names, shapes and dtypes of the REFERENCE's own networks (dnerf/network.py and
SealDNeRF/network.py imported from /root/reference with this package's op shims aliased in and the absent third-party modules
stubbed), for the checkpoint-compatibility test: a reference checkpoint's `model` state dict must load into our modules.

    python tests/golden/make_state_dict_golden.py       (needs /root/reference; writes tests/golden/state_dict_keys.json)
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import seald_nerf_b200  # noqa: E402

seald_nerf_b200.install_aliases()  # raymarching / gridencoder / ffmlp / freqencoder / shencoder / encoding / activation -> our shims


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Any:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, k):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


for name in ("trimesh", "trimesh.creation", "trimesh.primitives", "json5", "pytorch3d", "pytorch3d.structures", "skspatial", "skspatial.objects",
             "open3d", "mcubes", "tensorboardX", "lpips", "torch_ema", "torchmetrics", "torchmetrics.functional", "imageio", "dearpygui",
             "dearpygui.dearpygui", "matplotlib", "matplotlib.pyplot", "sklearn.decomposition"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            stub(name, __getattr__=lambda k: _Any())
sys.path.insert(0, "/root/reference")


def custom_meshgrid(*args):
    return torch.meshgrid(*args, indexing="ij")


# the reference's */utils.py drag in the trainer stack; the networks only need custom_meshgrid from them
stub("nerf.utils", custom_meshgrid=custom_meshgrid, Trainer=object)
stub("dnerf.utils", custom_meshgrid=custom_meshgrid, Trainer=object)
stub("SealDNeRF.utils", custom_meshgrid=custom_meshgrid)

out = {}
from dnerf.network import NeRFNetwork as RefDNeRF  # noqa: E402

torch.manual_seed(0)
net = RefDNeRF(encoding="hashgrid", bound=1, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=10)
out["dnerf"] = {k: [list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()}
try:
    from SealDNeRF.network import NeRFNetwork as RefSeal  # noqa: E402
    net2 = RefSeal(encoding="hashgrid", bound=1, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=10)
    out["seald"] = {k: [list(v.shape), str(v.dtype)] for k, v in net2.state_dict().items()}
except Exception as e:  # the SealD network subclasses the teacher renderer, which needs more of the stack
    out["seald_error"] = repr(e)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json"), "w"), indent=1, sort_keys=True)
print({k: (len(v) if isinstance(v, dict) else v) for k, v in out.items()})
