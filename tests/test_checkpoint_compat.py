"""CPU: parameter / buffer names, shapes and dtypes of our drop-in networks equal those of the REFERENCE's own networks
(tests/golden/state_dict_keys.json, produced by tests/golden/make_state_dict_golden.py from /root/reference/dnerf/network.py and
SealDNeRF/network.py), so the `model` entry of a reference checkpoint (nerf/utils.py:1033-1154: torch.save dict) loads with
strict=True — and a state dict saved here loads into the reference."""
import json
import os

import torch

from conftest import GOLDEN


def _ours(seald):
    if seald:
        from seald_nerf_b200.SealDNeRF.network import NeRFNetwork
    else:
        from seald_nerf_b200.dnerf.network import NeRFNetwork
    torch.manual_seed(0)
    return NeRFNetwork(encoding="hashgrid", bound=1, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=10)


def test_state_dict_layout_equals_reference():
    gold = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    for name, seald in (("dnerf", False), ("seald", True)):
        sd = _ours(seald).state_dict()
        mine = {k: [list(v.shape), str(v.dtype)] for k, v in sd.items()}
        assert sorted(mine) == sorted(gold[name]), (name, sorted(set(mine) ^ set(gold[name])))
        for k in mine:
            assert mine[k] == gold[name][k], (name, k, mine[k], gold[name][k])


def test_reference_style_checkpoint_roundtrip(tmp_path):
    """A checkpoint dict shaped like the reference's Trainer.save_checkpoint (nerf/utils.py:1033-1075) loads strictly."""
    gold = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))["dnerf"]
    g = torch.Generator().manual_seed(1)
    model_sd = {}
    for k, (shape, dtype) in gold.items():
        dt = getattr(torch, dtype.replace("torch.", ""))
        t = torch.rand(shape, generator=g) if dt.is_floating_point else torch.randint(0, 255, shape, generator=g).to(dt)
        model_sd[k] = t.to(dt)
    ckpt = {"epoch": 3, "global_step": 1234, "stats": {}, "mean_count": 4096, "mean_density": 0.5, "model": model_sd}
    path = os.path.join(tmp_path, "ngp_ep0003.pth")
    torch.save(ckpt, path)
    loaded = torch.load(path, map_location="cpu")
    net = _ours(False)
    missing, unexpected = net.load_state_dict(loaded["model"], strict=True)
    assert not missing and not unexpected
    for k, v in net.state_dict().items():
        assert torch.equal(v, model_sd[k]), k
