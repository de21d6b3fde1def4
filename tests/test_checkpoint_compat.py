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


def test_optimizer_state_dict_host_logic_loads_into_torch_adam():
    """checkpoint.optimizer_state_dict / scaler_state_dict only re-label views of the trainer's flat buffers: exercised on CPU with a
    stand-in trainer (the real one needs a GPU; tests/test_gpu_checkpoint.py covers it end to end).  The result must load into
    torch.optim.Adam built over the reference's seven parameter groups (dnerf/network.py:260-272) and a GradScaler."""
    from types import SimpleNamespace
    from seald_nerf_b200 import checkpoint as ckpt
    net = _ours(False)
    n_table = net.encoder.embeddings.numel()
    weights = net.mlp_weights()
    n = n_table + sum(w.numel() for w in weights)
    g = torch.Generator().manual_seed(2)
    tr = SimpleNamespace(model=net, lr=1e-2, lr_net=1e-3, betas=(0.9, 0.99), eps=1e-15, n_table=n_table, n_table_pad=n_table,
                         exp_avg=torch.rand(n, generator=g), exp_avg_sq=torch.rand(n, generator=g), step_dev=torch.tensor([17]),
                         loss_scale=torch.tensor([4096.0]), growth_tracker=torch.tensor([5]), growth_interval=2000,
                         lr_scale=torch.tensor([1.0]), sched_step=torch.tensor([17]))
    sd = ckpt.optimizer_state_dict(tr)
    assert [len(pg["params"]) for pg in sd["param_groups"]] == [1, 2, 0, 3, 0, 0, 8]
    assert [pg["lr"] for pg in sd["param_groups"]] == [1e-2, 1e-3, 1e-2, 1e-3, 1e-2, 1e-2, 1e-3]
    opt = torch.optim.Adam(net.get_params(1e-2, 1e-3), betas=(0.9, 0.99), eps=1e-15)
    opt.load_state_dict(sd)
    assert torch.equal(opt.state[net.encoder.embeddings]["exp_avg"].reshape(-1), tr.exp_avg[:n_table])
    o = n_table
    for w in weights:  # flat order of the trainer: deformation net, sigma net, colour net
        assert torch.equal(opt.state[w]["exp_avg_sq"].reshape(-1), tr.exp_avg_sq[o:o + w.numel()])
        assert float(opt.state[w]["step"]) == 17.0
        o += w.numel()
    # LambdaLR state: loads into torch's scheduler (main_dnerf.py:134) and reports the device-side factor
    tr.lr_scale = torch.tensor([0.5])
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 0.1 ** min(it / 100, 1))
    sched.load_state_dict(ckpt.lr_scheduler_state_dict(tr))
    assert sched.last_epoch == 17 and sched.get_last_lr() == [0.5 * g["lr"] for g in net.get_params(1e-2, 1e-3)]
    assert ckpt.optimizer_state_dict(tr)["param_groups"][0]["lr"] == 0.5e-2
    assert set(ckpt.default_stats()) == {"loss", "valid_loss", "results", "checkpoints", "best_result"}
    sc = ckpt.scaler_state_dict(tr)
    assert sc == {"scale": 4096.0, "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000, "_growth_tracker": 5}
