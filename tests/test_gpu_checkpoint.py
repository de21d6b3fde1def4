"""Checkpoints in the reference's format (seald_nerf_b200/checkpoint.py; nerf/utils.py:1033-1154): the optimiser / scaler entries
load into torch.optim.Adam / torch.amp.GradScaler built the way main_dnerf.py:129,136 builds them, and a second trainer resumed from
the file continues exactly like the first one."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_save_load_resume_and_torch_optimizer_compat(cuda_dev, tmp_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from seald_nerf_b200 import checkpoint as ckpt
    from seald_nerf_b200.trainer import FusedTrainer
    d = cuda_dev
    ro, rd, ts, gt = bench.make_batches(2, d, 0)
    a = bench.build_scene(d, seed=0)
    ta = FusedTrainer(a, num_rays=4096, max_samples=4096 * 16, lr=1e-2, lr_net=1e-3, perturb=False, init_loss_scale=1024.0, lr_decay_iters=30,
                      ema_decay=0.95)
    for i in range(6):
        ta.train_step(ro[i % 2], rd[i % 2], ts[i % 2], gt[i % 2])
    ta.ema_update()
    path = os.path.join(tmp_path, "ngp_ep0001.pth")
    state = ckpt.save_checkpoint(path, ta, epoch=1)
    assert set(state) >= {"epoch", "global_step", "stats", "mean_count", "mean_density", "model", "optimizer", "scaler", "lr_scheduler", "ema"}
    assert state["global_step"] == 6
    # the keys the reference's Trainer indexes after `self.stats = checkpoint_dict['stats']` (nerf/utils.py:1049,1060,1117)
    assert set(state["stats"]) >= {"loss", "valid_loss", "results", "checkpoints", "best_result"}
    # lr_scheduler: loads into torch's LambdaLR built like main_dnerf.py:134 and reports the learning rate this trainer had reached
    sopt = torch.optim.Adam(bench.build_scene(d, seed=9).get_params(1e-2, 1e-3), betas=(0.9, 0.99), eps=1e-15)
    sched = torch.optim.lr_scheduler.LambdaLR(sopt, lambda it: 0.1 ** min(it / 30, 1))
    sched.load_state_dict(state["lr_scheduler"])
    assert sched.last_epoch == 6 and sched.get_last_lr()[0] == pytest.approx(1e-2 * 0.1 ** (6 / 30), rel=1e-6)
    assert state["optimizer"]["param_groups"][0]["lr"] == pytest.approx(1e-2 * 0.1 ** (6 / 30), rel=1e-6)
    assert state["ema"]["num_updates"] == 1 and len(state["ema"]["shadow_params"]) == len(list(a.parameters()))

    # the reference builds torch.optim.Adam(model.get_params(lr, lr_net), betas=(0.9, 0.99), eps=1e-15) and a GradScaler: both accept the entries
    ref_model = bench.build_scene(d, seed=3)
    opt = torch.optim.Adam(ref_model.get_params(1e-2, 1e-3), betas=(0.9, 0.99), eps=1e-15)
    loaded = torch.load(path, map_location=d)
    opt.load_state_dict(loaded["optimizer"])
    ref_model.load_state_dict(loaded["model"], strict=True)
    emb = ref_model.encoder.embeddings
    assert torch.equal(opt.state[emb]["exp_avg"].reshape(-1), ta.exp_avg[:ta.n_table])
    assert float(opt.state[emb]["step"]) == float(int(ta.step_dev))
    assert [len(g["params"]) for g in opt.param_groups] == [1, 2, 0, 3, 0, 0, 8]
    scaler = torch.amp.GradScaler("cuda")
    scaler.load_state_dict(loaded["scaler"])
    assert scaler.get_scale() == float(ta.loss_scale)

    # resume in a fresh trainer: the next step is the same step
    b = bench.build_scene(d, seed=5)
    tb = FusedTrainer(b, num_rays=4096, max_samples=4096 * 16, lr=1e-2, lr_net=1e-3, perturb=False, init_loss_scale=7.0, lr_decay_iters=30,
                      ema_decay=0.95)
    info = ckpt.load_checkpoint(path, tb)
    assert int(tb.sched_step) == 6 and float(tb.lr_scale) == pytest.approx(float(ta.lr_scale), rel=1e-6)
    assert torch.equal(tb.ema_shadow[:tb.n_params], ta.ema_shadow[:ta.n_params]) and tb.ema_num_updates == 1
    assert not info["missing_keys"] and not info["unexpected_keys"] and info["epoch"] == 1
    assert tb.global_step == 6 and int(tb.step_dev) == int(ta.step_dev) and float(tb.loss_scale) == float(ta.loss_scale)
    assert torch.equal(tb.params[:tb.n_params], ta.params[:ta.n_params]) and torch.equal(tb.exp_avg_sq[:tb.n_params], ta.exp_avg_sq[:ta.n_params])
    assert torch.equal(tb.table16, ta.table16) and torch.equal(b.density_bitfield, a.density_bitfield)
    la = float(ta.train_step(ro[0], rd[0], ts[0], gt[0]))
    lb = float(tb.train_step(ro[0], rd[0], ts[0], gt[0]))
    assert la == pytest.approx(lb, rel=1e-5)
    ta.flush(); tb.flush()
    # (atomic order in the table scatter: entries whose gradient is round-off noise may take opposite Adam signs)
    differ = ((ta.params[:ta.n_table] - tb.params[:tb.n_table]).abs() > 1e-3 * float(ta.params[:ta.n_table].abs().max())).float().mean()
    assert float(differ) < 1e-3
