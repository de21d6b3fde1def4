"""Checkpoints of a FusedTrainer in the REFERENCE's on-disk format (Trainer.save_checkpoint / load_checkpoint,
nerf/utils.py:1033-1154): a `torch.save`d dict

    {epoch, global_step, stats, mean_count, mean_density, model: state_dict, [optimizer, lr_scheduler, scaler, ema]}

with the reference's parameter names (tests/test_checkpoint_compat.py), `optimizer` a torch.optim.Adam state dict over the
reference's parameter groups (`NeRFNetwork.get_params`: table, sigma net, colour net, deformation net; empty groups for the
parameter-free encoders) and `scaler` a torch GradScaler state dict — so a run can move between the reference trainer and
this one in either direction: `lr_scheduler` is a LambdaLR state dict (last_epoch = the device-side scheduler counter, the
groups' `lr` carry the current factor, `initial_lr` the base rate), `ema` a torch_ema.ExponentialMovingAverage state dict over
`model.parameters()`, `stats` has the keys the reference's Trainer indexes after loading (nerf/utils.py:1049,1060,1117).  The flat
fp32 buffers of the trainer ARE the parameters / Adam moments; this module only re-labels views of them.  Data parallel with a
sharded optimiser: the moments of the table shards other ranks own are gathered first (every rank writes a complete file).
"""
import torch


def _param_slots(trainer):
    """parameter object -> (offset, numel) in the trainer's flat buffers"""
    m = trainer.model
    slots = {id(m.encoder.embeddings): (0, trainer.n_table)}
    o = trainer.n_table_pad
    for w in m.mlp_weights():
        slots[id(w)] = (o, w.numel())
        o += w.numel()
    return slots


def optimizer_state_dict(trainer):
    """torch.optim.Adam.state_dict() equivalent of the trainer's optimiser state."""
    groups = trainer.model.get_params(trainer.lr, trainer.lr_net)
    slots = _param_slots(trainer)
    step = float(int(trainer.step_dev))
    factor = float(trainer.lr_scale)
    state, param_groups, idx = {}, [], 0
    for g in groups:
        ids = []
        for p in g["params"]:
            o, k = slots[id(p)]
            state[idx] = {"step": torch.tensor(step), "exp_avg": trainer.exp_avg[o:o + k].view_as(p).clone(),
                          "exp_avg_sq": trainer.exp_avg_sq[o:o + k].view_as(p).clone()}
            ids.append(idx)
            idx += 1
        param_groups.append({"lr": g["lr"] * factor, "initial_lr": g["lr"], "betas": tuple(trainer.betas), "eps": trainer.eps, "weight_decay": 0, "amsgrad": False,
                             "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                             "decoupled_weight_decay": False, "params": ids})
    return {"state": state, "param_groups": param_groups}


def scaler_state_dict(trainer):
    """torch.amp.GradScaler.state_dict() equivalent."""
    return {"scale": float(trainer.loss_scale), "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": trainer.growth_interval,
            "_growth_tracker": int(trainer.growth_tracker)}


def lr_scheduler_state_dict(trainer):
    """torch.optim.lr_scheduler.LambdaLR.state_dict() equivalent (the lambda itself is not picklable and saved as None, as torch does)."""
    base = [g["lr"] for g in trainer.model.get_params(trainer.lr, trainer.lr_net)]
    factor, epoch = float(trainer.lr_scale), int(trainer.sched_step)
    return {"base_lrs": base, "last_epoch": epoch, "_step_count": epoch + 1, "_get_lr_called_within_step": False, "_is_initial": False,
            "_last_lr": [b * factor for b in base], "lr_lambdas": [None] * len(base)}


def ema_state_dict(trainer):
    """torch_ema.ExponentialMovingAverage.state_dict() equivalent over model.parameters()."""
    slots = _param_slots(trainer)
    shadow = []
    for p in trainer.model.parameters():
        o, k = slots[id(p)]
        shadow.append(trainer.ema_shadow[o:o + k].view_as(p).clone())
    return {"decay": trainer.ema_decay, "num_updates": trainer.ema_num_updates, "shadow_params": shadow, "collected_params": None}


def default_stats():
    """The `stats` dict of the reference's Trainer (nerf/utils.py:341-347)."""
    return {"loss": [], "valid_loss": [], "results": [], "checkpoints": [], "best_result": None}


def save_checkpoint(path, trainer, epoch=0, stats=None, full=True):
    trainer.sync_params(moments=full)  # deferred table update applied; sharded master copy (and Adam moments) gathered
    m = trainer.model
    st = default_stats()
    st.update(stats or {})
    state = {"epoch": int(epoch), "global_step": int(trainer.global_step), "stats": st,
             "mean_count": m.mean_count, "mean_density": m.mean_density}
    if full:
        state["optimizer"] = optimizer_state_dict(trainer)
        state["lr_scheduler"] = lr_scheduler_state_dict(trainer)
        state["scaler"] = scaler_state_dict(trainer)
        if trainer.ema_shadow is not None:
            state["ema"] = ema_state_dict(trainer)
    state["model"] = m.state_dict()
    torch.save(state, path)
    return state


def load_checkpoint(path, trainer, model_only=False):
    """Load a checkpoint written by save_checkpoint() or by the reference's Trainer (same format)."""
    ck = torch.load(path, map_location=trainer.device)
    m = trainer.model
    sd = ck["model"] if "model" in ck else ck
    missing, unexpected = m.load_state_dict(sd, strict=False)  # the parameters are views of the flat buffer: copied in place
    if "mean_count" in ck:
        m.mean_count = ck["mean_count"]
    if "mean_density" in ck:
        m.mean_density = ck["mean_density"]
    if not model_only:
        trainer.global_step = int(ck.get("global_step", 0))
        opt = ck.get("optimizer")
        if opt is not None:
            groups = m.get_params(trainer.lr, trainer.lr_net)
            slots = _param_slots(trainer)
            idx, step = 0, 0
            for g, pg in zip(groups, opt["param_groups"]):
                for p, pid in zip(g["params"], pg["params"]):
                    st = opt["state"].get(pid)
                    o, k = slots[id(p)]
                    if st is not None:
                        trainer.exp_avg[o:o + k].copy_(st["exp_avg"].reshape(-1))
                        trainer.exp_avg_sq[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
                        step = max(step, int(float(st["step"])))
                    idx += 1
            trainer.step_dev.fill_(step)
            # the learning rate the run had reached: group lr / initial_lr (torch's schedulers rewrite group["lr"] in place)
            for pg in opt["param_groups"]:
                if pg.get("params") and pg.get("initial_lr"):
                    trainer.lr_scale.fill_(float(pg["lr"]) / float(pg["initial_lr"]))
                    break
        sch = ck.get("lr_scheduler")
        if sch is not None:
            epoch = int(sch.get("last_epoch", 0))
            trainer.sched_step.fill_(epoch)
            if sch.get("_last_lr") and sch.get("base_lrs") and float(sch["base_lrs"][0]) != 0.0:
                trainer.lr_scale.fill_(float(sch["_last_lr"][0]) / float(sch["base_lrs"][0]))
            elif trainer.lr_decay_iters:
                trainer.lr_scale.fill_(0.1 ** min(epoch / trainer.lr_decay_iters, 1))
        ema = ck.get("ema")
        if ema is not None and trainer.ema_shadow is not None:
            slots = _param_slots(trainer)
            for p, sp in zip(m.parameters(), ema["shadow_params"]):
                o, k = slots[id(p)]
                trainer.ema_shadow[o:o + k].copy_(sp.reshape(-1))
            trainer.ema_num_updates = int(ema.get("num_updates") or 0)
        sc = ck.get("scaler")
        if sc:
            trainer.loss_scale.fill_(float(sc["scale"]))
            trainer.growth_tracker.fill_(int(sc.get("_growth_tracker", 0)))
    # re-stage everything derived from the parameters / occupancy grid
    trainer.table16.copy_(m.encoder.embeddings.data)
    if trainer.dp_mode == "sharded":
        trainer.shard16.copy_(trainer.table16_pad[trainer.rank * trainer.shard_len:(trainer.rank + 1) * trainer.shard_len])
    trainer.hw.refresh(trainer.weight_views)
    trainer.pending[0] = 1  # no deferred update pending
    trainer.refresh_occupancy()
    return {"missing_keys": missing, "unexpected_keys": unexpected, "epoch": ck.get("epoch", 0)}
