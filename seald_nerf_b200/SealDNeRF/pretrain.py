"""Seal LOCAL PRE-TRAINING on the device.

Reference: `SealDNeRF/utils.py:308-335` (`sample_points`), `:386-562` (`init_pretraining`: the three point sets and their teacher
labels), `SealNeRF/trainer.py:363-462` (`pretrain_one_epoch` / `pretrain_part` / `pretrain_step`).  Before the distillation steps
start, the student's hash table is fitted directly — in 3-D, no rendering — to the teacher field at

  local        points on a `local_point_step` lattice inside the mapper's `force_fill_bound` that the proxy mapping moves; the label is
               the teacher evaluated at the MAPPED point, its colour passed through `map_color`;
  surrounding  lattice points of the bound grown by `surrounding_bounds_extend` that the mapping does NOT move; label = teacher there;
  global       a coarse lattice over the whole training box, unmapped points only.

`SealPretrainer.init_pretraining` builds those sets with this package's kernels (`SealMapper.map_to_origin` = csrc/seal.cu, the
teacher through the fused field kernels, in bounded chunks — the reference evaluates each set in one call); `pretrain_part` walks a
set in `batch_size` slices and hands each to `FusedTrainer.pretrain_step` (every MLP frozen, L1 on sigma and colour, loss-scaled Adam
on the table at `lr`).  Same names / arguments / dict layout (`pretraining_data[kind] = {points, dirs, sigma, color, steps}`) as the
reference so its GUI / training loop code reads the same.  D-NeRF specifics the reference leaves implicit: every teacher / student
query carries the edited `time_frame` (its `global` branch calls the teacher without a time stamp, which only the static Seal-3D
backbone accepts).
"""
import math

import torch


def sample_points(bounds, point_step=0.005, angle_step=45):
    """Lattice points per `point_step` inside bounds [B,2,3] or [2,3] and the Euler-angle lattice of view directions
    (SealDNeRF/utils.py:308-335).  Directions: the vector (1 - 1e-5, 0, 0) rotated by the extrinsic x-y-z Euler angles
    (a, b, c) on a lattice of `angle_step` degrees = Rz(c) Ry(b) Rx(a) v = v (cos b cos c, cos b sin c, -sin b); float64 like
    scipy's Rotation.apply."""
    bounds = torch.as_tensor(bounds, dtype=torch.float32)
    if bounds.ndim == 2:
        bounds = bounds[None]
    pts, drs = [], []
    ang = torch.arange(0, 360, step=angle_step)
    r_x, r_y, r_z = torch.meshgrid(ang, ang, ang, indexing="ij")
    e = torch.stack([r_x, r_y, r_z], dim=-1).reshape(-1, 3).double() * (math.pi / 180.0)
    v = 1 - 1e-5
    d = torch.stack([v * torch.cos(e[:, 1]) * torch.cos(e[:, 2]), v * torch.cos(e[:, 1]) * torch.sin(e[:, 2]), -v * torch.sin(e[:, 1])], dim=-1)
    for i in range(bounds.shape[0]):
        lo, hi = bounds[i]
        X, Y, Z = torch.meshgrid(torch.arange(lo[0], hi[0], step=point_step), torch.arange(lo[1], hi[1], step=point_step),
                                 torch.arange(lo[2], hi[2], step=point_step), indexing="ij")
        pts.append(torch.stack([X, Y, Z], dim=-1).reshape(-1, 3))
        drs.append(d)
    return torch.concat(pts), torch.concat(drs)


class SealPretrainer:
    """init_pretraining / pretrain_one_epoch of the reference's Seal trainer over a FusedTrainer (the student) and a teacher
    network with a `seal_mapper` (SealNeRFTeacherRenderer.init_mapper)."""

    CHUNK = 1 << 20  # teacher queries per launch sequence

    def __init__(self, trainer, teacher_model):
        self.trainer = trainer
        self.model = trainer.model
        self.teacher_model = teacher_model
        self.device = trainer.params.device
        self.pretraining_epochs = 0
        self.pretraining_batch_size = 4096
        self.pretraining_lr = 0.07
        self.pretraining_data = {}
        self.is_pretraining = False
        self.local_step = 0
        self.time_frame = None
        self.last_losses = {}

    sample_points = staticmethod(sample_points)

    # ---- teacher labels ------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _teacher(self, points, dirs, time_frame, map_color=False):
        sig, col = [], []
        mapper = self.teacher_model.seal_mapper
        for s in range(0, points.shape[0], self.CHUNK):
            p, d = points[s:s + self.CHUNK], dirs[s:s + self.CHUNK]
            with torch.autocast(device_type="cuda", dtype=torch.float16):
                out = self.teacher_model(p, d, time_frame)
            c = out[1].float()
            if map_color:
                c = mapper.map_color(p, d, c)
            sig.append(out[0].float().reshape(-1))
            col.append(c.reshape(-1, 3))
        if not sig:
            return torch.zeros(0, device=self.device), torch.zeros(0, 3, device=self.device)
        return torch.cat(sig), torch.cat(col)

    @staticmethod
    def _steps(n, batch):
        steps = list(range(0, n, batch))
        if not steps or steps[-1] != n:
            steps.append(n)
        return steps

    def _x_dirs(self, points):
        return torch.zeros_like(points) + torch.tensor([1.0, 0.0, 0.0], device=points.device)

    def _entry(self, points, dir_pool, sigma, color):
        n = points.shape[0]
        dirs = dir_pool[torch.randint(dir_pool.shape[0], (n,), device=self.device)]
        return dirs, {"points": points, "dirs": dirs, "sigma": sigma, "color": color, "steps": self._steps(n, self.pretraining_batch_size)}

    # ---- SealDNeRF/utils.py:386-562 --------------------------------------------------------------------------------------------
    @torch.no_grad()
    def init_pretraining(self, time_frame, epochs=0, batch_size=4096, lr=0.07, local_point_step=0.001, local_angle_step=45,
                         surrounding_point_step=0.01, surrounding_angle_step=45, surrounding_bounds_extend=0.2, global_point_step=0.05,
                         global_angle_step=45, no_debug=True):
        """Call once the teacher's seal_mapper is initialised.  The view direction of a pre-training sample is drawn once here
        (`randint` into the direction lattice) like the reference does, so an epoch is deterministic."""
        self.time_frame = float(time_frame)
        tf = torch.tensor([[self.time_frame]], dtype=torch.float32, device=self.device)
        self.pretraining_epochs, self.pretraining_batch_size, self.pretraining_lr = epochs, batch_size, lr
        self.pretraining_data = {}
        if epochs <= 0:
            return
        mapper = self.teacher_model.seal_mapper
        dev = self.device
        aabb = self.model.aabb_train.detach().float().cpu()
        fill = torch.as_tensor(mapper.map_data["force_fill_bound"], dtype=torch.float32).clone()

        if local_point_step > 0:
            pts, drs = self.sample_points(fill, local_point_step, local_angle_step)
            pts, drs = pts.to(dev, torch.float32), drs.to(dev, torch.float32)
            mp, md, mask = mapper.map_to_origin(pts, self._x_dirs(pts))
            if "map_source" in mapper.map_data:  # with a map source every point of the fill bound is kept (:416-417)
                mask[:] = True
            pts, mp, md = pts[mask], mp[mask], md[mask]
            # labels: the teacher at the MAPPED points, colours through map_color (:429-438)
            sigma, color = self._teacher(mp, md, tf, map_color=True)
            _, self.pretraining_data["local"] = self._entry(pts, drs, sigma, color)
            self.is_pretraining = True

        if surrounding_point_step > 0:
            b = fill.clone()  # (the reference grows map_data['force_fill_bound'] in place; the stored bound is left alone here)
            if b.ndim == 2:
                b = b[None]
            b[:, 0] = torch.max(b[:, 0] - surrounding_bounds_extend, aabb[:3])
            b[:, 1] = torch.min(b[:, 1] + surrounding_bounds_extend, aabb[3:])
            pts, drs = self.sample_points(b, surrounding_point_step, surrounding_angle_step)
            pts, drs = pts.to(dev, torch.float32), drs.to(dev, torch.float32)
            _, _, mask = mapper.map_to_origin(pts, self._x_dirs(pts))
            pts = pts[~mask]  # only points the edit leaves alone
            dirs = drs[torch.randint(drs.shape[0], (pts.shape[0],), device=dev)]
            sigma, color = self._teacher(pts, dirs, tf)
            self.pretraining_data["surrounding"] = {"points": pts, "dirs": dirs, "sigma": sigma, "color": color,
                                                    "steps": self._steps(pts.shape[0], batch_size)}

        if global_point_step > 0:
            pts, drs = self.sample_points(aabb.view(2, 3), global_point_step, global_angle_step)
            pts, drs = pts.to(dev, torch.float32), drs.to(dev, torch.float32)
            _, _, mask = mapper.map_to_origin(pts, self._x_dirs(pts))
            pts = pts[~mask]
            dirs = drs[torch.randint(drs.shape[0], (pts.shape[0],), device=dev)]
            sigma, color = self._teacher(pts, dirs, tf)
            self.pretraining_data["global"] = {"points": pts, "dirs": dirs, "sigma": sigma, "color": color,
                                               "steps": self._steps(pts.shape[0], batch_size)}

    # ---- SealNeRF/trainer.py:363-462 ----------------------------------------------------------------------------------------------
    def pretrain_part(self, source_type, silent=True):
        """One pass over a point set in `pretraining_batch_size` slices; returns the mean loss (one host read at the end)."""
        src = self.pretraining_data[source_type]
        steps = src["steps"]
        total = torch.zeros(1, device=self.device)
        n = 0
        for i in range(len(steps) - 1):
            a, b = steps[i], steps[i + 1]
            if b <= a:
                continue
            self.local_step += 1
            self.trainer.pretrain_step(src["points"][a:b], src["dirs"][a:b], src["sigma"][a:b], src["color"][a:b], self.time_frame,
                                       lr=self.pretraining_lr)
            total += self.trainer.loss
            n += 1
        mean = float(total) / max(n, 1)
        self.last_losses[source_type] = mean
        return mean

    def pretrain_one_epoch(self, silent=True):
        """Every point set once, then the EMA update of the reference's epoch end (:392-393)."""
        if not self.model.density_bitfield_hacked:
            self.model.hack_bitfield()
        self.local_step = 0
        for key in self.pretraining_data.keys():
            self.pretrain_part(key, silent)
        if getattr(self.trainer, "ema_decay", None):
            self.trainer.ema_update()
        return dict(self.last_losses)
