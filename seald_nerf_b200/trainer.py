"""FusedTrainer — the D-NeRF / SealD-student training step as one graph-captured kernel sequence.

Replaces, for `cuda_ray=True` training, the reference's `Trainer.train_step` + `scaler.scale(loss).backward()` +
`scaler.step(optimizer)` + `scaler.update()` (dnerf/utils.py:38-115, nerf/utils.py:879-886, Adam as configured in
main_dnerf.py:129) with:

    [pixel sampling + get_rays + target gather from a resident dataset]  ->  march (AABB fused, warp-compacted)  ->  deform MLP
    ->  grid encoder  ->  sigma/colour heads  ->  composite + MSE/background blend + composite bwd (one kernel)  ->  heads bwd
    ->  grid scatter + input grad  ->  deform bwd  ->  weight-gradient GEMMs  ->  [gradient exchange over the GPUs, fused with
    the optimiser: csrc/dp_fused.cu]  ->  overflow check + Adam + fp16 refresh + loss-scale update

Every buffer is preallocated, no host synchronisation happens inside a step (the sample count stays on the device), and the
whole step — exchange kernels and cross-rank barriers included — is ONE CUDA graph per rank.  Parameters stay the model's own
nn.Parameters (re-pointed into one flat fp32 buffer), so checkpoints and the drop-in `NeRFNetwork.forward` see the trained values
(call flush() / sync_params() first when the table pass of the optimiser is deferred, i.e. in the fused data-parallel mode).

Environment switches (measurement only; defaults follow the measurements recorded in DESIGN.md §3): SEALD_DP_MODE, SEALD_DP_MULTICAST,
SEALD_DP_MULTICAST_REDUCE, SEALD_DEFER, SEALD_FORK, SEALD_ADAM_BLOCKS, SEALD_GRID_AGG_LEVELS, SEALD_GRID_SCATTER_PPC.
"""
import math

import ctypes as C
import os

import torch

from . import _lib
from . import field as F
from . import parallel
from ._lib import ptr


class FusedTrainer:
    def __init__(self, model, num_rays=4096, max_samples=None, lr=1e-2, lr_net=None, betas=(0.9, 0.99), eps=1e-15, dt_gamma=0.0,
                 max_steps=1024, T_thresh=1e-4, perturb=True, init_loss_scale=65536.0, growth_interval=2000, train_deform=True,
                 use_graph=True, world_size=1, process_group=None, device=None, fuse_composite=True, shard_optimizer=True, dp_mode=None,
                 defer_table_update=True, lr_decay_iters=None, ema_decay=None):
        self.model = model
        self.device = device or model.encoder.embeddings.device
        if self.device.type != "cuda":
            raise RuntimeError("FusedTrainer needs a CUDA device (no CPU fallback)")
        self.N = int(num_rays)
        self.cfg = model._field_cfg
        self.cfg.density_scale = float(model.density_scale)
        self.M = int(max_samples) if max_samples else self.N * 64
        self.M += (128 - self.M % 128) % 128
        self.lr, self.lr_net = float(lr), float(lr if lr_net is None else lr_net)
        self.betas, self.eps = betas, float(eps)
        self.dt_gamma, self.max_steps, self.T_thresh, self.perturb = float(dt_gamma), int(max_steps), float(T_thresh), bool(perturb)
        self.train_deform = bool(train_deform)
        self.growth_interval = int(growth_interval)
        # LambdaLR(0.1 ** min(iter / iters, 1)) of main_dnerf.py:134, stepped every iteration: the factor lives on the device
        # (self.lr_scale), is advanced by the optimiser-tail kernel and read by both Adam passes, so the captured step follows it
        self.lr_decay_iters = int(lr_decay_iters) if lr_decay_iters else 0
        self.ema_decay = float(ema_decay) if ema_decay else None
        self.world_size, self.pg = int(world_size), process_group
        self.use_graph = bool(use_graph)
        self.fuse_composite = bool(fuse_composite)
        self.global_step = 0
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)

        # ---- flat fp32 parameter / gradient / Adam state: [grid table | MLP weights] ----------------------------
        weights = model.mlp_weights()
        table = model.encoder.embeddings
        self.n_table = table.numel()
        sizes = [w.numel() for w in weights]
        self.n_weights = sum(sizes)
        # data parallel with a sharded optimiser ("zero1"): the table region is padded to world_size equal shards of a multiple of
        # 8 elements; rank r owns elements [r * shard_len, (r + 1) * shard_len)
        # data-parallel exchange (world_size > 1):
        #   "fused"     gradient buffer + fp16 table in symmetric memory; ONE kernel per step sums this rank's shard of the table gradient
        #               over the ranks (in the NVSwitch when a multicast mapping exists), runs Adam on it and writes the refreshed fp16
        #               rows into every rank's table (csrc/dp_fused.cu); two cross-rank barriers per step, no NCCL call in the step
        #   "sharded"   NCCL reduce-scatter -> Adam on the own shard -> NCCL all-gather, captured in the step graph
        #   "allreduce" NCCL all-reduce of the whole flat buffer, every rank runs the whole optimiser
        W = self.world_size
        if dp_mode is None:
            dp_mode = os.environ.get("SEALD_DP_MODE", "fused" if shard_optimizer else "allreduce")
        self.dp_mode = dp_mode if W > 1 else "single"
        if self.dp_mode not in ("single", "fused", "sharded", "allreduce"):
            raise ValueError("dp_mode must be fused | sharded | allreduce")
        self.shard_optimizer = self.dp_mode == "sharded"
        self.rank = torch.distributed.get_rank(self.pg) if W > 1 else 0
        sharded_layout = self.dp_mode in ("fused", "sharded")  # table padded to W equal shards of a multiple of 8 elements
        self.shard_len, self.n_table_pad = parallel.shard_layout(self.n_table, W if sharded_layout else 1)
        if not sharded_layout:
            self.shard_len = 0
        n = self.n_table_pad + self.n_weights
        # flat buffers: [table (padded) | MLP weights | pad | overflow flag (4 floats)].  The flag lives INSIDE the gradient buffer so the
        # exchange that sums the MLP gradients also tells every rank whether any rank overflowed (GradScaler's found_inf).
        self.n_flag = (n + 3) // 4 * 4
        n_pad = self.n_flag + 4
        self._symm = None
        if self.dp_mode == "fused":
            import torch.distributed._symmetric_memory as symm
            group = self.pg if self.pg is not None else torch.distributed.group.WORLD
            self.grads = symm.empty(n_pad, dtype=torch.float32, device=dev)
            self.grads.zero_()
            self.table16_pad = symm.empty(self.n_table_pad, dtype=torch.float16, device=dev)
            self.table16_pad.zero_()
            hg, ht = symm.rendezvous(self.grads, group), symm.rendezvous(self.table16_pad, group)
            use_mc = os.environ.get("SEALD_DP_MULTICAST", "1") != "0" and bool(hg.multicast_ptr) and bool(ht.multicast_ptr)
            self._symm = (hg, ht)
            self._peer_grads = (C.c_void_p * W)(*[int(x) for x in hg.buffer_ptrs])
            self._peer_table16 = (C.c_void_p * W)(*[int(x) for x in ht.buffer_ptrs])
            self._mc_grads = int(hg.multicast_ptr) if use_mc else None      # in-switch reduction / replication (NVLS) when mapped
            self._mc_table16 = int(ht.multicast_ptr) if use_mc else None
            # the shard sum, measured (scripts/dp_kernel_bench.py): 2 GPUs peer loads 0.051 ms vs multimem.ld_reduce 0.085 ms; 8 GPUs peer
            # loads 0.096 ms (43 MB inbound) vs in-switch reduction 0.072 ms (6 MB inbound) -> multicast from 4 ranks up.  The small MLP
            # region is always summed with peer loads (8 GPUs: 0.040 vs 0.056 ms for the stage).
            mcr = os.environ.get("SEALD_DP_MULTICAST_REDUCE", "auto")
            self._mc_reduce = use_mc and (W >= 4 if mcr == "auto" else mcr == "1")
            self.found_inf_global = torch.zeros(1, **i32)
            self.grad_shard = torch.zeros(self.shard_len, **f32)
        else:
            self.grads = torch.zeros(n_pad, **f32)
            self.table16_pad = torch.zeros(self.n_table_pad, dtype=torch.float16, device=dev)
        self.params = torch.zeros(n_pad, **f32)
        self.exp_avg = torch.zeros(n_pad, **f32)
        self.exp_avg_sq = torch.zeros(n_pad, **f32)
        self.n_params = n
        self.n_real_params = self.n_table + self.n_weights
        with torch.no_grad():
            self.params[:self.n_table].copy_(table.data.reshape(-1))
            table.data = self.params[:self.n_table].view_as(table)
            o = self.n_table_pad
            self.weight_views, self.grad_views = [], []
            for w, k in zip(weights, sizes):
                self.params[o:o + k].copy_(w.data.reshape(-1))
                w.data = self.params[o:o + k].view_as(w)
                self.weight_views.append(w.data)
                self.grad_views.append(self.grads[o:o + k].view_as(w))
                o += k
        self.grad_table = self.grads[:self.n_table].view_as(table)
        self.table16 = self.table16_pad[:self.n_table].view(table.shape)
        self.table16.copy_(table.data)
        if self.dp_mode == "sharded":
            self.grad_shard = torch.zeros(self.shard_len, **f32)               # this rank's slice of the summed table gradient
            self.shard16 = torch.zeros(self.shard_len, dtype=torch.float16, device=dev)  # ... and of the refreshed fp16 table
            self.shard16.copy_(self.table16_pad[self.rank * self.shard_len:(self.rank + 1) * self.shard_len])
        self._side = torch.cuda.Stream(device=dev)
        # the table pass of the optimiser is software-pipelined into the next step (single GPU and the fused exchange); flush() / sync_params()
        # bring the parameters up to date.  pending = {found_inf (1 = nothing pending), step counter, loss-scale bits} of the stashed update
        defer = os.environ.get("SEALD_DEFER", "auto")
        if defer == "auto":
            # measured: on one GPU the overlap buys nothing (0.403 vs 0.406 ms: the march is a chain of dependent bitfield look-ups and
            # slows down beside a pass that saturates HBM); with the fused exchange it hides the NVLink stores (0.437 -> 0.417 ms on 2 GPUs)
            defer = "start" if self.dp_mode == "fused" else "0"
        self.defer_table_update = bool(defer_table_update) and self.dp_mode in ("single", "fused") and defer != "0"
        self.defer_point = defer
        self.pending = torch.tensor([1, 0, 0, 0], dtype=torch.int32, device=dev)
        self.adam_blocks = int(os.environ.get("SEALD_ADAM_BLOCKS", "296"))
        self.one_launch_optimizer = os.environ.get("SEALD_OPT_ONE_LAUNCH", "1") != "0"  # measurement switch (seald_optimizer_step)
        # one GPU: the scatter beside the deformation backward measured SLOWER than in line (0.414 vs 0.400-0.408 ms: its 2656 short CTAs
        # delay the persistent tensor-core kernel's CTAs); the second stream is for the data-parallel exchange
        self.fork_scatter = os.environ.get("SEALD_FORK", "0") != "0"
        self.hw = F.HalfWeights(self.cfg, dev)
        self.hw.refresh(self.weight_views)
        self.lr_scale = torch.ones(1, **f32)
        self.sched_step = torch.zeros(1, **i32)
        self._tail_sync = torch.zeros(2, **i32)
        self._tail_segs = self._build_tail_segs(sizes)
        # the fused MLP tail (csrc/optim_tail.cu): single GPU; the data-parallel modes keep their exchange-fused optimiser kernels
        self.fused_tail = self.dp_mode == "single" and os.environ.get("SEALD_FUSED_TAIL", "1") != "0"
        # data parallel with the fused exchange: the same single-launch tail with the gradients summed over the peers inside it
        # (seald_mlp_tail_dp), in its flags-only form — every rank flags its own gradients (table scatter + weight-gradient flush), the
        # tail sums the flags: no check pass, no grid barrier.  Measured on 2 GPUs: 0.3145 ms (17 launches) vs 0.3172 ms (20) with the
        # four small kernels; the variant WITH the check pass + grid barrier was slower (0.3276).  It also advances the LR schedule
        self.dp_tail = (self.dp_mode == "fused" and F.WGRAD_IMPL == "umma"
                        and (os.environ.get("SEALD_DP_TAIL", "1") != "0" or bool(self.lr_decay_iters)))
        # torch_ema.ExponentialMovingAverage over every parameter (ema_decay = 0.95 in main_dnerf.py:136): a shadow of the flat buffer
        self.ema_shadow = self.params.clone() if self.ema_decay else None
        self.ema_num_updates = 0
        self._ema_backup = None

        # ---- per-step buffers ----------------------------------------------------------------------------------
        N, M = self.N, self.M
        # the step's inputs live in ONE device buffer [rays_o | rays_d | gt | time] so that host batches arrive with a single H2D copy
        self.inputs = torch.zeros(9 * N + 4, **f32)
        self.rays_o = self.inputs[0:3 * N].view(N, 3)
        self.rays_d = self.inputs[3 * N:6 * N].view(N, 3)
        self.gt = self.inputs[6 * N:9 * N].view(N, 3)
        self.bg = torch.ones(N, 3, **f32)
        self.time = self.inputs[9 * N:9 * N + 1]
        self.noises = torch.zeros(N, **f32)
        self.noise_ctr = torch.zeros(2, dtype=torch.int64, device=dev)  # seald_step_begin: {draw counter, CTA ticket}
        self.nears = torch.empty(N, **f32)
        self.fars = torch.empty(N, **f32)
        self.bitfield_frame = torch.zeros(model.density_bitfield.shape[1], dtype=torch.uint8, device=dev)
        # box of the occupied cells per time frame (+ guard band): rays that cannot meet an occupied cell skip the walk
        self.occ_all = torch.zeros(model.density_bitfield.shape[0], 6, **f32)
        self.occ_frame = torch.zeros(1, 6, **f32)
        self.refresh_occupancy()
        self.xyzs = torch.zeros(M, 3, **f32)
        self.dirs = torch.zeros(M, 3, **f32)
        self.deltas = torch.zeros(M, 2, **f32)
        self.rays = torch.zeros(N, 3, **i32)
        self.counter = torch.zeros(2, **i32)
        self.weights_sum = torch.empty(N, **f32)
        self.depth = torch.empty(N, **f32)
        self.image = torch.empty(N, 3, **f32)
        self.pred = torch.empty(N, 3, **f32)
        self.loss = torch.zeros(1, **f32)
        self.grad_image = torch.empty(N, 3, **f32)
        self.grad_ws = torch.empty(N, **f32)
        self.grad_sigma = torch.zeros(M, **f32)
        self.grad_rgb = torch.zeros(M, 3, **f32)
        self.ws = F.FieldWorkspace(self.cfg, M, dev, training=True)
        self.jobs, self.n_jobs = F.wgrad_jobs(self.cfg, self.ws, self.grad_views, deform=self.train_deform)
        self.loss_scale = torch.full((1,), float(init_loss_scale), **f32)
        self.found_inf = self.grads[self.n_flag:self.n_flag + 1].view(torch.int32)  # 0 or the bits of a positive float
        self.growth_tracker = torch.zeros(1, **i32)
        self.step_dev = torch.zeros(1, **i32)  # optimiser step counter (device side: the optimiser is graph-replayed)
        # pinned staging for the end-to-end path
        self.h_inputs = torch.zeros(9 * N + 4).pin_memory()
        self.h_rays_o = self.h_inputs[0:3 * N].view(N, 3)
        self.h_rays_d = self.h_inputs[3 * N:6 * N].view(N, 3)
        self.h_gt = self.h_inputs[6 * N:9 * N].view(N, 3)
        self.h_time = self.h_inputs[9 * N:9 * N + 1]
        self.h_loss = torch.zeros(1).pin_memory()
        self._graphs = {}
        self._occ = None
        self._trace = None
        self._ds = None
        self._ds_active = False
        self.launches_per_step = 0

    # ------------------------------------------------------------------------------------------------------------
    def _build_tail_segs(self, sizes):
        """seald_tail_seg table of the MLP weights in flat-buffer order: where the fp16 copies of every matrix live (row-padded staging
        copy; K-major tcgen05 operand tile and its transpose for the deformation layers — the layouts of csrc/field_umma.cu)."""
        cfg, hw = self.cfg, self.hw
        nd = cfg.n_deform

        def fwd_bytes(l):
            return F.DEFORM_K0 * 128 * 2 if l == 0 else (128 * 16 * 2 if l == nd - 1 else 128 * 128 * 2)

        def bwd_bytes(l):
            return 16 * 128 * 2 if l == nd - 1 else 128 * 128 * 2

        segs, first = [], 0
        for k, ((rows, cols, ld), n) in enumerate(zip(hw.shapes, sizes)):
            assert rows * cols == n
            packed = packedT = None
            n_pad = 0
            if k < nd:
                packed = hw.packed_deform.data_ptr() + sum(fwd_bytes(i) for i in range(k))
                n_pad = 16 if k == nd - 1 else 128
                if k >= 1:
                    packedT = hw.packed_deform_T.data_ptr() + sum(bwd_bytes(i) for i in range(1, k))
            segs.append(_lib.TailSeg(first, rows, cols, ld, hw.views[k].data_ptr(), packed, n_pad, packedT))
            first += n
        return (_lib.TailSeg * len(segs))(*segs)

    # ---- exponential moving average of the parameters (torch_ema semantics; the reference updates it once per epoch) -----------
    def ema_update(self):
        """ExponentialMovingAverage.update(): decay = min(ema_decay, (1 + n) / (10 + n)); shadow -= (1 - decay) * (shadow - param)."""
        if self.ema_shadow is None:
            raise RuntimeError("FusedTrainer was built without ema_decay")
        self.sync_params()
        self.ema_num_updates += 1
        decay = min(self.ema_decay, (1 + self.ema_num_updates) / (10 + self.ema_num_updates))
        _lib.call("seald_ema_update", ptr(self.ema_shadow), ptr(self.params), self.n_params, float(decay), _lib.stream())

    def _restage(self):
        self.table16.copy_(self.model.encoder.embeddings.data)
        if self.dp_mode == "sharded":
            self.shard16.copy_(self.table16_pad[self.rank * self.shard_len:(self.rank + 1) * self.shard_len])
        self.hw.refresh(self.weight_views)

    def ema_copy_to(self):
        """ema.store(); ema.copy_to(): evaluate / export with the averaged parameters (nerf/utils.py:939-941, 1071-1073)."""
        self.sync_params()
        self._ema_backup = self.params.clone()
        self.params.copy_(self.ema_shadow)
        self._restage()

    def ema_restore(self):
        """ema.restore(): back to the raw parameters."""
        if self._ema_backup is None:
            return
        self.params.copy_(self._ema_backup)
        self._ema_backup = None
        self._restage()

    def set_lr_scale(self, factor, sched_step=None):
        """Overwrite the learning-rate factor (e.g. when resuming: LambdaLR's last_epoch)."""
        self.lr_scale.fill_(float(factor))
        if sched_step is not None:
            self.sched_step.fill_(int(sched_step))

    @property
    def current_lr(self):
        f = float(self.lr_scale)
        return self.lr * f, self.lr_net * f

    # ------------------------------------------------------------------------------------------------------------
    def refresh_occupancy(self):
        """Recompute the per-frame occupied-cell boxes; call after the model's density_bitfield changed (update_extra_state)."""
        from . import raymarching
        m = self.model
        for t in range(m.density_bitfield.shape[0]):
            raymarching.occupancy_aabb(m.density_bitfield[t], m.cascade, m.grid_size, m.bound, 2, out=self.occ_all[t])

    def update_extra_state(self, decay=0.95):
        """Occupancy-grid refresh (NeRFRenderer.update_extra_state, dnerf/renderer.py:453-555) through the fused device pipeline
        (occupancy_fused.py) on the trainer's own fp16 weights / table; time frames are sharded over the ranks."""
        from .occupancy_fused import FusedOccupancy
        self.sync_params()  # the density field must see the last update on every rank (sharded modes: fp16 rows exchanged now)
        if self._occ is None:
            self._occ = FusedOccupancy(self.model, hw=self.hw, table16=self.table16, rank=self.rank, world_size=self.world_size,
                                       process_group=self.pg)
        self._occ.update(decay)
        self.refresh_occupancy()

    # ---- step inputs generated on the device from a resident dataset ----------------------------------------
    def attach_dataset(self, poses, intrinsics, H, W, images, times):
        """Keep the training set on the GPU (the reference's `preload`, dnerf/provider.py:246-252): poses [F,4,4] cam2world, intrinsics
        (fx, fy, cx, cy), images [F, H*W, 3|4] float32, times [F].  train_step_frame(i) then draws the pixels, builds the rays
        (get_rays, nerf/utils.py:54-137) and gathers the targets (collate, dnerf/provider.py:340-343) INSIDE the step graph."""
        dev = self.device
        self._ds = dict(poses=poses.to(dev, torch.float32).contiguous(), times=times.to(dev, torch.float32).reshape(-1).contiguous(),
                        images=images.to(dev, torch.float32).contiguous(), H=int(H), W=int(W), C=int(images.shape[-1]),
                        intr=[float(v) for v in intrinsics])
        if self._ds["C"] not in (3, 4) or self._ds["images"].shape[1] != H * W:
            raise ValueError("images must be [F, H*W, 3|4]")
        self.inds = torch.zeros(self.N, dtype=torch.int64, device=dev)
        self.frame_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.h_frame = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._graphs.pop(True, None)

    def _dataset_prologue(self):
        ds = self._ds
        self.inds.random_(0, ds["H"] * ds["W"])  # torch.randint(0, H*W, [N]) of get_rays (:95-97); philox state is graph-safe
        n_launch = 2
        if ds["C"] == 4:
            # RGBA frames train against a fresh random background every step (dnerf/utils.py:68-73: bg_color = rand_like(rgb)); the
            # gather kernel blends the target with bg[n] and the compositing kernel blends the prediction with the same row
            self.bg.uniform_(0, 1)
            n_launch = 3
        fx, fy, cx, cy = ds["intr"]
        _lib.call("seald_get_rays_gather", ptr(ds["poses"]), ptr(ds["times"]), ptr(ds["images"]), ptr(self.frame_dev), ptr(self.inds), self.N,
                  ds["H"], ds["W"], ds["C"], fx, fy, cx, cy, ptr(self.bg), ptr(self.rays_o), ptr(self.rays_d), ptr(self.gt), ptr(self.time),
                  _lib.stream())
        return n_launch

    def train_step_frame(self, frame, host_loss=False):
        """One step on N random pixels of training frame `frame` of the attached dataset: the host sends 4 bytes.  host_loss=True
        reads the loss back (one D2H + sync), the end-to-end variant."""
        if self._ds is None:
            raise RuntimeError("attach_dataset() first")
        self.h_frame[0] = int(frame)
        self.frame_dev.copy_(self.h_frame, non_blocking=True)
        self.step(dataset=True)
        if not host_loss:
            return self.loss
        self.h_loss.copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.h_loss[0])

    def set_inputs(self, rays_o, rays_d, time, gt_rgb, bg_color=None):
        """Device-resident inputs of the next step (copied into the static buffers the graph reads)."""
        self.rays_o.copy_(rays_o.reshape(-1, 3), non_blocking=True)
        self.rays_d.copy_(rays_d.reshape(-1, 3), non_blocking=True)
        self.gt.copy_(gt_rgb.reshape(-1, 3), non_blocking=True)
        if torch.is_tensor(time):
            self.time.copy_(time.reshape(-1)[:1], non_blocking=True)
        else:
            self.time.fill_(float(time))
        if bg_color is None:
            self.bg.fill_(1.0)
        elif torch.is_tensor(bg_color):
            self.bg.copy_(bg_color.reshape(-1, 3).expand(self.N, 3), non_blocking=True)
        else:
            self.bg.fill_(float(bg_color))

    def set_inputs_host(self, rays_o, rays_d, time, gt_rgb):
        """Host inputs: staged through pinned memory, H2D on the current stream (counted in the e2e timing)."""
        self.h_rays_o.copy_(rays_o.reshape(-1, 3))
        self.h_rays_d.copy_(rays_d.reshape(-1, 3))
        self.h_gt.copy_(gt_rgb.reshape(-1, 3))
        self.h_time[0] = float(time)
        self.inputs.copy_(self.h_inputs, non_blocking=True)

    @property
    def h2d_bytes_per_step(self):
        return self.inputs.numel() * 4

    # ------------------------------------------------------------------------------------------------------------
    def _stages(self):
        """The step as an ordered list of (name, callable, kernel launches): shared by the real step and by the per-stage timer."""
        m, cfg = self.model, self.cfg
        N, M = self.N, self.M
        ws, hw = self.ws, self.hw
        m_dev = self.counter[0:1]
        offsets = m.encoder.offsets
        inv_count = parallel.loss_inv_count(N, self.world_size)
        F16, F32 = _lib.F16, _lib.F32

        def select_frame():
            # occupancy frame of this time stamp (dnerf/renderer.py:285) + its occupied-cell box, selected on the device; sample counter
            # and loss accumulator reset; this step's perturbation noise (a counter-based hash: graph replays draw fresh numbers)
            _lib.call("seald_step_begin", ptr(self.time), int(m.time_size), ptr(m.density_bitfield), int(m.density_bitfield.shape[1]),
                      ptr(self.bitfield_frame), ptr(self.occ_all), ptr(self.occ_frame), ptr(self.counter), ptr(self.loss),
                      ptr(self.noises) if self.perturb else None, N, ptr(self.noise_ctr), _lib.stream())

        def march():
            _lib.call("seald_march_rays_train", ptr(self.rays_o), ptr(self.rays_d), ptr(self.bitfield_frame), float(m.bound), self.dt_gamma,
                      self.max_steps, N, int(m.cascade), int(m.grid_size), M, None, None, ptr(m.aabb_train), float(m.min_near),
                      ptr(self.nears), ptr(self.fars), ptr(self.xyzs), ptr(self.dirs), ptr(self.deltas), ptr(self.rays), ptr(self.counter),
                      ptr(self.noises), ptr(self.occ_frame), _lib.stream())

        def deform_fwd():
            F.deform_forward(cfg, hw, self.xyzs, self.time, M, m_dev, 1, ws.deform, ws.x01, ws.in_buf, ws.fwd_d)

        def grid_fwd():
            _lib.call("seald_grid_encode_forward", ptr(ws.x01), ptr(self.table16), ptr(offsets), ptr(ws.feat), None, M, 3, cfg.grid_dim,
                      cfg.grid_levels, cfg.grid_S, cfg.grid_base, cfg.gridtype, int(cfg.align_corners), cfg.interp, F16, ptr(m_dev),
                      _lib.stream())

        def heads_fwd():
            F.heads_forward(cfg, hw, ws, self.dirs, M, m_dev, True)

        def grid_heads_fwd():
            # hash-grid encoder inside the heads kernel: one launch, the features never reach HBM
            F.grid_heads_forward(cfg, hw, ws, self.dirs, self.table16, offsets, M, m_dev, True)

        def composite_fwd():
            _lib.call("seald_composite_rays_train_forward", ptr(ws.sigma), ptr(ws.rgb), ptr(self.deltas), ptr(self.rays), M, N, self.T_thresh,
                      ptr(self.weights_sum), ptr(self.depth), ptr(self.image), _lib.stream())

        def loss():
            self.loss.zero_()
            _lib.call("seald_mse_loss_bg", ptr(self.image), ptr(self.weights_sum), ptr(self.bg), ptr(self.gt), N, inv_count,
                      ptr(self.loss_scale), ptr(self.pred), ptr(self.loss), ptr(self.grad_image), ptr(self.grad_ws), _lib.stream())

        def composite_bwd():
            self.grad_sigma.zero_()
            self.grad_rgb.zero_()
            _lib.call("seald_composite_rays_train_backward", ptr(self.grad_ws), ptr(self.grad_image), ptr(ws.sigma), ptr(ws.rgb),
                      ptr(self.deltas), ptr(self.rays), ptr(self.weights_sum), ptr(self.image), M, N, self.T_thresh, ptr(self.grad_sigma),
                      ptr(self.grad_rgb), _lib.stream())

        def heads_bwd():
            F.heads_backward(cfg, hw, ws, self.grad_sigma, self.grad_rgb, M, m_dev)

        def grid_scatter():
            # table gradient; also raises the overflow flag when dfeat holds an inf/nan (== the table gradient would)
            _lib.call("seald_grid_encode_backward_table", ptr(ws.dfeat), ptr(ws.x01), ptr(offsets), ptr(self.grad_table), M, 3, cfg.grid_dim,
                      cfg.grid_levels, cfg.grid_S, cfg.grid_base, cfg.gridtype, int(cfg.align_corners), cfg.interp, F16, F32, ptr(m_dev),
                      ptr(self.found_inf), _lib.stream())

        def grid_bwd_both():
            # table scatter + input gradient in one launch (CTA roles): both only need dfeat and x01 and are latency bound at this size
            _lib.call("seald_grid_encode_backward_both", ptr(ws.dfeat), ptr(ws.x01), ptr(self.table16), ptr(offsets), ptr(self.grad_table),
                      ptr(ws.grad_x01), M, 3, cfg.grid_dim, cfg.grid_levels, cfg.grid_S, cfg.grid_base, cfg.gridtype, int(cfg.align_corners),
                      cfg.interp, F16, F32, ptr(m_dev), ptr(self.found_inf), _lib.stream())

        def grid_input_bwd():
            _lib.call("seald_grid_encode_backward_input", ptr(ws.dfeat), ptr(ws.x01), ptr(self.table16), ptr(offsets), None, ptr(ws.grad_x01),
                      M, 3, cfg.grid_dim, cfg.grid_levels, cfg.grid_S, cfg.grid_base, cfg.gridtype, int(cfg.align_corners), cfg.interp, F16,
                      ptr(m_dev), _lib.stream())

        def deform_bwd():
            F.deform_backward(cfg, hw, ws.grad_x01, self.time, M, m_dev, ws.fwd_d, ws.bwd_d, ws.gout_d)

        def wgrad():
            if self._wgrad_flags():
                # + this rank's overflow flag for its weight gradients (no separate finite-check pass; the tail sums the ranks' flags)
                _lib.call("seald_mlp_wgrad_umma_flag", C.cast(self.jobs, C.c_void_p), self.n_jobs, M, ptr(m_dev), ptr(self.found_inf), _lib.stream())
            else:
                F.mlp_wgrad(self.jobs, self.n_jobs, M, m_dev)

        def composite_loss_fused():
            # composite forward + loss + composite backward in one kernel (bg/gt are indexed by ray); the loss accumulator was zeroed by
            # the step's first kernel
            _lib.call("seald_composite_train_loss_fused", ptr(ws.sigma), ptr(ws.rgb), ptr(self.deltas), ptr(self.rays), M, N, self.T_thresh,
                      ptr(self.bg), ptr(self.gt), inv_count, ptr(self.loss_scale), ptr(self.weights_sum), ptr(self.depth), ptr(self.image),
                      ptr(self.pred), ptr(self.loss), ptr(self.grad_sigma), ptr(self.grad_rgb), _lib.stream())

        if self.fuse_composite:
            tail = [("composite_loss_fused", composite_loss_fused, 1)]
        else:
            tail = [("composite_fwd", composite_fwd, 1), ("loss", loss, 2), ("composite_bwd", composite_bwd, 3)]
        field = [("grid_heads_fwd", grid_heads_fwd, 1)] if F.grid_heads_fusable(cfg, ws, True) else [("grid_fwd", grid_fwd, 1), ("heads_fwd", heads_fwd, 1)]
        stages = [("select_frame", select_frame, 1), ("march", march, 1), ("deform_fwd", deform_fwd, 1)] + field + tail + [("heads_bwd", heads_bwd, 1)]
        # one GPU: scatter and input gradient share a launch.  Data parallel: the scatter runs on the side stream ahead of the exchange
        # beside the input gradient (measured on 2 GPUs: 0.3088 ms split vs 0.3135 ms merged)
        both = self.train_deform and self.dp_mode == "single" and not self.fork_scatter and os.environ.get("SEALD_GRID_BWD_SPLIT", "0") == "0"
        if both:
            stages += [("grid_bwd_both", grid_bwd_both, 1), ("deform_bwd", deform_bwd, 1)]
        else:
            stages.append(("grid_scatter", grid_scatter, 1))
            if self.train_deform:
                stages += [("grid_input_bwd", grid_input_bwd, 1), ("deform_bwd", deform_bwd, 1)]
        stages.append(("wgrad", wgrad, 1))
        return stages

    # ------------------------------------------------------------------------------------------------------------
    # The step body.  ONE function describes the step for 1 GPU, data parallel with a flat all-reduce, and data parallel with a
    # sharded optimiser; it is either run eagerly or captured — collectives included — into ONE CUDA graph.  Two streams:
    #
    #   main : [select frame, march, deformation forward] . grid forward .. heads backward | input grad, deform backward, wgrad, check | Adam
    #   side : [all-gather fp16 table (sharded)]                                           | table scatter, reduce-scatter / all-reduce |
    #
    # The table scatter (atomic/latency bound) runs beside the tensor-core backward of the deformation net; with a sharded
    # optimiser the reduce-scatter of the table gradient hides behind the same kernels and the all-gather of the refreshed
    # fp16 table behind the NEXT step's march + deformation forward, which do not read the table.
    def _step_body(self):
        S = {name: (fn, k) for name, fn, k in self._stages()}
        n = [0]

        def run(*names):
            for nm in names:
                if nm in S:
                    S[nm][0]()
                    n[0] += S[nm][1]
                    mark(nm)

        def mark(label):
            # optional in-graph timeline: external event-record nodes after every stage of the main stream (step_timeline())
            if self._trace is not None:
                ev = torch.cuda.Event(enable_timing=True, external=True)
                ev.record(torch.cuda.current_stream())
                self._trace.append((label, ev))

        dist = torch.distributed
        W, mode = self.world_size, self.dp_mode
        mark("begin")
        if self._ds is not None and self._use_dataset:
            n[0] += self._dataset_prologue()
            mark("get_rays")
        main, side = torch.cuda.current_stream(), self._side
        ntp = self.n_table_pad
        # ---- beginning of the step: the march and the deformation forward do not read the hash table, so the table part of the
        # PREVIOUS step's optimiser (the largest memory pass of a step) runs beside them on the second stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if self.defer_table_update and self.defer_point == "start":
                n[0] += self._optimizer_table_deferred()
            if mode == "sharded":  # last step's refreshed table shards (a no-op exchange before the first step)
                dist.all_gather_into_tensor(self.table16_pad, self.shard16, group=self.pg)
            elif mode == "fused":
                # every rank finished writing its fp16 rows into our table, and nobody reads our gradient buffer any more: clear it
                self._symm[0].barrier(1)
                self.grads.zero_()
                n[0] += 2
        run("select_frame", "march")
        if self.defer_table_update and self.defer_point == "after_march":
            # (the march is a chain of dependent bitfield look-ups: its latency suffers beside a pass that saturates HBM, so on one GPU the
            # table pass only shares the GPU with the tensor-core deformation forward)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                n[0] += self._optimizer_table_deferred()
        run("deform_fwd")
        main.wait_stream(side)
        run("grid_fwd", "heads_fwd", "grid_heads_fwd", "composite_fwd", "loss", "composite_bwd", "composite_loss_fused", "heads_bwd")
        run("grid_bwd_both")  # (one GPU: scatter + input gradient in one launch; then "grid_scatter" / "grid_input_bwd" below are absent)
        # ---- the table scatter (atomics) beside the tensor-core backward of the deformation net
        fork = self.fork_scatter or mode in ("fused", "sharded", "allreduce")
        if fork:
            side.wait_stream(main)
        with torch.cuda.stream(side if fork else main):
            run("grid_scatter")
            if mode == "sharded":
                dist.reduce_scatter_tensor(self.grad_shard, self.grads[:ntp], op=dist.ReduceOp.SUM, group=self.pg)
                self.grads[:ntp].zero_()  # consumed; the Adam kernel only sees the shard
            elif mode == "allreduce":
                dist.all_reduce(self.grads[:ntp], op=dist.ReduceOp.SUM, group=self.pg)
            elif mode == "fused":
                # every rank's table gradient is complete -> pull this rank's shard of the sum over NVLink, underneath the backward
                # of the deformation net
                self._symm[0].barrier(0)
                _lib.call("seald_dp_reduce_shard", C.cast(self._peer_grads, C.c_void_p), self._mc_grads if self._mc_reduce else None, W,
                          self.rank * self.shard_len, self.shard_len, ptr(self.grad_shard), _lib.stream())
                n[0] += 2
        run("grid_input_bwd", "deform_bwd", "wgrad")
        if not self.fused_tail and not self.dp_tail:
            _lib.call("seald_grad_finite_check", self.grads.data_ptr() + 4 * ntp, self.n_weights, ptr(self.found_inf), _lib.stream())
            n[0] += 1
        main.wait_stream(side)  # (also orders the overflow flag written by the scatter before it is exchanged)
        if mode in ("sharded", "allreduce"):  # MLP gradients + the overflow flag (a float: > 0 on every rank if any rank overflowed)
            dist.all_reduce(self.grads[ntp:], op=dist.ReduceOp.SUM, group=self.pg)
        elif mode == "fused":
            self._symm[0].barrier(2)  # every rank's MLP gradients and overflow flag are complete
            n[0] += 1
        mark("check+exchange")
        n[0] += self._optimizer()
        mark("optimizer")
        return n[0]

    def step_timeline(self, reps=50):
        """Time between consecutive stage boundaries INSIDE the replayed step graph (external event-record nodes), averaged over
        `reps` replays: what each stage costs in its real position, caches as the previous stage left them."""
        if not self.use_graph:
            raise RuntimeError("step_timeline() needs use_graph=True")
        snapshot = [t.clone() for t in self._state()]
        self._warmup()
        self._trace = []
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step_body()
        trace, self._trace = self._trace, None
        acc = [0.0] * (len(trace) - 1)
        for _ in range(3):
            g.replay()
        for _ in range(reps):
            g.replay()
            torch.cuda.synchronize()
            for i in range(len(trace) - 1):
                acc[i] += trace[i][1].elapsed_time(trace[i + 1][1])
        out = {}
        for i in range(len(trace) - 1):
            out[trace[i + 1][0]] = out.get(trace[i + 1][0], 0.0) + acc[i] / reps
        out["total"] = sum(acc) / reps
        del g
        torch.cuda.synchronize()
        self._restore(snapshot)
        return out

    def _forward_backward(self):
        """Every stage of the forward + backward pass in order on the current stream (no optimiser, no exchange): tests / debugging."""
        n = 0
        for _, fn, k in self._stages():
            fn()
            n += k
        return n

    def _allreduce(self):
        """Flat all-reduce of the whole gradient buffer, overflow flag included (the unsharded data-parallel exchange, in one piece)."""
        if self.world_size > 1:
            parallel.allreduce_flat_grads(self.grads, self.pg)

    def _adam_table(self, step_dev, loss_scale, found_inf, st, lr_scale=None):
        b1, b2 = self.betas
        if lr_scale is None:
            lr_scale = ptr(self.lr_scale)
        if self.dp_mode == "fused":  # own shard (gradient already summed over the ranks) + fp16 rows to every rank's table
            _lib.call("seald_dp_adam_shard_broadcast", C.cast(self._peer_table16, C.c_void_p), self._mc_table16, self.world_size,
                      ptr(self.params), ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self.grad_shard), self.rank * self.shard_len,
                      self.shard_len, self.lr, b1, b2, self.eps, step_dev, loss_scale, found_inf,
                      lr_scale if self.dp_tail else None, st)
        elif self.dp_mode == "sharded":
            off = self.rank * self.shard_len
            _lib.call("seald_adam_step", self.params.data_ptr() + 4 * off, ptr(self.grad_shard), self.exp_avg.data_ptr() + 4 * off,
                      self.exp_avg_sq.data_ptr() + 4 * off, self.shard_len, self.lr, b1, b2, self.eps, 1, step_dev, loss_scale, found_inf,
                      ptr(self.shard16), 0, st)
        else:
            # beside the march the pass must leave room on every SM (a full-occupancy grid would push the latency-bound march behind it)
            cap = self.adam_blocks if self.defer_table_update else 0
            _lib.call("seald_adam_step_lr", ptr(self.params), ptr(self.grads), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.n_table_pad,
                      self.lr, lr_scale, b1, b2, self.eps, 1, step_dev, loss_scale, found_inf, ptr(self.table16_pad), 1, cap, st)
        return 1

    def _wgrad_flags(self):
        """The weight-gradient kernel raises GradScaler's overflow flag itself (seald_mlp_wgrad_umma_flag): the single-launch optimiser
        (one GPU, table pass in place) and the data-parallel tail then need no check pass and no grid barrier."""
        if F.WGRAD_IMPL != "umma":
            return False
        return self.dp_tail or (self.fused_tail and not self.defer_table_update and self.dp_mode == "single" and self.one_launch_optimizer
                                and os.environ.get("SEALD_FLAGS_FINAL", "1") != "0")

    def _optimizer_table_deferred(self):
        """The hash-table pass of the previous step's optimiser, with the overflow decision / step number / loss scale that step
        stashed (the first call finds found_inf = 1: nothing pending)."""
        p = self.pending.data_ptr()
        return self._adam_table(p + 4, p + 8, p, _lib.stream(), p + 12 if (self.fused_tail or self.dp_tail) else None)

    def _optimizer(self):
        """GradScaler.step + optimizer.step + GradScaler.update as device kernels (nerf/utils.py:884-886): the whole update is
        skipped when the overflow flag is set, the loss scale backs off / grows, the fp16 copies are refreshed.  With
        defer_table_update the hash-table pass is not launched here but at the beginning of the next step (or by flush())."""
        st = _lib.stream()
        b1, b2 = self.betas
        ntp = self.n_table_pad
        found = self.found_inf
        if self.fused_tail:
            # ONE launch: overflow check of the MLP gradients, Adam on the weights, fp16 copies + tcgen05 tiles, GradScaler.update,
            # lr_scheduler.step; it stashes {found_inf, step, loss scale, lr factor} of this step for the table pass
            o = 4 * ntp
            tail_args = (self.params.data_ptr() + o, self.grads.data_ptr() + o, self.exp_avg.data_ptr() + o, self.exp_avg_sq.data_ptr() + o,
                         C.cast(self._tail_segs, C.c_void_p), len(self._tail_segs), self.lr_net, b1, b2, self.eps, ptr(self.step_dev),
                         ptr(self.loss_scale), ptr(found), ptr(self.growth_tracker), 2.0, 0.5, self.growth_interval, ptr(self.pending),
                         ptr(self.lr_scale), ptr(self.sched_step), self.lr_decay_iters, ptr(self._tail_sync))
            if not self.defer_table_update and self.dp_mode == "single" and self.one_launch_optimizer:
                # one GPU, table pass in place: the whole scaler.step / scaler.update / lr_scheduler.step in ONE launch
                _lib.call("seald_optimizer_step", *tail_args, ptr(self.params), ptr(self.grads), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                          self.n_table_pad, self.lr, ptr(self.table16_pad), 1 if self._wgrad_flags() else 0, st)
                return 1
            _lib.call("seald_mlp_tail", *tail_args, st)
            n = 1
            if not self.defer_table_update:
                n += self._optimizer_table_deferred()
            return n
        if self.dp_tail:
            # ONE launch: flags + MLP gradients summed over the peers, Adam on the replicated weights, fp16 copies / tcgen05 tiles,
            # GradScaler.update, LambdaLR; the stash carries {found_inf, step, loss scale, lr factor} to the deferred shard pass
            o = 4 * ntp
            _lib.call("seald_mlp_tail_dp", C.cast(self._peer_grads, C.c_void_p), self.world_size, ntp, self.n_flag, 1, ptr(self.found_inf_global),
                      self.params.data_ptr() + o, self.exp_avg.data_ptr() + o, self.exp_avg_sq.data_ptr() + o,
                      C.cast(self._tail_segs, C.c_void_p), len(self._tail_segs), self.lr_net, b1, b2, self.eps, ptr(self.step_dev),
                      ptr(self.loss_scale), ptr(self.growth_tracker), 2.0, 0.5, self.growth_interval, ptr(self.pending), ptr(self.lr_scale),
                      ptr(self.sched_step), self.lr_decay_iters, ptr(self._tail_sync), st)
            n = 1
            if not self.defer_table_update:
                n += self._optimizer_table_deferred()
            return n
        n = 3
        if self.dp_mode == "fused":  # overflow decision over the ranks + Adam on the replicated MLP weights (gradient summed over the peers)
            found = self.found_inf_global
            _lib.call("seald_dp_adam_weights", C.cast(self._peer_grads, C.c_void_p), None, self.world_size, ptr(self.params),
                      ptr(self.exp_avg), ptr(self.exp_avg_sq), ntp, self.n_weights, self.n_flag, self.lr_net, b1, b2, self.eps,
                      ptr(self.step_dev), ptr(self.loss_scale), ptr(found), st)
        else:
            _lib.call("seald_adam_step", self.params.data_ptr() + 4 * ntp, self.grads.data_ptr() + 4 * ntp, self.exp_avg.data_ptr() + 4 * ntp,
                      self.exp_avg_sq.data_ptr() + 4 * ntp, self.n_weights, self.lr_net, b1, b2, self.eps, 1, ptr(self.step_dev),
                      ptr(self.loss_scale), ptr(found), None, 1, st)
        if not self.defer_table_update:
            n += self._adam_table(ptr(self.step_dev), ptr(self.loss_scale), ptr(found), st)
        self.hw.refresh(self.weight_views)
        _lib.call("seald_loss_scale_update_stash", ptr(self.loss_scale), ptr(found), ptr(self.growth_tracker), 2.0, 0.5,
                  self.growth_interval, ptr(self.step_dev), ptr(self.pending), st)
        return n

    def pretrain_step(self, points, dirs, gt_sigma, gt_color, time, lr=0.07, optimize=True):
        """One step of Seal's local pre-training (SealNeRF/trainer.py:396-462 pretrain_part / pretrain_step; data from
        SealDNeRF/utils.py:386-562): the student field is evaluated at `points` [B,3] / `dirs` [B,3] of time stamp `time` and pulled
        towards the teacher's (gt_sigma [B], gt_color [B,3]) with L1Loss(sigma) + L1Loss(colour); every MLP is frozen (freeze_mlp,
        SealDNeRF/utils.py:365-376; the student's deformation net is frozen for the whole SealD run), so the only gradient is the hash
        table's.  GradScaler.scale(loss).backward() -> scaler.step(optimizer) with every group's lr = `lr` (set_lr, :564-577) ->
        scaler.update(), as device kernels on the trainer's own Adam state: deformation forward (tcgen05, nothing saved) -> grid
        forward -> heads forward -> L1 loss + seeds -> heads backward (input gradient only) -> table scatter -> Adam on the table.
        Returns nothing; `self.loss` holds the loss.  optimize=False stops after the backward pass (gradients stay in self.grad_table,
        tests).  (One step counter serves all groups here: frozen weights keep torch's per-parameter `step`, which only matters
        for Adam's bias correction during the first few hundred steps of a run.)"""
        if self.world_size > 1:
            raise NotImplementedError("Seal pre-training runs the same batch on every replica in the reference; use one GPU")
        m, cfg, ws, hw = self.model, self.cfg, self.ws, self.hw
        B = int(points.shape[0])
        if B == 0:
            return
        if B > self.M:
            raise ValueError("pre-training batch of %d points exceeds the trainer's sample capacity %d" % (B, self.M))
        _lib.require_cuda(points, dirs, gt_sigma, gt_color)
        self.flush()  # a pending table pass of the last training step comes first
        pts, drs = points.detach().float().contiguous(), dirs.detach().float().contiguous()
        gs, gc = gt_sigma.detach().float().contiguous().view(-1), gt_color.detach().float().contiguous().view(-1, 3)
        if torch.is_tensor(time):
            self.time.copy_(time.reshape(-1)[:1])
        else:
            self.time.fill_(float(time))
        st = _lib.stream()
        offsets = m.encoder.offsets
        F16, F32 = _lib.F16, _lib.F32
        F.deform_forward(cfg, hw, pts, self.time, B, None, 1, ws.deform, ws.x01, None, None)
        F.grid_heads_forward(cfg, hw, ws, drs, self.table16, offsets, B, None, True)
        self.loss.zero_()
        _lib.call("seald_l1_pretrain_loss", ptr(ws.sigma), ptr(ws.rgb), ptr(gs), ptr(gc), B, ptr(self.loss_scale), ptr(self.loss),
                  ptr(self.grad_sigma), ptr(self.grad_rgb), st)
        F.heads_backward(cfg, hw, ws, self.grad_sigma, self.grad_rgb, B, None)
        _lib.call("seald_grid_encode_backward_table", ptr(ws.dfeat), ptr(ws.x01), ptr(offsets), ptr(self.grad_table), B, 3, cfg.grid_dim,
                  cfg.grid_levels, cfg.grid_S, cfg.grid_base, cfg.gridtype, int(cfg.align_corners), cfg.interp, F16, F32, None,
                  ptr(self.found_inf), st)
        if optimize:
            self._pretrain_optimize(lr)

    def _pretrain_optimize(self, lr):
        """scaler.step(optimizer) + scaler.update() of a pre-training step: Adam on the hash table only (the MLP groups are frozen)."""
        st = _lib.stream()
        b1, b2 = self.betas
        _lib.call("seald_adam_step_lr", ptr(self.params), ptr(self.grads), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.n_table_pad,
                  float(lr), None, b1, b2, self.eps, 1, ptr(self.step_dev), ptr(self.loss_scale), ptr(self.found_inf), ptr(self.table16_pad),
                  1, 0, st)
        _lib.call("seald_loss_scale_update", ptr(self.loss_scale), ptr(self.found_inf), ptr(self.growth_tracker), 2.0, 0.5,
                  self.growth_interval, ptr(self.step_dev), st)
        self.global_step += 1

    def flush(self):
        """Apply the deferred hash-table update of the last step now (parameters are about to be read: evaluation, checkpoint)."""
        if self.defer_table_update:
            self._optimizer_table_deferred()
            self.pending[0] = 1  # nothing pending any more
        if self.dp_mode == "fused":
            # every rank's rows have arrived in our fp16 table; also the point after which a rank may free its symmetric buffers
            torch.cuda.current_stream().synchronize()
            torch.distributed.barrier(group=self.pg)

    def _state(self):
        return (self.params, self.exp_avg, self.exp_avg_sq, self.loss_scale, self.growth_tracker, self.table16_pad, self.hw.flat, self.step_dev,
                self.pending, self.lr_scale, self.sched_step)

    def stage_timings(self, reps=20):
        """Average device time (ms) of every stage, each timed alone: `reps` back-to-back launches captured in a CUDA graph and
        replayed, so the host's launch cost is not in the number.  Uses the inputs currently staged; the optimiser state is
        restored afterwards."""
        out = {}
        stages = [(nm, fn) for nm, fn, _k in self._stages()] + [("optimizer", self._optimizer)]
        if self.defer_table_update:
            stages.append(("optimizer_table", self._optimizer_table_deferred))
        snapshot = [t.clone() for t in self._state()]
        for _, fn in stages[:len(self._stages())]:  # one full pass so every stage sees valid inputs
            fn()
        torch.cuda.synchronize()
        out["live_samples"] = int(self.counter[0].item())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        S = dict(stages)

        def march_with_reset():  # the march appends to the sample counter: reset it per launch, like the step does
            S["select_frame"]()
            S["march"]()

        for name, fn in stages:
            if name == "march":
                fn = march_with_reset
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                fn()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(reps):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            out[name] = e0.elapsed_time(e1) / reps
            if name == "march":
                out[name] = max(out[name] - out["select_frame"], 0.0)
            del g
        torch.cuda.synchronize()
        self._restore(snapshot)
        return out

    def sync_params(self, moments=False):
        """Data parallel with a sharded table: bring every rank's fp16 table and fp32 master copy up to date (before evaluation /
        checkpoints); the per-step exchange of the fp16 table completes at the beginning of the NEXT step.  moments=True also gathers
        the Adam moments of the shards the other ranks own (a checkpoint must hold all of them)."""
        self.flush()
        if self.dp_mode not in ("fused", "sharded"):
            return
        dist = torch.distributed
        off = self.rank * self.shard_len
        if self.dp_mode == "sharded":
            dist.all_gather_into_tensor(self.table16_pad, self.shard16, group=self.pg)
        bufs = [self.params] + ([self.exp_avg, self.exp_avg_sq] if moments else [])
        for b in bufs:
            dist.all_gather_into_tensor(b[:self.n_table_pad], b[off:off + self.shard_len].clone(), group=self.pg)

    def _restore(self, snap):
        for dst, src in zip(self._state(), snap):
            dst.copy_(src)
        self.hw.refresh(self.weight_views)
        self.grads.zero_()
        if self.dp_mode == "fused":
            self.grad_shard.zero_()
        if self.dp_mode == "sharded":
            self.grad_shard.zero_()
            self.shard16.copy_(self.table16_pad[self.rank * self.shard_len:(self.rank + 1) * self.shard_len])

    def _rank_sync(self):
        torch.cuda.synchronize()
        if self.world_size > 1:
            torch.distributed.barrier(group=self.pg)

    def _warmup(self):
        """One eager step on a side stream (module loading, communicator set-up), then the state is put back.  The ranks are
        synchronised around the restore: in the fused mode the peers write their shard of the warm-up update into OUR table."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            snap = [t.clone() for t in self._state()]
            self._step_body()
        self._rank_sync()
        self._restore(snap)
        self._rank_sync()

    # ------------------------------------------------------------------------------------------------------------
    def step(self, dataset=False):
        """One optimisation step on the inputs staged by set_inputs*/() — or, with dataset=True, on pixels drawn inside the step from
        the attached dataset.  No host synchronisation.  One CUDA graph per input mode."""
        self.global_step += 1
        self._ds_active = bool(dataset)
        if not self.use_graph:
            self.launches_per_step = self._step_body()
            return
        g = self._graphs.get(self._ds_active)
        if g is None:
            self._warmup()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.launches_per_step = self._step_body()
            self._graphs[self._ds_active] = g
        g.replay()

    @property
    def _use_dataset(self):
        return self._ds is not None and self._ds_active

    def train_step(self, rays_o, rays_d, time, gt_rgb, bg_color=None):
        """Public API: one training step on device tensors; returns the (device) loss of this step."""
        self.set_inputs(rays_o, rays_d, time, gt_rgb, bg_color)
        self.step()
        return self.loss

    def train_step_host(self, rays_o, rays_d, time, gt_rgb):
        """End-to-end variant: host inputs in, host loss out (one D2H read + sync per step)."""
        self.set_inputs_host(rays_o, rays_d, time, gt_rgb)
        self.step()
        self.h_loss.copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.h_loss[0])

    # ---- software-pipelined end-to-end loop ---------------------------------------------------------------------------------------
    # train_step_host() serialises host and device: the host waits for the loss before it stages the next batch (~35 us per 0.33 ms
    # step).  Pipelined: the batch of step k is copied host -> device on a copy stream WHILE step k-1 runs (two pinned + two device
    # staging slots), step k starts with a device-to-device copy into the graph's static input buffer, and the host reads the loss
    # of step k-1 — every step's inputs still cross PCIe and every step's loss still reaches the host, one step later.
    def _pipe_init(self):
        dev = self.params.device
        n = self.inputs.numel()
        self._pipe = dict(k=0, copy_stream=torch.cuda.Stream(device=dev),
                          pinned=[torch.zeros(n).pin_memory() for _ in range(2)], staging=[torch.zeros(n, device=dev) for _ in range(2)],
                          hloss=[torch.zeros(1).pin_memory() for _ in range(2)],
                          h2d_done=[torch.cuda.Event() for _ in range(2)], d2d_done=[torch.cuda.Event() for _ in range(2)],
                          loss_done=[torch.cuda.Event() for _ in range(2)])

    def train_step_host_pipelined(self, rays_o, rays_d, time, gt_rgb):
        """End-to-end step with host inputs; returns the loss of the PREVIOUS call (None on the first one) — see drain_host_pipeline()."""
        if getattr(self, "_pipe", None) is None:
            self._pipe_init()
        P, N = self._pipe, self.N
        k = P["k"]
        s = k & 1
        if k >= 2:
            P["h2d_done"][s].synchronize()  # the pinned slot is free again (its copy was issued two calls ago: long done)
        h = P["pinned"][s]
        h[0:3 * N].view(N, 3).copy_(rays_o.reshape(-1, 3))
        h[3 * N:6 * N].view(N, 3).copy_(rays_d.reshape(-1, 3))
        h[6 * N:9 * N].view(N, 3).copy_(gt_rgb.reshape(-1, 3))
        h[9 * N] = float(time)
        main = torch.cuda.current_stream()
        cs = P["copy_stream"]
        if k >= 2:
            cs.wait_event(P["d2d_done"][s])  # the device slot was consumed by step k-2's copy into the static inputs
        with torch.cuda.stream(cs):
            P["staging"][s].copy_(h, non_blocking=True)
            P["h2d_done"][s].record(cs)
        main.wait_event(P["h2d_done"][s])
        self.inputs.copy_(P["staging"][s], non_blocking=True)
        P["d2d_done"][s].record(main)
        self.step()
        P["hloss"][s].copy_(self.loss, non_blocking=True)
        P["loss_done"][s].record(main)
        prev = None
        if k >= 1:
            P["loss_done"][1 - s].synchronize()
            prev = float(P["hloss"][1 - s][0])
        P["k"] = k + 1
        return prev

    def drain_host_pipeline(self):
        """Loss of the last pipelined step (waits for it)."""
        P = getattr(self, "_pipe", None)
        if P is None or P["k"] == 0:
            return None
        s = (P["k"] - 1) & 1
        P["loss_done"][s].synchronize()
        return float(P["hloss"][s][0])

    def calibrate_max_samples(self, rays_o, rays_d, time, margin=1.25):
        """Run the march once with a generous bound to size M (the reference's `mean_count` estimate, raymarching.py:200-203)."""
        m = self.model
        nears, fars = None, None
        from . import raymarching
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, m.aabb_train, m.min_near)
        t = m._frame_index(torch.as_tensor([[float(time)]], device=self.device))
        counter = torch.zeros(2, dtype=torch.int32, device=self.device)
        raymarching.march_rays_train(rays_o, rays_d, m.bound, m.density_bitfield[t], m.cascade, m.grid_size, nears, fars, counter, -1, False, 128,
                                     True, self.dt_gamma, self.max_steps)
        return int(counter[0].item() * margin)
