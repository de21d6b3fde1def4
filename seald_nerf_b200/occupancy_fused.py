"""FusedOccupancy — NeRFRenderer.update_extra_state (dnerf/renderer.py:453-555) as a preallocated, sync-free device pipeline.

Same schedule as the reference (full sweep of every cell of every time frame for the first 16 calls, then H^3/4 random cells +
H^3/4 re-sampled occupied cells per frame), same random-number consumption for the full sweep (so it is bit-identical to the
torch-composed `NeRFRenderer.update_extra_state` of this package under the same seed), but per time frame it is
    1 kernel   jittered sample points + Morton indices          (csrc/occupancy.cu; the reference: ~10 torch kernels)
    2 kernels  density query: tcgen05 deformation net -> hash grid + sigma head in one launch, on buffers allocated ONCE for H^3
               points (the all-in-one tcgen05 launch seald_field_density_umma is slower at this size: field.DENSITY_IMPL)
    2 kernels  store into a per-frame temporary, decayed maximum (no 512 MiB temporaries, no boolean-mask passes)
followed by ONE packbits launch over all time frames.  The partial pass draws the reference's random numbers in the reference's
order (cells, ranks among the occupied cells, jitter), so it consumes the same generator stream and samples the SAME points; the
rank-th occupied cell is found by a prefix sum + in-kernel binary search instead of `nonzero` (one host read of the 64 occupied-cell
counts per refresh instead of one synchronisation per frame).
"""
import torch

from . import _lib
from . import field as F
from . import raymarching
from ._lib import ptr


class FusedOccupancy:
    def __init__(self, model, hw=None, table16=None, rank=0, world_size=1, process_group=None):
        self.model = model
        # data parallel: the time frames are dealt to the ranks in contiguous blocks and the refreshed grids all-gathered
        self.rank, self.world_size, self.pg = int(rank), int(world_size), process_group
        if model.time_size % self.world_size:
            raise ValueError("time_size must be divisible by the number of ranks")
        m = model
        dev = m.density_grid.device
        if dev.type != "cuda":
            raise RuntimeError("FusedOccupancy needs the model on a CUDA device (no CPU fallback)")
        self.H = int(m.grid_size)
        self.n_full = self.H ** 3
        self.cfg = m._field_cfg
        self.hw, self.table16 = hw, table16  # shared with a FusedTrainer (already fp16 / packed); else refreshed per update
        # Two frame pipelines on two streams: the tensor-core deformation net of one time frame runs beside the gather-bound hash-grid
        # pass of the previous one (different SM resources).  Random numbers are drawn on the host thread in the reference's order, so
        # the result does not depend on how the streams interleave.
        self.n_pipes = 2
        self.pipes = []
        for _ in range(self.n_pipes):
            self.pipes.append(dict(ws=F.FieldWorkspace(self.cfg, self.n_full, dev, training=False), xyzs=torch.empty(self.n_full, 3, device=dev),
                                   indices=torch.empty(self.n_full, dtype=torch.int32, device=dev), tmp=torch.full((self.n_full,), -1.0, device=dev),
                                   time_dev=torch.zeros(1, device=dev), stream=torch.cuda.Stream(device=dev)))

    def _weights(self):
        m = self.model
        if self.hw is not None:
            return self.hw, self.table16
        hw = m._half_weights()
        hw.refresh([w.detach() for w in m.mlp_weights()])
        return hw, m.encoder.embeddings.detach().to(torch.float16)

    def _query_and_store(self, pp, n, t, cas, hw, table16):
        """density at pp.xyzs[:n] (time pp.time_dev) -> tmp[indices] -> density_grid[t, cas] = max(grid * decay, tmp)"""
        m, cfg = self.model, self.cfg
        saved, cfg.density_scale = cfg.density_scale, 1.0  # like NeRFNetwork.density: the renderer applies density_scale itself
        try:
            F.field_density(cfg, hw, pp["ws"], pp["xyzs"], pp["time_dev"], table16, m.encoder.offsets, M=n, sigma_packed=True,
                            scatter=(pp["indices"], float(m.density_scale), pp["tmp"]), sigma_only=True)
        finally:
            cfg.density_scale = saved
        st = _lib.stream()
        _lib.call("seald_occ_ema_max", ptr(m.density_grid[t, cas]), ptr(pp["tmp"]), self.n_full, self._decay, st)

    @torch.no_grad()
    def update(self, decay=0.95):
        m = self.model
        if not m.cuda_ray:
            return
        H, n_full = self.H, self.n_full
        dev = m.density_grid.device
        hw, table16 = self._weights()
        self._decay = float(decay)
        half_time = 0.5 / m.time_size
        st = _lib.stream
        full = m.iter_density < 16
        if not full and m.iter_density >= 100:
            pass  # the reference stops re-sampling after 100 refreshes (only decay-free bookkeeping below)
        else:
            n_rand = n_full // 4
            if not full:
                # occupied-cell count of every frame in ONE host read (the reference synchronises per frame for `nonzero`, :509)
                occ_counts = (m.density_grid > 0).sum(-1).cpu()
            per = m.time_size // self.world_size
            if F.density_fused_ok(self.cfg):
                hw.pack_sigma()  # operand tiles of the sigma head for the one-launch density query
            main = torch.cuda.current_stream()
            for pp in self.pipes:
                pp["stream"].wait_stream(main)
            job = 0
            for t in range(self.rank * per, (self.rank + 1) * per):
                time = m.times[t]
                for cas in range(m.cascade):
                    bound = min(2 ** cas, m.bound)
                    half_cell = bound / H
                    pp = self.pipes[job % self.n_pipes]
                    job += 1
                    with torch.cuda.stream(pp["stream"]):
                        if full:
                            n = n_full
                            rnd = torch.rand(n, 3, device=dev)          # == torch.rand_like(cas_xyzs) of the reference loop
                            _lib.call("seald_occ_cell_points", None, ptr(rnd), n, H, float(bound - half_cell), float(half_cell), ptr(pp["xyzs"]),
                                      ptr(pp["indices"]), st())
                        else:
                            # the reference's draws, in its order and dtypes (same generator stream): cells, then ranks among the occupied
                            # cells of this frame, then the jitter of all 2n points (dnerf/renderer.py:507-524)
                            coords = torch.randint(0, H, (n_rand, 3), device=dev)
                            n_occ = int(occ_counts[t, cas])
                            if n_occ == 0:  # (the reference's randint(0, 0) raises here; an empty frame just gets no re-sampled cells)
                                n_occ = 1
                            rand_mask = torch.randint(0, n_occ, [n_rand], dtype=torch.long, device=dev)
                            csum = torch.cumsum(m.density_grid[t, cas] > 0, 0, dtype=torch.int32)
                            n = 2 * n_rand
                            rnd = torch.rand(n, 3, device=dev)
                            _lib.call("seald_occ_partial_points", ptr(coords), ptr(rand_mask), ptr(csum), n_full, ptr(rnd), n_rand, H,
                                      float(bound - half_cell), float(half_cell), ptr(pp["xyzs"]), ptr(pp["indices"]), st())
                        pp["time_dev"].copy_((time + (torch.rand_like(time) * 2 - 1) * half_time).reshape(-1)[:1])
                        self._query_and_store(pp, n, t, cas, hw, table16)
            for pp in self.pipes:
                main.wait_stream(pp["stream"])
            if self.world_size > 1:
                mine = m.density_grid[self.rank * per:(self.rank + 1) * per].clone()
                torch.distributed.all_gather_into_tensor(m.density_grid.view(-1), mine.view(-1), group=self.pg)
        m.mean_density = torch.mean(m.density_grid.clamp(min=0)).item()
        m.iter_density += 1
        density_thresh = min(m.mean_density, m.density_thresh)
        # density_grid [T, cascade, H^3] and density_bitfield [T, cascade * H^3 / 8] are contiguous: one launch packs every frame
        raymarching.packbits(m.density_grid.view(1, -1), density_thresh, m.density_bitfield.view(-1))
        total_step = min(16, m.local_step)
        if total_step > 0:
            m.mean_count = int(m.step_counter[:total_step, 0].sum().item() / total_step)
        m.local_step = 0
