"""FusedRenderer — full-frame inference of the D-NeRF / SealD-teacher field (`run_cuda`, eval branch) with preallocated
buffers, and its ray-partitioned multi-GPU form.

Replaces the loop of dnerf/renderer.py:332-381 and SealDNeRF/renderer.py:214-286:

    while rays alive:  march_rays [+ Seal proxy mapping, fused]  ->  deform MLP -> grid encoder -> sigma/colour heads
                       [-> map_color on mapped samples]  ->  composite_rays (in place)  ->  compact alive list

The reference's schedule (n_step = clamp(N // n_alive, 1, 8)) with the cap raised to 32 — how a ray's samples are cut into rounds
does not change its compositing sequence, so the image is the same — stop when no ray is alive or `max_steps` is reached,
so images match the drop-in `NeRFRenderer.run_cuda`; what changes is the plumbing: no per-iteration allocation or
zero-fill (the march kernel writes the terminator slots itself), one fused field pass over exactly the live rows,
device-side order-preserving compaction, and NO host synchronisation inside the loop: n_alive / n_step live on the device
(`seald_render_schedule`), two rounds are replayed as one CUDA graph, and the host only polls a lagging copy of n_alive.

Multi-GPU: rays are independent, the model is replicated.  `render_sharded` gives every rank the rays of
interleaved tiles (tile t -> rank t % world; background/object load balances out), renders them locally with no
exchange, then all-gathers `image | depth | weights_sum` (5 floats per ray) and un-permutes.
"""
import ctypes as C

import torch

from . import _lib
from . import field as F
from . import parallel
from ._lib import ptr


class FusedRenderer:
    def __init__(self, model, max_rays, device=None, use_graph=True, max_n_step=None, min_samples=1 << 16):
        self.model = model
        self.device = device or model.encoder.embeddings.device
        if self.device.type != "cuda":
            raise RuntimeError("FusedRenderer needs a CUDA device (no CPU fallback)")
        self.cfg = model._field_cfg
        self.N = int(max_rays)
        # Sample slots per round.  The reference marches n_step = clamp(N // n_alive, 1, 8) samples per live ray, i.e. at most
        # N slots per round (dnerf/renderer.py:360); that is kept for full frames.  Small batches (training-size teacher
        # renders) would need ~15 nearly empty rounds that way, so the slot budget has a floor of `min_samples` and n_step may
        # grow to 32: the per-ray compositing sequence - and therefore the image - does not depend on how a ray's samples are
        # cut into rounds.
        # (measurement switches: SEALD_RENDER_SLOTS_MULT scales the slot budget of full frames, SEALD_RENDER_MAX_NSTEP caps n_step)
        import os
        # Sample-packed rounds (csrc/raymarch.cu k_march_round_pack): a round's samples are stored back to back, so the field never
        # sees the empty terminator rows of the n_step-rows-per-ray layout (22% of a frame's rows) and round 0 may march several
        # samples per ray although only `slots` rows exist (a CTA that does not fit defers its rays to the next round).
        # SEALD_RENDER_PACK=0: the reference's layout (measurement switch).
        self.pack = os.environ.get("SEALD_RENDER_PACK", "1") != "0"
        # Row budget of a round.  Fixed-stride layout: N (larger budgets cost more in empty rows than they save in rounds, measured).
        # Packed: 2 N — rows are real samples, so a larger budget only costs the samples a ray evaluates past its termination inside
        # its last round (none in the synthetic benchmark scene, where 1x / 1.5x / 2x / 3x give 3.55 / 3.49 / 3.37 / 3.33 ms per
        # frame in 10 / 8 / 6 / 5 rounds; a trained scene with opaque surfaces pays up to n_step - 1 samples per terminated ray).
        mult = float(os.environ.get("SEALD_RENDER_SLOTS_MULT", "2" if self.pack else "1"))
        self.slots = max(int(self.N * mult), int(min_samples))
        self.cap = self.slots + 128  # rounded up to the MLP tile
        self.use_graph = bool(use_graph)
        if max_n_step is None:
            # measured (800x800 frame): n_step up to 32 instead of the reference's 8 -> 11 rounds instead of 15 for +7% samples:
            # 4.50 -> 4.39 ms on one GPU, 3.09 -> 2.06 ms per frame on 4 (fewer latency-bound late rounds); a larger slot budget
            # (2x / 4x: 7 / 5 rounds) costs more in wasted samples than it saves in rounds
            max_n_step = 32
            if os.environ.get("SEALD_RENDER_MAX_NSTEP"):
                max_n_step = int(os.environ["SEALD_RENDER_MAX_NSTEP"])
        self.max_n_step = int(max_n_step)
        self.n_step0 = max(1, min(int(os.environ.get("SEALD_RENDER_NSTEP0", "4")), self.max_n_step))
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        N, cap = self.N, self.cap
        self.rays_o = torch.zeros(N, 3, **f32)
        self.rays_d = torch.zeros(N, 3, **f32)
        self.bitfield = torch.zeros(model.density_bitfield.shape[1], dtype=torch.uint8, device=dev)
        self.occ = torch.zeros(6, **f32)  # box of the occupied cells of the current frame (+ guard band): march early-out
        self.occ_scratch = torch.zeros(6 * model.cascade, **i32)
        self.nears = torch.empty(N, **f32)
        self.fars = torch.empty(N, **f32)
        self.rays_t = torch.empty(N, **f32)
        self.noises = torch.zeros(N, **f32)
        self.alive = [torch.empty(N, **i32), torch.empty(N, **i32)]
        self.ray_rows = torch.zeros(N, 2, **i32)  # packed rounds: {first row, count} of every alive-list entry
        # packed rounds, opt-in (SEALD_RENDER_COARSE=1): one bit per 8^3 block of the frame's bitfield (single cascade, H <= 128) lets the
        # march leave an empty block in one step — same samples bit for bit, but measured a wash on the benchmark frame (3.33 vs 3.38 ms on
        # one GPU, 1.06 vs 1.03 ms for an 8-way share: after the alive-list filter the rays start next to occupied cells)
        H = int(model.grid_size)
        self.coarse = (torch.zeros((H // 8) ** 3 // 32, **i32)
                       if (int(model.cascade) == 1 and H % 32 == 0 and H <= 128 and os.environ.get("SEALD_RENDER_COARSE", "0") == "1") else None)
        # packed rounds: (t, dt, dt_ray) of every marched sample before its row is known; n_alive * n_step <= slots, round 0: N * n_step0
        self.stage = torch.empty(3 * max(self.slots, N * self.n_step0) + 16, **f32) if self.pack else None
        # device-side loop state {n_alive, n_step, n_alive * n_step, steps done} and the compaction's output count
        self.state = torch.zeros(8, **i32)  # + [4] live samples evaluated so far, [5] non-empty rounds so far, [6] rows of this round (packed)
        self.counters = torch.zeros(2, **i32)  # survivors appended so far / CTAs finished (seald_composite_rays_compact; left zero)
        self.n_new = torch.zeros(1, **i32)
        self.scratch = torch.empty((N + 1023) // 1024 + 1, **i32)
        self.xyzs = torch.zeros(cap, 3, **f32)
        self.dirs = torch.zeros(cap, 3, **f32)
        self.deltas = torch.zeros(cap, 2, **f32)
        self.mask = torch.zeros(cap, dtype=torch.bool, device=dev)
        self.ws = F.FieldWorkspace(self.cfg, cap, dev, training=False)
        self.weights_sum = torch.empty(N, **f32)
        self.depth = torch.empty(N, **f32)
        self.image = torch.empty(N, 3, **f32)
        self.time = torch.zeros(1, **f32)
        self.h_state = torch.zeros(8, dtype=torch.int32).pin_memory()
        self._poll = [(torch.zeros(8, dtype=torch.int32).pin_memory(), torch.cuda.Event()) for _ in range(4)]
        self.hw = F.HalfWeights(self.cfg, dev)
        self.hw.pack_transposed = False  # inference only
        self.table16 = None
        self.refresh_weights()
        self._graphs = {}
        self.iterations = 0
        self.samples = 0
        self.launches = 0
        self.deferred = 0

    def refresh_weights(self):
        """Re-stage the fp16 copies of the model's parameters (call after the model was trained / loaded)."""
        m = self.model
        self.hw.refresh([w.detach() for w in m.mlp_weights()])
        t16 = m.encoder.embeddings.detach().to(torch.float16).contiguous()
        if self.table16 is None:
            self.table16 = t16
        else:
            self.table16.copy_(t16)  # (captured graphs hold this pointer)

    # ---- one round of the loop; every size is an upper bound, the live counts are read on the device ---------------------
    def _round(self, N, cur, first, opts, mapper, desc):
        m, cfg, st = self.model, self.cfg, _lib.stream()
        alive, nxt = self.alive[cur], self.alive[1 - cur]
        n_alive_dev, n_step_dev, m_dev = self.state[0:1], self.state[1:2], self.state[2:3]
        n_bound = N  # launch bound; kernels stop at *n_alive_dev
        # fixed stride: round 0 perturbs; packed: every round gets the buffer (a deferred ray is perturbed when it is first marched)
        noises = self.noises if (opts["perturb"] and (first or self.pack)) else None
        launches = 0
        if self.pack:
            m_dev = self.state[6:7]  # rows the march really wrote
            fused_map = mapper is not None and mapper.fusable
            _lib.call("seald_march_rays_pack", n_bound, self.max_n_step, ptr(alive), ptr(self.rays_t), ptr(self.rays_o), ptr(self.rays_d),
                      float(m.bound), opts["dt_gamma"], opts["max_steps"], int(m.cascade), int(m.grid_size), ptr(self.bitfield), ptr(self.fars),
                      ptr(self.xyzs), ptr(self.dirs), ptr(self.deltas), ptr(noises), ptr(self.state), self.cap, ptr(self.ray_rows),
                      ptr(self.stage), C.byref(desc) if fused_map else None, ptr(self.mask) if fused_map else None, ptr(self.occ),
                      ptr(self.coarse), st)
            if mapper is not None and not fused_map:  # anchor mapper: batch-wide early exit, separate op
                _lib.call("seald_seal_map_to_origin", C.byref(desc), ptr(self.xyzs), ptr(self.dirs), self.cap, ptr(m_dev), ptr(self.xyzs),
                          ptr(self.dirs), ptr(self.mask), ptr(mapper._dev_cache["scratch_i"]), st)
                launches += 3
            F.field_forward(cfg, self.hw, self.ws, self.xyzs, self.dirs, self.time, self.table16, m.encoder.offsets, m_dev, 1, M=self.cap)
            if mapper is not None and mapper.has_color_map:
                _lib.call("seald_seal_map_color", C.byref(mapper._dev_cache["color"]), ptr(self.xyzs), ptr(self.mask), ptr(self.ws.rgb), self.cap,
                          ptr(m_dev), ptr(mapper._dev_cache["scratch_f"]), st)
                launches += 4
            _lib.call("seald_composite_rays_pack", n_bound, opts["T_thresh"], ptr(alive), ptr(self.rays_t), ptr(self.ws.sigma), ptr(self.ws.rgb),
                      ptr(self.deltas), ptr(self.weights_sum), ptr(self.depth), ptr(self.image), ptr(nxt), ptr(self.state), ptr(self.counters),
                      ptr(self.ray_rows), max(N, self.slots), opts["max_steps"], self.max_n_step, self.cap, st)
            return launches + 5
        if mapper is not None and mapper.fusable:
            _lib.call("seald_march_rays_seal", n_bound, self.max_n_step, ptr(alive), ptr(self.rays_t), ptr(self.rays_o), ptr(self.rays_d), float(m.bound),
                      opts["dt_gamma"], opts["max_steps"], int(m.cascade), int(m.grid_size), ptr(self.bitfield), ptr(self.nears),
                      ptr(self.fars), ptr(self.xyzs), ptr(self.dirs), ptr(self.deltas), ptr(noises), ptr(n_alive_dev), ptr(n_step_dev),
                      C.byref(desc), ptr(self.mask), ptr(self.occ), st)
        else:
            _lib.call("seald_march_rays", n_bound, self.max_n_step, ptr(alive), ptr(self.rays_t), ptr(self.rays_o), ptr(self.rays_d), float(m.bound),
                      opts["dt_gamma"], opts["max_steps"], int(m.cascade), int(m.grid_size), ptr(self.bitfield), ptr(self.nears),
                      ptr(self.fars), ptr(self.xyzs), ptr(self.dirs), ptr(self.deltas), ptr(noises), ptr(n_alive_dev), ptr(n_step_dev),
                      ptr(self.occ), st)
            if mapper is not None:  # anchor mapper: batch-wide early exit, separate op
                _lib.call("seald_seal_map_to_origin", C.byref(desc), ptr(self.xyzs), ptr(self.dirs), self.cap, ptr(m_dev), ptr(self.xyzs),
                          ptr(self.dirs), ptr(self.mask), ptr(mapper._dev_cache["scratch_i"]), st)
                launches += 3
        F.field_forward(cfg, self.hw, self.ws, self.xyzs, self.dirs, self.time, self.table16, m.encoder.offsets, m_dev, 1, M=self.cap)
        if mapper is not None and mapper.has_color_map:
            _lib.call("seald_seal_map_color", C.byref(mapper._dev_cache["color"]), ptr(self.xyzs), ptr(self.mask), ptr(self.ws.rgb), self.cap,
                      ptr(m_dev), ptr(mapper._dev_cache["scratch_f"]), st)
            launches += 4
        # composite + survivor compaction + next round's schedule in one launch (the reference: composite_rays, a boolean-mask gather
        # with a host synchronisation, and Python bookkeeping)
        _lib.call("seald_composite_rays_compact", n_bound, self.max_n_step, opts["T_thresh"], ptr(alive), ptr(self.rays_t), ptr(self.ws.sigma),
                  ptr(self.ws.rgb), ptr(self.deltas), ptr(self.weights_sum), ptr(self.depth), ptr(self.image), ptr(nxt), ptr(self.state),
                  ptr(self.counters), max(N, self.slots), opts["max_steps"], self.max_n_step, st)
        return launches + 5

    def _double_round_graph(self, N, opts, mapper, desc):
        """Rounds 2k+1 and 2k+2 (alive buffers 1 -> 0 -> 1) as one CUDA graph; cached per ray count / options / mapper."""
        key = (N, tuple(sorted(opts.items())), id(mapper), None if mapper is None else id(mapper._dev_cache["desc"]))
        g = self._graphs.get(key)
        if g is None:
            saved = self.state.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):  # warm-up outside capture (function attributes, lazy module loads) on an empty state
                self.state.zero_()
                self._round(N, 1, False, opts, mapper, desc)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            self.state.zero_()
            with torch.cuda.graph(g):
                n = self._round(N, 1, False, opts, mapper, desc)
                n += self._round(N, 0, False, opts, mapper, desc)
            self.state.copy_(saved)
            g.launches = n
            self._graphs[key] = g
        return g

    @torch.no_grad()
    def render(self, rays_o, rays_d, time, bg_color=None, perturb=False, dt_gamma=0, max_steps=1024, T_thresh=None, normalize_depth=None,
               _packed_out=None, **kwargs):
        """rays_o, rays_d [..., 3] (<= max_rays rays); time [1,1] or float -> dict(image [...,3], depth [...], weights_sum [N]).

        The loop runs without host synchronisation: round sizes (n_alive, n_step = clamp(N // n_alive, 1, 8)) live on the
        device (seald_render_schedule), launches use upper bounds, and the host only polls a lagging copy of n_alive to
        know when to stop (extra rounds after the last ray died are empty launches)."""
        m = self.model
        prefix = rays_o.shape[:-1]
        N = rays_o.numel() // 3
        if N > self.N:
            raise RuntimeError("FusedRenderer was sized for %d rays, got %d" % (self.N, N))
        mapper = getattr(m, "seal_mapper", None)
        seald = hasattr(m, "init_mapper")  # SealD teacher/student renderers: T_thresh 1e-4, raw depth (SealDNeRF/renderer.py:114,284)
        if T_thresh is None:
            T_thresh = 1e-4 if seald else 1e-2
        if normalize_depth is None:
            normalize_depth = not seald
        if bg_color is None:
            bg_color = 1
        opts = {"perturb": bool(perturb), "dt_gamma": float(dt_gamma), "max_steps": int(max_steps), "T_thresh": float(T_thresh)}
        st = _lib.stream()
        if rays_o.data_ptr() != self.rays_o.data_ptr():  # (render_sharded gathers its tiles straight into the staging buffers)
            self.rays_o[:N].copy_(rays_o.reshape(-1, 3), non_blocking=True)
            self.rays_d[:N].copy_(rays_d.reshape(-1, 3), non_blocking=True)
        if torch.is_tensor(time):
            self.time.copy_(time.reshape(-1)[:1])
            t_idx = m._frame_index(self.time.view(1, 1))  # (device scalar -> index: the same host sync as the reference, renderer.py:285)
        else:
            self.time.fill_(float(time))
            t_idx = min(max(int(float(time) * m.time_size), 0), m.time_size - 1)
        self.bitfield.copy_(m.density_bitfield[t_idx], non_blocking=True)
        _lib.call("seald_occupancy_aabb", ptr(self.bitfield), int(m.cascade), int(m.grid_size), float(m.bound), 2, ptr(self.occ_scratch),
                  ptr(self.occ), st)
        if self.pack and self.coarse is not None:
            _lib.call("seald_occupancy_coarse_bits", ptr(self.bitfield), int(m.grid_size), ptr(self.coarse), st)
        aabb = m.aabb_train if m.training else m.aabb_infer
        nears, fars = self.nears[:N], self.fars[:N]
        ws_out, depth, image = self.weights_sum[:N], self.depth[:N], self.image[:N]
        if perturb:
            self.noises[:N].uniform_(0, 1)
        budget = max(N, self.slots)
        if self.pack:
            # one launch: near / far, rays_t, zeroed outputs, alive list of the rays that can meet an occupied cell, round-0 schedule
            _lib.call("seald_render_init_pack", ptr(self.rays_o), ptr(self.rays_d), ptr(aabb), ptr(self.occ), N, float(m.min_near), ptr(nears),
                      ptr(fars), ptr(self.rays_t), ptr(ws_out), ptr(depth), ptr(image), ptr(self.alive[0]), ptr(self.state), ptr(self.counters),
                      budget, self.n_step0, self.max_n_step, st)
        else:
            _lib.call("seald_near_far_from_aabb", ptr(self.rays_o), ptr(self.rays_d), ptr(aabb), N, float(m.min_near), ptr(nears), ptr(fars), st)
            self.rays_t[:N].copy_(nears)
            ws_out.zero_(); depth.zero_(); image.zero_()
            torch.arange(N, dtype=torch.int32, device=self.device, out=self.alive[0][:N])
            n_step0 = max(min(budget // N, self.max_n_step), 1)
            self.h_state.zero_()
            self.h_state[0] = N; self.h_state[1] = n_step0; self.h_state[2] = N * n_step0
            self.state.copy_(self.h_state, non_blocking=True)
        desc = mapper.descriptor(self.device) if mapper is not None else None

        launches = (7 if self.pack else 12) + self._round(N, 0, True, opts, mapper, desc)  # round 0 (perturbed start, alive 0 -> 1)
        rounds = 1
        graph = self._double_round_graph(N, opts, mapper, desc) if self.use_graph else None
        k = 0
        max_rounds = int(max_steps) + 2
        while rounds < max_rounds:
            if graph is not None:
                graph.replay()
                launches += graph.launches
            else:
                launches += self._round(N, 1, False, opts, mapper, desc) + self._round(N, 0, False, opts, mapper, desc)
            rounds += 2
            # lagging poll of n_alive: read the state copied one double-round ago, never wait for the newest one
            buf, ev = self._poll[k % 4]
            buf.copy_(self.state, non_blocking=True)
            ev.record()
            if k >= 1:
                pbuf, pev = self._poll[(k - 1) % 4]
                pev.synchronize()
                if int(pbuf[0]) == 0:
                    break
            k += 1
        self.h_state.copy_(self.state, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.iterations, self.samples, self.launches = int(self.h_state[5]), int(self.h_state[4]), launches
        self.deferred = int(self.h_state[2]) if self.pack else 0  # rays a packed round could not fit and marched one round later
        if isinstance(bg_color, (int, float)):
            # epilogue in one launch (background blend, depth normalisation, copies out of the persistent accumulators)
            if _packed_out is not None:  # tile-sharded frame: rows {r, g, b, depth, weights_sum} straight into the all-gather input
                _lib.call("seald_render_finish", ptr(image), ptr(ws_out), ptr(depth), ptr(nears), ptr(fars), N, float(bg_color),
                          int(bool(normalize_depth)), None, None, None, ptr(_packed_out), st)
                return {"packed": _packed_out}
            o_img, o_dep, o_ws = torch.empty_like(image), torch.empty_like(depth), torch.empty_like(ws_out)
            _lib.call("seald_render_finish", ptr(image), ptr(ws_out), ptr(depth), ptr(nears), ptr(fars), N, float(bg_color),
                      int(bool(normalize_depth)), ptr(o_img), ptr(o_dep), ptr(o_ws), None, st)
            return {"image": o_img.view(*prefix, 3), "depth": o_dep.view(*prefix), "weights_sum": o_ws}
        image = image + (1 - ws_out).unsqueeze(-1) * bg_color
        depth_out = torch.clamp(depth - nears, min=0) / (fars - nears) if normalize_depth else depth.clone()
        out = {"image": image.view(*prefix, 3), "depth": depth_out.view(*prefix), "weights_sum": ws_out.clone()}
        if _packed_out is not None:
            _packed_out[:N] = torch.cat([out["image"].view(-1, 3), out["depth"].view(-1, 1), out["weights_sum"].view(-1, 1)], dim=1)
            return {"packed": _packed_out}
        return out

    @torch.no_grad()
    def render_one_pass(self, rays_o, rays_d, time, bg_color=None, dt_gamma=0, max_steps=1024, T_thresh=None, normalize_depth=None, **kwargs):
        """Small-batch render (a training-size teacher batch of the SealD distillation step) WITHOUT the round loop: every ray is
        marched to the end in one launch (the training march, force_all_rays semantics, with the Seal proxy mapping fused in), the
        field is evaluated once over all samples, and one compositing kernel applies the same front-to-back recurrence with the same
        T_thresh early stop.  5 launches instead of ~4 rounds x 9.  Against render(): identical sample positions; the stop test sits
        after the sample here (composite_rays_train, raymarching.cu:560-566) and before the next one there (composite_rays, :880-886),
        so a ray can take one sample fewer, whose weight is < T_thresh: image / weights_sum agree to T_thresh.  Falls back to render()
        when the batch needs more sample slots than this renderer was sized for."""
        m = self.model
        prefix = rays_o.shape[:-1]
        N = rays_o.numel() // 3
        if N > self.N:
            raise RuntimeError("FusedRenderer was sized for %d rays, got %d" % (self.N, N))
        mapper = getattr(m, "seal_mapper", None)
        seald = hasattr(m, "init_mapper")
        if mapper is not None and not mapper.fusable:
            return self.render(rays_o, rays_d, time, bg_color=bg_color, dt_gamma=dt_gamma, max_steps=max_steps, T_thresh=T_thresh,
                               normalize_depth=normalize_depth, **kwargs)
        if T_thresh is None:
            T_thresh = 1e-4 if seald else 1e-2
        if normalize_depth is None:
            normalize_depth = not seald
        if bg_color is None:
            bg_color = 1
        st = _lib.stream()
        self.rays_o[:N].copy_(rays_o.reshape(-1, 3), non_blocking=True)
        self.rays_d[:N].copy_(rays_d.reshape(-1, 3), non_blocking=True)
        if torch.is_tensor(time):
            self.time.copy_(time.reshape(-1)[:1])
            t_idx = m._frame_index(self.time.view(1, 1))
        else:
            self.time.fill_(float(time))
            t_idx = min(max(int(float(time) * m.time_size), 0), m.time_size - 1)
        self.bitfield.copy_(m.density_bitfield[t_idx], non_blocking=True)
        _lib.call("seald_occupancy_aabb", ptr(self.bitfield), int(m.cascade), int(m.grid_size), float(m.bound), 2, ptr(self.occ_scratch),
                  ptr(self.occ), st)
        aabb = m.aabb_train if m.training else m.aabb_infer
        nears, fars = self.nears[:N], self.fars[:N]
        if getattr(self, "_rays", None) is None:
            self._rays = torch.zeros(self.N, 3, dtype=torch.int32, device=self.device)
            self._counter = torch.zeros(2, dtype=torch.int32, device=self.device)
            self._h_counter = torch.zeros(2, dtype=torch.int32).pin_memory()
            self._zero_noise = torch.zeros(self.N, dtype=torch.float32, device=self.device)
        self._counter.zero_()
        m_dev = self._counter[0:1]
        M = self.cap - 128
        args = (ptr(self.rays_o), ptr(self.rays_d), ptr(self.bitfield), float(m.bound), float(dt_gamma), int(max_steps), N, int(m.cascade),
                int(m.grid_size), M, None, None, ptr(aabb), float(m.min_near), ptr(nears), ptr(fars), ptr(self.xyzs), ptr(self.dirs),
                ptr(self.deltas), ptr(self._rays), ptr(self._counter), ptr(self._zero_noise))
        if mapper is not None:
            desc = mapper.descriptor(self.device)
            _lib.call("seald_march_rays_train_seal", *args, C.byref(desc), ptr(self.mask), ptr(self.occ), st)
        else:
            _lib.call("seald_march_rays_train", *args, ptr(self.occ), st)
        F.field_forward(self.cfg, self.hw, self.ws, self.xyzs, self.dirs, self.time, self.table16, m.encoder.offsets, m_dev, 1, M=M)
        launches = 8
        if mapper is not None and mapper.has_color_map:
            _lib.call("seald_seal_map_color", C.byref(mapper._dev_cache["color"]), ptr(self.xyzs), ptr(self.mask), ptr(self.ws.rgb), M,
                      ptr(m_dev), ptr(mapper._dev_cache["scratch_f"]), st)
            launches += 4
        ws_out, depth, image = self.weights_sum[:N], self.depth[:N], self.image[:N]
        _lib.call("seald_composite_rays_train_forward", ptr(self.ws.sigma), ptr(self.ws.rgb), ptr(self.deltas), ptr(self._rays), M, N,
                  float(T_thresh), ptr(ws_out), ptr(depth), ptr(image), st)
        self._h_counter.copy_(self._counter, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        n_samples = int(self._h_counter[0])
        if n_samples > M:  # not every ray found room for its samples: the round loop has no such limit
            return self.render(rays_o, rays_d, time, bg_color=bg_color, dt_gamma=dt_gamma, max_steps=max_steps, T_thresh=T_thresh,
                               normalize_depth=normalize_depth, **kwargs)
        self.iterations, self.samples, self.launches = 1, n_samples, launches
        # the training compositor measures depth from the ray's first sample parameter (its t starts at 0, raymarching.cu:538); the
        # inference compositor from the ray origin (t starts at rays_t = near, :869): add near * weights_sum
        depth = depth + nears * ws_out
        image = image + (1 - ws_out).unsqueeze(-1) * bg_color
        depth_out = torch.clamp(depth - nears, min=0) / (fars - nears) if normalize_depth else depth
        return {"image": image.view(*prefix, 3), "depth": depth_out.view(*prefix), "weights_sum": ws_out.clone()}

    @torch.no_grad()
    def render_sharded(self, rays_o, rays_d, time, rank=0, world_size=1, group=None, tile=256, **kwargs):
        """Every rank passes the SAME full ray set; each renders its interleaved tiles and all ranks get the full frame."""
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        if world_size == 1:
            return self.render(rays_o, rays_d, time, **kwargs)
        key = (N, world_size, rank, tile)
        if getattr(self, "_shard_key", None) != key:  # cached: building the index on the host costs more than the render
            self._shard_idx = parallel.shard_tiles(N, world_size, rank, tile).to(rays_o.device)
            self._shard_key = key
            cap = parallel.shard_capacity(N, world_size, tile)
            self._gather_in = torch.zeros(cap, 5, dtype=torch.float32, device=rays_o.device)   # rows past this rank's share stay zero
            self._gather_out = torch.empty(world_size * cap, 5, dtype=torch.float32, device=rays_o.device)
        idx = self._shard_idx
        n = idx.shape[0]
        # this rank's tiles gathered straight into the staging buffers; the render's epilogue writes {r, g, b, depth, weights_sum} rows
        # into the all-gather input; one index pass puts the gathered rows back into ray order (the outputs are VIEWS of that tensor)
        torch.index_select(rays_o, 0, idx, out=self.rays_o[:n])
        torch.index_select(rays_d, 0, idx, out=self.rays_d[:n])
        self.render(self.rays_o[:n], self.rays_d[:n], time, _packed_out=self._gather_in, **kwargs)
        torch.distributed.all_gather_into_tensor(self._gather_out, self._gather_in, group=group)
        full = self._gather_out.index_select(0, parallel.unshard_index(N, world_size, tile, rays_o.device))
        return {"image": full[:, :3], "depth": full[:, 3], "weights_sum": full[:, 4]}
