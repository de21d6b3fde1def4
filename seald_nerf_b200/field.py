"""Low-level driver of the fused D-NeRF field kernels (csrc/field.cu + grid_encoder.cu).

`FieldWorkspace` owns every intermediate buffer of one forward/backward pass for up to `M` samples; `field_forward`
/ `field_backward` enqueue the kernel sequence

    deform MLP (freq-encode in the input stage)  ->  grid encoder  ->  sigma + colour heads (SH in the input stage)

and its reverse (heads bwd -> grid scatter + input grad -> deform bwd -> weight-gradient GEMMs).  Both the autograd
wrapper used by the drop-in `NeRFNetwork` and the graph-captured `FusedTrainer` call these two functions.
Reference: dnerf/network.py:123-169.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import ptr, F16, F32

DEFORM_W, DEFORM_K0, DEFORM_IN = 128, 80, 76
HEAD_W, HEAD_K0 = 64, 32


class FieldConfig:
    """Static description of the field (D-NeRF defaults, dnerf/network.py:11-26)."""

    def __init__(self, n_deform=8, n_sigma=2, n_color=3, bound=1.0, density_scale=1.0, grid_levels=16, grid_dim=2, grid_base=16,
                 grid_S=None, gridtype=0, align_corners=False, interp=0):
        self.n_deform, self.n_sigma, self.n_color = n_deform, n_sigma, n_color
        self.bound, self.density_scale = float(bound), float(density_scale)
        self.grid_levels, self.grid_dim, self.grid_base = grid_levels, grid_dim, grid_base
        self.grid_S = float(grid_S)
        self.gridtype, self.align_corners, self.interp = int(gridtype), bool(align_corners), int(interp)
        if grid_levels * grid_dim != HEAD_K0:
            raise NotImplementedError("fused heads expect 32 grid features (16 levels x 2)")


def half_weight_shapes(cfg):
    """(rows, cols, ld) of the fp16 staging copy of every nn.Linear weight, in the order deform, sigma, colour."""
    shapes = [(DEFORM_W, DEFORM_IN, DEFORM_K0)] + [(DEFORM_W, DEFORM_W, DEFORM_W)] * (cfg.n_deform - 2) + [(3, DEFORM_W, DEFORM_W)]
    shapes += [(HEAD_W, HEAD_K0, HEAD_K0)] + [(HEAD_W, HEAD_W, HEAD_W)] * (cfg.n_sigma - 2) + [(16, HEAD_W, HEAD_W)]
    shapes += [(HEAD_W, 31, HEAD_K0)] + [(HEAD_W, HEAD_W, HEAD_W)] * (cfg.n_color - 2) + [(3, HEAD_W, HEAD_W)]
    return shapes


class HalfWeights:
    """fp16, row-padded staging copies of the MLP weights (one flat buffer, 16-byte aligned slices)."""

    def __init__(self, cfg, device):
        self.cfg = cfg
        self.shapes = half_weight_shapes(cfg)
        sizes = [(r * ld + 7) // 8 * 8 for r, c, ld in self.shapes]
        self.flat = torch.zeros(sum(sizes), dtype=torch.float16, device=device)
        self.views, o = [], 0
        for (r, c, ld), n in zip(self.shapes, sizes):
            self.views.append(self.flat[o:o + r * ld].view(r, ld))
            o += n
        nd, ns = cfg.n_deform, cfg.n_sigma
        self.deform, self.sigma, self.color = self.views[:nd], self.views[nd:nd + ns], self.views[nd + ns:]
        self.p_deform, self.p_sigma, self.p_color = _lib.ptr_array(self.deform), _lib.ptr_array(self.sigma), _lib.ptr_array(self.color)

        # the deformation weights again, as canonical K-major UMMA operand tiles (tcgen05 path, csrc/field_umma.cu)
        self.packed_deform = torch.zeros(int(_lib.load().seald_field_umma_deform_bytes(cfg.n_deform)) // 2, dtype=torch.float16, device=device)
        self.packed_deform_T = torch.zeros(int(_lib.load().seald_field_umma_deform_bytes_T(cfg.n_deform)) // 2, dtype=torch.float16, device=device)
        self.pack_transposed = True  # the transposed tiles are only needed for training (deformation backward)
        # the sigma head as operand tiles for the one-launch density query (seald_field_density_umma); packed on demand (pack_sigma)
        self.packed_sigma = torch.zeros(max(int(_lib.load().seald_field_umma_sigma_bytes(cfg.n_sigma)) // 2, 8), dtype=torch.float16, device=device)

        n = len(self.views)
        self._dst = _lib.ptr_array(self.views)
        self._rows = (C.c_uint32 * n)(*[r for r, c, ld in self.shapes])
        self._cols = (C.c_uint32 * n)(*[c for r, c, ld in self.shapes])
        self._ld = (C.c_uint32 * n)(*[ld for r, c, ld in self.shapes])

    def refresh(self, weights32):
        """weights32: list of fp32 [out,in] tensors in the same order; one batched cast+pad launch."""
        ws = []
        for w, (r, c, ld) in zip(weights32, self.shapes):
            assert tuple(w.shape) == (r, c), (tuple(w.shape), (r, c))
            ws.append(w if w.is_contiguous() else w.contiguous())
        src = _lib.ptr_array(ws)
        _lib.call("seald_cast_pad_f16_batch", src, self._dst, self._rows, self._cols, self._ld, len(ws), _lib.stream())
        if self.pack_transposed:
            _lib.call("seald_field_umma_pack_deform_both", self.p_deform, self.cfg.n_deform, ptr(self.packed_deform), ptr(self.packed_deform_T),
                      _lib.stream())
        else:
            _lib.call("seald_field_umma_pack_deform", self.p_deform, self.cfg.n_deform, ptr(self.packed_deform), _lib.stream())


    def pack_sigma(self):
        """Operand tiles of the sigma head from the current fp16 staging copies (one tiny launch; the optimiser tail keeps the staging
        copies and the deformation tiles fresh, the sigma tiles are only needed by density queries)."""
        _lib.call("seald_field_umma_pack_sigma", self.p_sigma, self.cfg.n_sigma, ptr(self.packed_sigma), _lib.stream())


WGRAD_IMPL = os.environ.get("SEALD_WGRAD_IMPL", "umma")  # "umma": tcgen05 kernel (csrc/wgrad_umma.cu); "mma": mma.sync kernel (field.cu)
DEFORM_IMPL = os.environ.get("SEALD_DEFORM_IMPL", "umma")  # "umma": tcgen05 kernel (default); "mma": the mma.sync kernel of field.cu
TILE_ROWS = 128


def heads_tiled():
    """The heads save what the weight-gradient GEMMs read as 128-row tile images when the tcgen05 weight-gradient kernel consumes it
    (every job of that kernel then takes the bulk-copy path); the mma.sync kernel reads row-major tensors."""
    return WGRAD_IMPL == "umma"


def tile_image(x):
    """[M, w] row-major fp16 -> the 128-row TILE-IMAGE layout the tcgen05 deformation kernels save their activations in
    (csrc/field_umma.cu, csrc/wgrad_umma.cu): [tile][w / 8][128 rows][8 halves], returned as [ceil128(M), w] (rows beyond M zero)."""
    M, w = x.shape
    T = (M + TILE_ROWS - 1) // TILE_ROWS
    xp = x.new_zeros(T * TILE_ROWS, w)
    xp[:M] = x
    return xp.view(T, TILE_ROWS, w // 8, 8).permute(0, 2, 1, 3).contiguous().view(T * TILE_ROWS, w)


def from_tile_image(img, M=None):
    """Inverse of tile_image: [ceil128(M), w] tile images -> [M, w] row-major."""
    Mp, w = img.shape
    out = img.view(Mp // TILE_ROWS, w // 8, TILE_ROWS, 8).permute(0, 2, 1, 3).contiguous().view(Mp, w)
    return out if M is None else out[:M]


class FieldWorkspace:
    def __init__(self, cfg, M, device, training=True):
        self.cfg, self.M, self.training = cfg, int(M), training
        f16 = dict(dtype=torch.float16, device=device)
        f32 = dict(dtype=torch.float32, device=device)
        M = self.M
        # tcgen05 path: the deformation net's saved tensors hold whole 128-row tiles; the mma.sync kernels index [layer][M][width]
        Mp = (M + TILE_ROWS - 1) // TILE_ROWS * TILE_ROWS if DEFORM_IMPL == "umma" else M
        self.deform = torch.empty(M, 3, **f32)
        self.x01 = torch.empty(M, 3, **f32)
        self.feat = torch.empty(M, HEAD_K0, **f16)
        self.sigma = torch.empty(M, **f32)
        self.rgb = torch.empty(M, 3, **f32)
        if training:
            # tcgen05 path (DEFORM_IMPL "umma"): tile images (tile_image()); mma.sync path: row-major, first M rows
            self.in_buf = torch.empty(Mp, DEFORM_K0, **f16)
            self.fwd_d = torch.empty(cfg.n_deform - 1, Mp, DEFORM_W, **f16)
            self.bwd_d = torch.empty(cfg.n_deform - 1, Mp, DEFORM_W, **f16)
            self.gout_d = torch.empty(Mp, 16, **f16)
            self.hs = torch.empty(M, 16, **f16)
            # heads: tile images (whole 128-row tiles) with the tcgen05 weight-gradient kernel, row-major otherwise
            self.heads_tiled = heads_tiled()
            Mh = (M + TILE_ROWS - 1) // TILE_ROWS * TILE_ROWS if self.heads_tiled else M
            self.cin = torch.empty(Mh, HEAD_K0, **f16)
            self.feat_img = torch.empty(Mh, HEAD_K0, **f16) if self.heads_tiled else None  # tile-image copy of `feat` (sigma layer 0's A operand)
            self.fwd_s = torch.empty(cfg.n_sigma - 1, Mh, HEAD_W, **f16)
            self.fwd_c = torch.empty(cfg.n_color - 1, Mh, HEAD_W, **f16)
            self.bwd_s = torch.empty(cfg.n_sigma - 1, Mh, HEAD_W, **f16)
            self.bwd_c = torch.empty(cfg.n_color - 1, Mh, HEAD_W, **f16)
            self.gout_s = torch.empty(Mh, 16, **f16)
            self.gout_c = torch.empty(Mh, 16, **f16)
            self.dfeat = torch.empty(M, HEAD_K0, **f16)
            self.grad_x01 = torch.empty(M, 3, **f32)


def deform_forward(cfg, hw, xyzs, time_dev, M, m_dev, t0_mode, deform, x01, in_buf, fwd_buf):
    st = _lib.stream()
    if DEFORM_IMPL == "umma":
        _lib.call("seald_field_deform_forward_umma", ptr(xyzs), ptr(time_dev), ptr(hw.packed_deform), cfg.n_deform, M, ptr(m_dev), cfg.bound,
                  int(t0_mode), ptr(deform), ptr(x01), ptr(in_buf), ptr(fwd_buf), st)
    else:
        _lib.call("seald_field_deform_forward", ptr(xyzs), ptr(time_dev), hw.p_deform, cfg.n_deform, M, ptr(m_dev), cfg.bound, int(t0_mode),
                  ptr(deform), ptr(x01), ptr(in_buf), ptr(fwd_buf), st)


def deform_backward(cfg, hw, grad_x01, time_dev, M, m_dev, fwd_d, bwd_d, gout_d):
    st = _lib.stream()
    if DEFORM_IMPL == "umma":
        _lib.call("seald_field_deform_backward_umma", ptr(grad_x01), ptr(time_dev), ptr(hw.packed_deform_T), cfg.n_deform, M, ptr(m_dev),
                  cfg.bound, ptr(fwd_d), ptr(bwd_d), ptr(gout_d), st)
    else:
        _lib.call("seald_field_deform_backward", ptr(grad_x01), ptr(time_dev), hw.p_deform, cfg.n_deform, M, ptr(m_dev), cfg.bound, ptr(fwd_d),
                  ptr(bwd_d), ptr(gout_d), st)


def heads_forward(cfg, hw, ws, dirs, M, m_dev, save):
    """Sigma + colour heads on ws.feat -> ws.sigma, ws.rgb (+ the tensors the backward / weight gradients need when `save`)."""
    st = _lib.stream()
    if save and ws.heads_tiled:
        _lib.call("seald_field_heads_forward_tiled", ptr(ws.feat), ptr(dirs), hw.p_sigma, cfg.n_sigma, hw.p_color, cfg.n_color, M, ptr(m_dev),
                  cfg.density_scale, ptr(ws.sigma), ptr(ws.rgb), ptr(ws.hs), ptr(ws.cin), ptr(ws.fwd_s), ptr(ws.fwd_c), ptr(ws.feat_img), st)
    else:
        _lib.call("seald_field_heads_forward", ptr(ws.feat), ptr(dirs), hw.p_sigma, cfg.n_sigma, hw.p_color, cfg.n_color, M, ptr(m_dev),
                  cfg.density_scale, ptr(ws.sigma), ptr(ws.rgb), ptr(ws.hs) if save else None, ptr(ws.cin) if save else None,
                  ptr(ws.fwd_s) if save else None, ptr(ws.fwd_c) if save else None, st)


FUSE_GRID_HEADS = os.environ.get("SEALD_FUSE_GRID_HEADS", "1") != "0"  # measurement switch


def grid_heads_fusable(cfg, ws, save):
    """The encoder runs inside the heads kernel for the field's own grid shape (3-D, 16 levels x 2 features, fp16 table) in TRAINING
    (tile-image saves, tcgen05 weight gradients): at training-batch size the step is bound by kernel boundaries and the fused kernel
    saves one (0.0201 -> 0.0177 ms + the boundary).  Inference keeps the two kernels: on millions of samples the fused kernel's
    two-levels-in-flight gather (its register budget is shared with the MLP fragments) is slower (800x800 frame 4.13 vs 4.58 ms)."""
    return FUSE_GRID_HEADS and save and cfg.grid_dim == 2 and cfg.grid_levels == 16 and ws.heads_tiled


def grid_heads_forward(cfg, hw, ws, dirs, table16, offsets, M, m_dev, save):
    """ws.x01 -> hash-grid features -> sigma / colour heads -> ws.sigma, ws.rgb: ONE launch when grid_heads_fusable (the features stay
    in shared memory; ws.feat is not written), else encoder + heads."""
    st = _lib.stream()
    if grid_heads_fusable(cfg, ws, save):
        _lib.call("seald_field_grid_heads_forward", ptr(ws.x01), ptr(table16), ptr(offsets), 3, cfg.grid_dim, cfg.grid_levels, cfg.grid_S,
                  cfg.grid_base, cfg.gridtype, int(cfg.align_corners), cfg.interp, ptr(dirs), hw.p_sigma, cfg.n_sigma, hw.p_color, cfg.n_color,
                  M, ptr(m_dev), cfg.density_scale, ptr(ws.sigma), ptr(ws.rgb), ptr(ws.hs) if save else None, ptr(ws.cin) if save else None,
                  ptr(ws.fwd_s) if save else None, ptr(ws.fwd_c) if save else None, ptr(ws.feat_img) if save else None, st)
        return 1
    _lib.call("seald_grid_encode_forward", ptr(ws.x01), ptr(table16), ptr(offsets), ptr(ws.feat), None, M, 3, cfg.grid_dim, cfg.grid_levels,
              cfg.grid_S, cfg.grid_base, cfg.gridtype, int(cfg.align_corners), cfg.interp, F16, ptr(m_dev), st)
    heads_forward(cfg, hw, ws, dirs, M, m_dev, save)
    return 2


def heads_backward(cfg, hw, ws, grad_sigma, grad_rgb, M, m_dev):
    """dL/d(sigma, rgb) -> ws.dfeat (+ bwd_s, bwd_c, gout_s, gout_c for the weight gradients)."""
    name = "seald_field_heads_backward_tiled" if ws.heads_tiled else "seald_field_heads_backward"
    _lib.call(name, ptr(grad_sigma), ptr(grad_rgb), ptr(ws.rgb), ptr(ws.hs), hw.p_sigma, cfg.n_sigma, hw.p_color, cfg.n_color, M, ptr(m_dev),
              cfg.density_scale, ptr(ws.fwd_s), ptr(ws.fwd_c), ptr(ws.bwd_s), ptr(ws.bwd_c), ptr(ws.gout_s), ptr(ws.gout_c), ptr(ws.dfeat),
              _lib.stream())


def mlp_wgrad(jobs, n_jobs, M, m_dev):
    name = "seald_mlp_wgrad_umma" if WGRAD_IMPL == "umma" else "seald_mlp_wgrad"
    _lib.call(name, C.cast(jobs, C.c_void_p), n_jobs, M, ptr(m_dev), _lib.stream())


def field_forward(cfg, hw, ws, xyzs, dirs, time_dev, table16, offsets, m_dev=None, t0_mode=1, M=None):
    """Enqueue the forward pass for rows [0, M) of the workspace.  Results: ws.sigma, ws.rgb, ws.deform."""
    M = ws.M if M is None else int(M)
    st = _lib.stream()
    save = ws.training
    deform_forward(cfg, hw, xyzs, time_dev, M, m_dev, t0_mode, ws.deform, ws.x01, ws.in_buf if save else None, ws.fwd_d if save else None)
    # rows >= *m_dev keep stale x01: the grid kernel clamps nothing, so feed it only well-defined rows
    grid_heads_forward(cfg, hw, ws, dirs, table16, offsets, M, m_dev, save)


def _job(G, A, dW, N, K, ldg, lda, ldw, n_real, k_real):
    return _lib.WgradJob(G.data_ptr(), A.data_ptr(), dW.data_ptr(), N, K, ldg, lda, ldw, n_real, k_real)


def wgrad_jobs(cfg, ws, grads32, deform=True):
    """ctypes array of weight-gradient jobs; grads32: fp32 [out,in] gradient tensors in weight order."""
    nd, ns, nc = cfg.n_deform, cfg.n_sigma, cfg.n_color
    gd, gs, gc = grads32[:nd], grads32[nd:nd + ns], grads32[nd + ns:]
    jobs = []
    M = ws.M
    if deform:
        # the tcgen05 deformation kernels save tile images: leading dimensions 0 tell the weight-gradient kernel (bulk-copy path)
        tiled = DEFORM_IMPL == "umma"
        if tiled and WGRAD_IMPL != "umma":
            raise RuntimeError("SEALD_WGRAD_IMPL=mma reads row-major activations: use it with SEALD_DEFORM_IMPL=mma")
        for l in range(nd):
            last = l == nd - 1
            G = ws.gout_d if last else ws.bwd_d[l]
            A = ws.in_buf if l == 0 else ws.fwd_d[l - 1]
            jobs.append(_job(G, A, gd[l], 16 if last else DEFORM_W, DEFORM_K0 if l == 0 else DEFORM_W, 0 if tiled else (16 if last else DEFORM_W),
                             0 if tiled else (DEFORM_K0 if l == 0 else DEFORM_W), gd[l].shape[1], 3 if last else DEFORM_W,
                             DEFORM_IN if l == 0 else DEFORM_W))
    ht = ws.heads_tiled  # tile images: leading dimensions 0, sigma layer 0 reads the tile-image copy of the features
    for l in range(ns):
        last = l == ns - 1
        G = ws.gout_s if last else ws.bwd_s[l]
        A = (ws.feat_img if ht else ws.feat) if l == 0 else ws.fwd_s[l - 1]
        jobs.append(_job(G, A, gs[l], 16 if last else HEAD_W, HEAD_K0 if l == 0 else HEAD_W, 0 if ht else (16 if last else HEAD_W),
                         0 if ht else (HEAD_K0 if l == 0 else HEAD_W), gs[l].shape[1], 16 if last else HEAD_W, HEAD_K0 if l == 0 else HEAD_W))
    for l in range(nc):
        last = l == nc - 1
        G = ws.gout_c if last else ws.bwd_c[l]
        A = ws.cin if l == 0 else ws.fwd_c[l - 1]
        jobs.append(_job(G, A, gc[l], 16 if last else HEAD_W, HEAD_K0 if l == 0 else HEAD_W, 0 if ht else (16 if last else HEAD_W),
                         0 if ht else (HEAD_K0 if l == 0 else HEAD_W), gc[l].shape[1], 3 if last else HEAD_W, 31 if l == 0 else HEAD_W))
    arr = (_lib.WgradJob * len(jobs))(*jobs)
    return arr, len(jobs)


def field_backward(cfg, hw, ws, grad_sigma, grad_rgb, time_is_zero, table16, offsets, grad_table32, jobs, n_jobs, m_dev=None,
                   deform_grad=True, M=None, time_dev=None):
    """Enqueue the backward pass.  grad_table32 (fp32 [rows, C]) and the fp32 weight gradients referenced by `jobs`
    are ACCUMULATED into.  deform_grad=False: frozen deformation net (SealD student, SealDNeRF/utils.py:337-359).
    time_is_zero: host knowledge of t == 0 (skips the deformation backward); when the host does not know (graph replay)
    pass time_is_zero=False and time_dev: the kernel then zeroes the gradient itself if *time_dev == 0."""
    M = ws.M if M is None else int(M)
    st = _lib.stream()
    heads_backward(cfg, hw, ws, grad_sigma, grad_rgb, M, m_dev)
    want_dx = deform_grad and not time_is_zero
    _lib.call("seald_grid_encode_backward", ptr(ws.dfeat), ptr(ws.x01), ptr(table16), ptr(offsets), ptr(grad_table32), None,
              ptr(ws.grad_x01) if want_dx else None, M, 3, cfg.grid_dim, cfg.grid_levels, cfg.grid_S, cfg.grid_base, cfg.gridtype,
              int(cfg.align_corners), cfg.interp, F16, F32, ptr(m_dev), st)
    if want_dx:
        deform_backward(cfg, hw, ws.grad_x01, time_dev, M, m_dev, ws.fwd_d, ws.bwd_d, ws.gout_d)
    mlp_wgrad(jobs, n_jobs, M, m_dev)


# "split" (default): tcgen05 deformation kernel + encoder/sigma-head kernel.  "umma": sigma-only queries in ONE tcgen05 launch
# (seald_field_density_umma) — measured SLOWER above ~64k points (profiles/r2_occupancy.md: the tile groups of a CTA share the weight
# ring, so they run in lock step and all gather at the same time with the tensor pipe idle; 2^21 points 1.20-1.36 ms vs 0.97 ms),
# faster below (32k points: 0.027 vs 0.033 ms); kept as an opt-in with its parity test.
DENSITY_IMPL = os.environ.get("SEALD_DENSITY_IMPL", "split")


def density_fused_ok(cfg):
    return (DENSITY_IMPL == "umma" and DEFORM_IMPL == "umma" and cfg.grid_dim == 2 and cfg.grid_levels == 16 and cfg.n_sigma >= 2
            and cfg.n_deform >= 2)


def field_density(cfg, hw, ws, xyzs, time_dev, table16, offsets, M=None, sigma_packed=False, scatter=None, sigma_only=False):
    """Density-only query (NeRFNetwork.density, dnerf/network.py:171-208): deform -> grid -> sigma net.  scatter = (indices int32 [M],
    scale, tmp): also tmp[indices] = sigma * scale (occupancy refresh, dnerf/renderer.py:497-499).  sigma_only: the caller reads
    nothing but ws.sigma / the scatter (no ws.deform, ws.x01): the one-launch kernel may serve it.  sigma_packed: hw.pack_sigma() was
    called since the weights last changed."""
    M = ws.M if M is None else int(M)
    st = _lib.stream()
    if sigma_only and density_fused_ok(cfg):
        # ONE launch: the sigma head and the hash-grid gather ride the deformation kernel's tile pipeline (csrc/field_umma.cu, DENS)
        if not sigma_packed:
            hw.pack_sigma()
        idx, scale, tmp = scatter if scatter is not None else (None, 1.0, None)
        _lib.call("seald_field_density_umma", ptr(xyzs), ptr(time_dev), ptr(hw.packed_deform), cfg.n_deform, ptr(hw.packed_sigma), cfg.n_sigma,
                  ptr(table16), ptr(offsets), 3, cfg.grid_dim, cfg.grid_levels, cfg.grid_S, cfg.grid_base, cfg.gridtype, int(cfg.align_corners),
                  cfg.interp, M, None, cfg.bound, 2, cfg.density_scale, ptr(ws.sigma), ptr(idx) if idx is not None else None, float(scale),
                  ptr(tmp) if tmp is not None else None, st)
        return
    deform_forward(cfg, hw, xyzs, time_dev, M, None, 2, ws.deform, ws.x01, None, None)
    # encoder inside the density head: one launch, no feature round trip (occupancy refresh: partial pass 61.8 -> 53.5 ms, full sweep equal)
    if FUSE_GRID_HEADS and cfg.grid_dim == 2 and cfg.grid_levels == 16:
        _lib.call("seald_field_grid_sigma_forward", ptr(ws.x01), ptr(table16), ptr(offsets), 3, cfg.grid_dim, cfg.grid_levels, cfg.grid_S,
                  cfg.grid_base, cfg.gridtype, int(cfg.align_corners), cfg.interp, hw.p_sigma, cfg.n_sigma, M, cfg.density_scale, ptr(ws.sigma),
                  None, st)
        if scatter is not None:
            idx, scale, tmp = scatter
            _lib.call("seald_occ_store", ptr(ws.sigma), ptr(idx), M, float(scale), ptr(tmp), st)
        return
    _lib.call("seald_grid_encode_forward", ptr(ws.x01), ptr(table16), ptr(offsets), ptr(ws.feat), None, M, 3, cfg.grid_dim, cfg.grid_levels,
              cfg.grid_S, cfg.grid_base, cfg.gridtype, int(cfg.align_corners), cfg.interp, F16, None, st)
    _lib.call("seald_field_sigma_forward", ptr(ws.feat), hw.p_sigma, cfg.n_sigma, M, cfg.density_scale, ptr(ws.sigma), None, st)
    if scatter is not None:
        idx, scale, tmp = scatter
        _lib.call("seald_occ_store", ptr(ws.sigma), ptr(idx), M, float(scale), ptr(tmp), st)
