"""ctypes binding of libseald_b200.so (the C-ABI declared in include/seald_b200.h).

There is NO fallback: if the shared library is missing or was not built for sm_100a every op raises.
PyTorch only provides device memory and the current CUDA stream; kernels are ours.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libseald_b200.so")

F32, F16 = 0, 1

_vp, _u32, _i32, _f32 = C.c_void_p, C.c_uint32, C.c_int, C.c_float

# name -> argtypes (restype is always int except where noted)
_SIGS = {
    "seald_grid_encode_forward": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _i32, _vp, _vp],
    "seald_grid_encode_backward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _i32, _i32, _vp, _vp],
    "seald_grid_encode_backward_both": [_vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _i32, _i32, _vp, _vp, _vp],
    "seald_grid_encode_backward_table": [_vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _i32, _i32, _vp, _vp, _vp],
    "seald_grid_encode_backward_input": [_vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _i32, _vp, _vp],
    "seald_grid_debug_indices": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _vp],
    "seald_near_far_from_aabb": [_vp, _vp, _vp, _u32, _f32, _vp, _vp, _vp],
    "seald_sph_from_ray": [_vp, _vp, _f32, _u32, _vp, _vp],
    "seald_morton3D": [_vp, _u32, _vp, _vp],
    "seald_morton3D_invert": [_vp, _u32, _vp, _vp],
    "seald_packbits": [_vp, _u32, _f32, _vp, _vp],
    "seald_march_rays_train": [_vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_occupancy_aabb": [_vp, _u32, _u32, _f32, _i32, _vp, _vp, _vp],
    "seald_composite_rays_train_forward": [_vp, _vp, _vp, _vp, _u32, _u32, _f32, _vp, _vp, _vp, _vp],
    "seald_composite_rays_train_backward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _f32, _vp, _vp, _vp],
    "seald_composite_train_loss_fused": [_vp, _vp, _vp, _vp, _u32, _u32, _f32, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_march_rays": [_u32, _u32, _vp, _vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_composite_rays": [_u32, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_march_rays_pack": [_u32, _u32, _vp, _vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _vp, _vp,
                              _vp, _vp, _vp, _vp, _vp],
    "seald_occupancy_coarse_bits": [_vp, _u32, _vp, _vp],
    "seald_render_init_pack": [_vp, _vp, _vp, _vp, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _vp],
    "seald_render_finish": [_vp, _vp, _vp, _vp, _vp, _u32, _f32, _i32, _vp, _vp, _vp, _vp, _vp],
    "seald_composite_rays_pack": [_u32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _vp],
    "seald_composite_rays_compact": [_u32, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _vp],
    "seald_render_schedule": [_vp, _vp, _u32, _u32, _u32, _vp],
    "seald_compact_alive": [_vp, _u32, _vp, _vp, _vp, _vp, _vp],
    "seald_seal_map_to_origin": [_vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_seal_map_color": [_vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp],
    "seald_march_rays_seal": [_u32, _u32, _vp, _vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_march_rays_train_seal": [_vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                    _vp, _vp, _vp, _vp, _vp],
    "seald_freq_encode_forward": [_vp, _u32, _u32, _u32, _u32, _vp, _vp],
    "seald_freq_encode_backward": [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp],
    "seald_sh_encode_forward": [_vp, _vp, _u32, _u32, _u32, _vp, _vp],
    "seald_sh_encode_backward": [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp],
    "seald_field_deform_forward": [_vp, _vp, _vp, _i32, _u32, _vp, _f32, _i32, _vp, _vp, _vp, _vp, _vp],
    "seald_field_umma_pack_deform": [_vp, _i32, _vp, _vp],
    "seald_field_deform_forward_umma": [_vp, _vp, _vp, _i32, _u32, _vp, _f32, _i32, _vp, _vp, _vp, _vp, _vp],
    "seald_field_umma_pack_sigma": [_vp, _i32, _vp, _vp],
    "seald_field_density_umma": [_vp, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _u32, _vp, _f32, _i32,
                                 _f32, _vp, _vp, _f32, _vp, _vp],
    "seald_field_umma_pack_deform_T": [_vp, _i32, _vp, _vp],
    "seald_field_umma_pack_deform_both": [_vp, _i32, _vp, _vp, _vp],
    "seald_field_deform_backward_umma": [_vp, _vp, _vp, _i32, _u32, _vp, _f32, _vp, _vp, _vp, _vp],
    "seald_field_deform_backward": [_vp, _vp, _vp, _i32, _u32, _vp, _f32, _vp, _vp, _vp, _vp],
    "seald_field_heads_forward": [_vp, _vp, _vp, _i32, _vp, _i32, _u32, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_field_sigma_forward": [_vp, _vp, _i32, _u32, _f32, _vp, _vp, _vp],
    "seald_field_grid_heads_forward": [_vp, _vp, _vp, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _vp, _vp, _i32, _vp, _i32, _u32, _vp, _f32,
                                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_field_grid_sigma_forward": [_vp, _vp, _vp, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _vp, _i32, _u32, _f32, _vp, _vp, _vp],
    "seald_field_heads_forward_tiled": [_vp, _vp, _vp, _i32, _vp, _i32, _u32, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_field_heads_backward_tiled": [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _u32, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_field_heads_backward": [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _u32, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_mlp_wgrad": [_vp, _i32, _u32, _vp, _vp],
    "seald_mlp_wgrad_umma": [_vp, _i32, _u32, _vp, _vp],
    "seald_mlp_wgrad_umma_flag": [_vp, _i32, _u32, _vp, _vp, _vp],
    "seald_ffmlp_forward": [_vp, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp],
    "seald_ffmlp_backward": [_vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp],
    "seald_select_frame": [_vp, _u32, _vp, _u32, _vp, _vp, _vp, _vp, _vp],
    "seald_step_begin": [_vp, _u32, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _vp, _vp],
    "seald_mse_loss_bg": [_vp, _vp, _vp, _vp, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_l1_pretrain_loss": [_vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp],
    "seald_cast_pad_f16": [_vp, _vp, _u32, _u32, _u32, _vp],
    "seald_cast_pad_f16_batch": [_vp, _vp, _vp, _vp, _vp, _i32, _vp],
    "seald_dp_reduce_shard": [_vp, _vp, _i32, C.c_uint64, C.c_uint64, _vp, _vp],
    "seald_dp_adam_weights": [_vp, _vp, _i32, _vp, _vp, _vp, C.c_uint64, C.c_uint64, C.c_uint64, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp],
    "seald_dp_adam_shard_broadcast": [_vp, _vp, _i32, _vp, _vp, _vp, _vp, C.c_uint64, C.c_uint64, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _vp],
    "seald_loss_scale_update_stash": [_vp, _vp, _vp, _f32, _f32, _i32, _vp, _vp, _vp],
    "seald_occ_cell_points": [_vp, _vp, _u32, _u32, _f32, _f32, _vp, _vp, _vp],
    "seald_occ_partial_points": [_vp, _vp, _vp, _u32, _vp, _u32, _u32, _f32, _f32, _vp, _vp, _vp],
    "seald_occ_store": [_vp, _vp, _u32, _f32, _vp, _vp],
    "seald_occ_ema_max": [_vp, _vp, _u32, _f32, _vp],
    "seald_get_rays_gather": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_trunc_exp_forward": [_vp, _vp, C.c_uint64, _vp],
    "seald_trunc_exp_backward": [_vp, _vp, _vp, C.c_uint64, _vp],
    "seald_grad_finite_check": [_vp, C.c_uint64, _vp, _vp],
    "seald_adam_advance": [_vp, _vp, _vp],
    "seald_adam_step": [_vp, _vp, _vp, _vp, C.c_uint64, _f32, _f32, _f32, _f32, _u32, _vp, _vp, _vp, _vp, _i32, _vp],
    "seald_adam_step_ex": [_vp, _vp, _vp, _vp, C.c_uint64, _f32, _f32, _f32, _f32, _u32, _vp, _vp, _vp, _vp, _i32, _u32, _vp],
    "seald_adam_step_lr": [_vp, _vp, _vp, _vp, C.c_uint64, _f32, _vp, _f32, _f32, _f32, _u32, _vp, _vp, _vp, _vp, _i32, _u32, _vp],
    "seald_mlp_tail_dp": [_vp, _i32, C.c_uint64, C.c_uint64, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _f32, _f32,
                          _i32, _vp, _vp, _vp, _i32, _vp, _vp],
    "seald_optimizer_step": [_vp, _vp, _vp, _vp, _vp, _i32, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _f32, _f32, _i32, _vp, _vp, _vp, _i32, _vp,
                             _vp, _vp, _vp, _vp, C.c_uint64, _f32, _vp, _i32, _vp],
    "seald_mlp_tail": [_vp, _vp, _vp, _vp, _vp, _i32, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _f32, _f32, _i32, _vp, _vp, _vp, _i32, _vp, _vp],
    "seald_ema_update": [_vp, _vp, C.c_uint64, _f32, _vp],
    "seald_umma_probe": [_i32, _i32, _i32, _i32, _vp, _vp],
    "seald_loss_scale_update": [_vp, _vp, _vp, _f32, _f32, _i32, _vp, _vp],
}


class TailSeg(C.Structure):
    """seald_tail_seg of include/seald_b200.h."""
    _fields_ = [("first", C.c_uint32), ("rows", C.c_uint32), ("cols", C.c_uint32), ("ld", C.c_uint32), ("dst16", C.c_void_p),
                ("packed", C.c_void_p), ("n_pad", C.c_uint32), ("packedT", C.c_void_p)]


class WgradJob(C.Structure):
    """seald_wgrad_job of include/seald_b200.h."""
    _fields_ = [("G", C.c_void_p), ("A", C.c_void_p), ("dW", C.c_void_p), ("N", C.c_int), ("K", C.c_int), ("ldg", C.c_int),
                ("lda", C.c_int), ("ldw", C.c_int), ("n_real", C.c_int), ("k_real", C.c_int)]


def ptr_array(tensors):
    """Host array of device pointers (const void* const*)."""
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr

_lib = None


def exported_symbols():
    """Every entry point include/seald_b200.h declares (used by the CPU-side ABI test)."""
    return ["seald_version", "seald_sm_arch", "seald_strerror", "seald_field_umma_deform_bytes", "seald_field_umma_deform_bytes_T",
            "seald_field_umma_sigma_bytes"] + list(_SIGS)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libseald_b200.so not found at %s — build it with `python -m seald_nerf_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.seald_version.restype = C.c_int
    lib.seald_sm_arch.restype = C.c_int
    lib.seald_strerror.restype = C.c_char_p
    lib.seald_strerror.argtypes = [C.c_int]
    lib.seald_field_umma_deform_bytes.restype = C.c_uint64
    lib.seald_field_umma_deform_bytes.argtypes = [C.c_int]
    lib.seald_field_umma_sigma_bytes.restype = C.c_uint64
    lib.seald_field_umma_sigma_bytes.argtypes = [C.c_int]
    lib.seald_field_umma_deform_bytes_T.restype = C.c_uint64
    lib.seald_field_umma_deform_bytes_T.argtypes = [C.c_int]
    for name, sig in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = sig
        fn.restype = C.c_int
    if lib.seald_sm_arch() != 100:
        raise RuntimeError("libseald_b200.so was not built for sm_100a")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().seald_strerror(rc).decode()
        raise RuntimeError("seald_b200 %s failed: %s (status %d)" % (what, msg, rc))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("seald_b200 ops need CUDA tensors (no CPU fallback)")


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    check(rc, name)
