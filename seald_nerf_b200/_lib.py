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
    "seald_grid_encode_forward": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _i32, _vp],
    "seald_grid_encode_backward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _i32, _i32, _vp],
    "seald_grid_debug_indices": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _vp],
    "seald_near_far_from_aabb": [_vp, _vp, _vp, _u32, _f32, _vp, _vp, _vp],
    "seald_sph_from_ray": [_vp, _vp, _f32, _u32, _vp, _vp],
    "seald_morton3D": [_vp, _u32, _vp, _vp],
    "seald_morton3D_invert": [_vp, _u32, _vp, _vp],
    "seald_packbits": [_vp, _u32, _f32, _vp, _vp],
    "seald_march_rays_train": [_vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_composite_rays_train_forward": [_vp, _vp, _vp, _vp, _u32, _u32, _f32, _vp, _vp, _vp, _vp],
    "seald_composite_rays_train_backward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _f32, _vp, _vp, _vp],
    "seald_march_rays": [_u32, _u32, _vp, _vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_composite_rays": [_u32, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "seald_compact_alive": [_vp, _u32, _vp, _vp, _vp, _vp, _vp],
    "seald_freq_encode_forward": [_vp, _u32, _u32, _u32, _u32, _vp, _vp],
    "seald_freq_encode_backward": [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp],
    "seald_sh_encode_forward": [_vp, _vp, _u32, _u32, _u32, _vp, _vp],
    "seald_sh_encode_backward": [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp],
}

_lib = None


def exported_symbols():
    """Every entry point include/seald_b200.h declares (used by the CPU-side ABI test)."""
    return ["seald_version", "seald_sm_arch", "seald_strerror"] + list(_SIGS)


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
