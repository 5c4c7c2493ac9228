"""ctypes binding of libmpiflow_sm100.so (include/mpiflow.h) for the Python host layer.

There is deliberately no CPU or torch fallback here: if the CUDA library is missing or a tensor is
not on a CUDA device, the call fails loudly.  torch is used only for device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# MPIFLOW_LIB: development override (A/B builds of the same library, tools/ab.py)
LIB_PATH = os.environ.get("MPIFLOW_LIB") or os.path.join(_HERE, "libmpiflow_sm100.so")

BG_DEPTH_INF = 1
HARD_FLOW = 2
E_SINGULAR = -3

_lib = None
_c = ctypes
_P, _I, _I64, _U, _SZ = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_uint, _c.c_size_t

_SIGNATURES = {
    "mpiflow_version": (_I, []),
    "mpiflow_shutdown": (_I, []),
    "mpiflow_error_string": (_c.c_char_p, [_I]),
    "mpiflow_plane_homographies_f32": (_I, [_P, _P, _P, _P, _I, _P, _P, _P, _P]),
    "mpiflow_render_workspace_bytes": (_SZ, [_I, _I]),
    "mpiflow_render_workspace_bytes_hw": (_SZ, [_I, _I, _I, _I]),
    "mpiflow_render_fused_f32": (_I, [_P, _I64, _I64, _I64, _P, _I64, _I64, _P, _I64, _I64, _P, _P, _P, _P,
                                      _I, _I, _I, _I, _U, _P, _P, _P, _P, _P, _P, _SZ, _P, _P]),
    "mpiflow_render_fused_f16": (_I, [_P, _I64, _I64, _I64, _P, _I64, _I64, _P, _I64, _I64, _P, _P, _P, _P,
                                      _I, _I, _I, _I, _U, _P, _P, _P, _P, _P, _P, _SZ, _P, _P]),
    "mpiflow_debug_div_f32": (_I, [_P, _P, _P, _I, _P, _P, _P]),
    "mpiflow_debug_packed_f32": (_I, [_P, _P, _P, _I, _P, _P]),
    "mpiflow_render_fused_host": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _U, _P, _P, _P, _P, _P]),
    "mpiflow_render_fused_host_f16": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _U, _P, _P, _P, _P, _P]),
    "mpiflow_homography_sample_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "mpiflow_homography_flow_f32": (_I, [_P, _I, _I, _I, _P, _P]),
    "mpiflow_src_xyz_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "mpiflow_tgt_xyz_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "mpiflow_plane_volume_rendering_f32": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _U, _P, _P, _P, _P, _P, _P, _P]),
    "mpiflow_plane_volume_rendering_flow_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _U, _P, _P]),
    "mpiflow_alpha_composition_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "forward_warping": (None, [_P, _P, _P, _P, _P, _I, _I]),
    "mpiflow_forward_warp_host": (_I, [_P, _P, _P, _P, _P, _I, _I]),
    "mpiflow_forward_warp_workspace_bytes": (_SZ, [_I, _I]),
    "mpiflow_forward_warp_f32": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _SZ, _P]),
    "mpiflow_rasterize_workspace_bytes": (_SZ, [_I, _I, _I]),
    "mpiflow_blend_src_f32": (_I, [_P, _I64, _I64, _I64, _P, _I64, _I64, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "mpiflow_blend_src_f16": (_I, [_P, _I64, _I64, _I64, _P, _I64, _I64, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "mpiflow_assemble_sample_f32": (_I, [_P, _P, _P, _P, _P, _P, _P, _c.c_float, _I, _I, _P, _P, _P, _P]),
    "mpiflow_inpaint_ns_workspace_bytes": (_SZ, [_I, _I, _I]),
    "mpiflow_inpaint_ns_u8": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _SZ, _P]),
    "mpiflow_rasterize_mesh_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _c.c_float, _P, _P, _P, _P, _P, _SZ, _P]),
}
EXPORTS = tuple(_SIGNATURES)


def lib():
    """The loaded library; raises if it has not been built (python mpi-flow_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the MPI renderer has no CPU fallback. "
                "Build it with `python mpi-flow_b200/build.py` (needs nvcc, targets sm_100a).")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


class MpiflowError(RuntimeError):
    pass


def check(rc):
    if rc == 0:
        return
    if rc == E_SINGULAR:
        raise Exception("Matrix inverse contains nan!")  # same text as homography_sampler.py:26
    raise MpiflowError(f"libmpiflow_sm100: {lib().mpiflow_error_string(rc).decode()} (code {rc})")


def dev(t, name="tensor"):
    """fp32, CUDA, returns the tensor unchanged otherwise (callers decide about contiguity)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: libmpiflow_sm100 is CUDA-only (no CPU fallback)")
    return t if t.dtype == torch.float32 else t.float()


def dev_storage(t, name="tensor"):
    """like dev(), but an fp16 tensor stays fp16 (the *_f16 entry points read halves); anything else becomes fp32"""
    if isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float16:
        return t
    return dev(t, name)


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def inner_contiguous(t):
    """True if the last two dims (H, W) are dense so that outer strides can be passed to the kernel."""
    return t.stride(-1) == 1 and t.stride(-2) == t.shape[-1]
