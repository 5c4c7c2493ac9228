"""numpy/ctypes front end of oracle/liboracle.so (mpi_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Parity pin: checked against outputs of the reference itself (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference run on CPU) in tests/test_oracle_golden.py.
Function names follow the reference functions they restate (see the C file for file:line).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NUM_CPU, NUM_CUDA = 0, 1

_lib = None


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when the reference tree is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(
            os.path.join(_HERE, "mpi_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_plane_homographies.restype = ctypes.c_int
        _lib.orc_render_tgt_rgb_depth.restype = ctypes.c_int
    return _lib


def ref_splat_lib():
    """The reference's own warping.c compiled from source by oracle/Makefile (None if absent)."""
    p = os.path.join(_HERE, "_ref", "libwarping_ref.so")
    return ctypes.CDLL(p) if os.path.exists(p) else None


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def plane_homographies(G, Kinv, Kt, depth, numerics=NUM_CPU):
    G, Kinv, Kt, depth = _f(G), _f(Kinv), _f(Kt), _f(depth)
    N = depth.size
    Hts = np.empty((N, 3, 3), np.float32)
    Hst = np.empty((N, 3, 3), np.float32)
    bad = lib().orc_plane_homographies(_p(G), _p(Kinv), _p(Kt), _p(depth), N, numerics, _p(Hts), _p(Hst))
    if bad:
        raise Exception("Matrix inverse contains nan!")
    return Hts, Hst


def get_src_xyz_from_plane_disparity(disparity, Kinv, H, W):
    disparity, Kinv = _f(disparity), _f(Kinv)
    B, S = disparity.shape
    out = np.empty((B, S, 3, H, W), np.float32)
    lib().orc_src_xyz_from_plane_disparity(_p(Kinv), _p(disparity), B, S, H, W, _p(out))
    return out


def get_tgt_xyz_from_plane_disparity(xyz_src, G):
    xyz_src, G = _f(xyz_src), _f(G)
    B, S, _, H, W = xyz_src.shape
    out = np.empty_like(xyz_src)
    lib().orc_tgt_xyz_from_plane_disparity(_p(xyz_src), _p(G), B, S, H, W, _p(out))
    return out


def homography_sample(src, Hst, numerics=NUM_CPU):
    src, Hst = _f(src), _f(Hst)
    N, C, H, W = src.shape
    tgt = np.empty_like(src)
    valid = np.empty((N, H, W), np.uint8)
    flow = np.empty((N, H, W, 2), np.float32)
    lib().orc_homography_sample(_p(src), _p(Hst), N, C, H, W, numerics, _p(tgt), _p(valid), _p(flow))
    return tgt, valid.astype(bool), flow


def homography_sample_inverse(Hts, H, W):
    Hts = _f(Hts)
    N = Hts.shape[0]
    flow = np.empty((N, H, W, 2), np.float32)
    lib().orc_homography_sample_inverse(_p(Hts), N, H, W, _p(flow))
    return flow


def plane_volume_rendering(rgb, sigma, xyz, src_flow=None, obj=None, is_bg_depth_inf=False):
    rgb, sigma, xyz = _f(rgb), _f(sigma), _f(xyz)
    B, S, _, H, W = sigma.shape
    src_flow = None if src_flow is None else _f(src_flow)
    obj = None if obj is None else _f(obj)
    rgb_out = np.empty((B, 3, H, W), np.float32)
    depth_out = np.empty((B, 1, H, W), np.float32)
    tacc = np.empty((B, S, 1, H, W), np.float32)
    wts = np.empty((B, S, 1, H, W), np.float32)
    flow_out = np.empty((B, 2, H, W), np.float32) if src_flow is not None else None
    obj_out = np.empty((B, 1, H, W), np.float32) if src_flow is not None else None
    lib().orc_plane_volume_rendering(_p(rgb), _p(sigma), _p(xyz), _p(src_flow), _p(obj), B, S, H, W,
                                     int(is_bg_depth_inf), _p(rgb_out), _p(depth_out), _p(tacc),
                                     _p(wts), _p(flow_out), _p(obj_out))
    return rgb_out, depth_out, tacc, wts, flow_out, obj_out


def plane_volume_rendering_flow(src_sigma, src_flow, src_xyz, hard_flow=False):
    src_sigma, src_flow, src_xyz = _f(src_sigma), _f(src_flow), _f(src_xyz)
    B, S, _, H, W = src_sigma.shape
    out = np.empty((B, 2, H, W), np.float32)
    lib().orc_plane_volume_rendering_flow(_p(src_sigma), _p(src_flow), _p(src_xyz), B, S, H, W,
                                          int(hard_flow), _p(out))
    return out


def alpha_composition(alpha, value):
    alpha, value = _f(alpha), _f(value)
    B, K, C, H, W = value.shape
    assert C <= 16
    comp = np.empty((B, C, H, W), np.float32)
    wts = np.empty((B, K, 1, H, W), np.float32)
    lib().orc_alpha_composition(_p(alpha), _p(value), B, K, C, H, W, _p(comp), _p(wts))
    return comp, wts


def render_tgt_rgb_depth(mpi_rgb, mpi_sigma, disparity, xyz_tgt, xyz_src, G, Kinv, Kt, obj,
                         is_bg_depth_inf=False, hard_flow=False, numerics=NUM_CPU):
    """-> (rgb [B,3,H,W], depth [B,1,H,W], mask/count [B,1,H,W], flowA2B [B,2,H,W], obj [B,1,H,W])"""
    mpi_rgb, mpi_sigma, disparity = _f(mpi_rgb), _f(mpi_sigma), _f(disparity)
    xyz_tgt, xyz_src, G, Kinv, Kt, obj = _f(xyz_tgt), _f(xyz_src), _f(G), _f(Kinv), _f(Kt), _f(obj)
    B, S, _, H, W = mpi_rgb.shape
    assert obj.shape == (B, S, 1, H, W) and G.shape == (B, 4, 4)
    o_rgb = np.empty((B, 3, H, W), np.float32)
    o_depth = np.empty((B, 1, H, W), np.float32)
    o_mask = np.empty((B, 1, H, W), np.float32)
    o_flow = np.empty((B, 2, H, W), np.float32)
    o_obj = np.empty((B, 1, H, W), np.float32)
    rc = lib().orc_render_tgt_rgb_depth(_p(mpi_rgb), _p(mpi_sigma), _p(disparity), _p(xyz_tgt),
                                        _p(xyz_src), _p(G), _p(Kinv), _p(Kt), _p(obj), B, S, H, W,
                                        int(is_bg_depth_inf), int(hard_flow), numerics, _p(o_rgb),
                                        _p(o_depth), _p(o_mask), _p(o_flow), _p(o_obj))
    if rc < 0:
        raise MemoryError("oracle allocation failed")
    if rc > 0:
        raise Exception("Matrix inverse contains nan!")
    return o_rgb, o_depth, o_mask, o_flow, o_obj


def render_frame(mpi, disparity, G, Kinv, Kt, obj_mask, numerics=NUM_CPU, **kw):
    """Whole frame from the kernel's own inputs: mpi [B,S,4,H,W], obj_mask [B,1,H,W]."""
    mpi = _f(mpi)
    B, S, _, H, W = mpi.shape
    xyz_src = get_src_xyz_from_plane_disparity(disparity, Kinv, H, W)
    xyz_tgt = get_tgt_xyz_from_plane_disparity(xyz_src, G)
    obj = np.ascontiguousarray(np.broadcast_to(_f(obj_mask)[:, None], (B, S, 1, H, W)))
    return render_tgt_rgb_depth(mpi[:, :, 0:3], mpi[:, :, 3:4], disparity, xyz_tgt, xyz_src, G, Kinv,
                                Kt, obj, numerics=numerics, **kw)


def forward_warping(src, idx, idy, z, use_ref=False):
    """src u8[h,w,3], idx/idy i64[h,w], z f32[h,w] -> warped u8[h,w,5]."""
    src = np.ascontiguousarray(src, np.uint8)
    idx = np.ascontiguousarray(idx, np.int64)
    idy = np.ascontiguousarray(idy, np.int64)
    z = _f(z)
    h, w = z.shape
    warped = np.zeros((h, w, 5), np.uint8)
    if use_ref:
        r = ref_splat_lib()
        if r is None:
            raise FileNotFoundError("oracle/_ref/libwarping_ref.so not built")
        r.forward_warping(_p(src), _p(idx), _p(idy), _p(z), _p(warped), ctypes.c_int(h), ctypes.c_int(w))
    else:
        lib().orc_forward_warping(_p(src), _p(idx), _p(idy), _p(z), _p(warped), h, w)
    return warped


def transformation_from_parameters(axisangle, translation, invert=False):
    a = _f(axisangle).reshape(-1, 3)
    t = _f(translation).reshape(-1, 3)
    M = np.empty((a.shape[0], 4, 4), np.float32)
    lib().orc_transformation_from_parameters(_p(a), _p(t), a.shape[0], int(invert), _p(M))
    return M
