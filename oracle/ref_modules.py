"""Loader of the UNMODIFIED reference hot-path modules -- TEST INFRASTRUCTURE ONLY.

oracle/Makefile copies utils/mpi/{homography_sampler,mpi_rendering,rendering_utils}.py and geometry.py of
Sharpiless/MPI-Flow, byte for byte, into oracle/_ref/refsrc/ (git-ignored; it travels to the GPU box with the
snapshot, where /root/reference does not exist).  This module imports those files under private module names, so
that the reference's own `from utils.mpi... import` lines resolve to the reference's files and not to the
product's modules of the same names, and leaves sys.modules / sys.path as it found them.

On CUDA tensors the reference runs unshimmed (its inverse() hard-codes .cuda(), homography_sampler.py:15-17); on CPU
tensors -- CPU-only box, or the CPU reference arm of bench.py on the GPU box -- the two calls are shimmed for the
duration of a call with cpu_shims().

Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs may import this.
"""
import contextlib
import importlib
import os
import sys

import torch

REFSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "refsrc")
_NAMES = ("utils", "utils.mpi", "utils.mpi.homography_sampler", "utils.mpi.rendering_utils", "utils.mpi.mpi_rendering",
          "geometry")
_cache = None


def available():
    return os.path.isfile(os.path.join(REFSRC, "utils", "mpi", "mpi_rendering.py"))


def load():
    """-> (mpi_rendering, homography_sampler, rendering_utils, geometry): the reference's own modules."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError(f"{REFSRC} is missing: run `make -C oracle ref` in the build container (needs /root/reference)")
    saved_modules = {n: sys.modules.pop(n) for n in _NAMES if n in sys.modules}
    saved_path = list(sys.path)
    # only the reference's files may answer to `utils.*` / `geometry` while they import each other
    sys.path[:] = [REFSRC] + [p for p in saved_path if not os.path.isdir(os.path.join(p, "utils", "mpi"))
                              and not os.path.isfile(os.path.join(p, "geometry.py"))]
    importlib.invalidate_caches()
    try:
        mods = tuple(importlib.import_module(n) for n in ("utils.mpi.mpi_rendering", "utils.mpi.homography_sampler",
                                                          "utils.mpi.rendering_utils", "geometry"))
        for m in mods:
            assert os.path.abspath(m.__file__).startswith(REFSRC), m.__file__
    finally:
        for n in _NAMES:
            sys.modules.pop(n, None)
        sys.modules.update(saved_modules)
        sys.path[:] = saved_path
        importlib.invalidate_caches()
    _cache = mods
    return mods


@contextlib.contextmanager
def cpu_shims(active=None):
    """The reference's inverse() calls .cuda() and torch.cuda.synchronize() (homography_sampler.py:15-17); for a run on
    CPU tensors both are made no-ops -- on a CPU-only box, and on the GPU box when the reference's CPU path is what is
    being timed (bench.py --impl reference).  active=None: shim iff there is no CUDA device."""
    if active is None:
        active = not torch.cuda.is_available()
    if not active:
        yield
        return
    sync, cuda = torch.cuda.synchronize, torch.Tensor.cuda
    torch.cuda.synchronize = lambda *a, **k: None
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.cuda.synchronize, torch.Tensor.cuda = sync, cuda


def render_frame(mpi, disparity, G, Kinv, K, obj_mask=None, is_bg_depth_inf=False, hard_flow=False):
    """The reference's own call sequence of one render (utils/utils.py:304-347 without the dead mpi_flow_src):
    get_src_xyz_from_plane_disparity -> get_tgt_xyz_from_plane_disparity -> render_tgt_rgb_depth, on whatever device
    the tensors live.  mpi BxSx4xHxW, disparity BxS, G Bx4x4, Kinv / K Bx3x3, obj_mask Bx1xHxW.
    -> (rgb, depth, mask, flowA2B, obj)."""
    R, HS, _, _ = load()
    B, S, _, H, W = mpi.shape
    with cpu_shims(active=mpi.device.type == "cpu"):
        sampler = HS.HomographySample(H, W, mpi.device)
        xyz_src = R.get_src_xyz_from_plane_disparity(sampler.meshgrid, disparity, Kinv)
        xyz_tgt = R.get_tgt_xyz_from_plane_disparity(xyz_src, G)
        obj = None if obj_mask is None else obj_mask.unsqueeze(1).repeat(1, S, 1, 1, 1)
        return R.render_tgt_rgb_depth(sampler, mpi[:, :, 0:3], mpi[:, :, 3:], disparity, xyz_tgt, xyz_src, G, Kinv, K,
                                      None, use_alpha=False, is_bg_depth_inf=is_bg_depth_inf, hard_flow=hard_flow,
                                      obj_mask=obj)
