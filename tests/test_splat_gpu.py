"""Bit-exactness of the parallel forward splat against the serial reference semantics
(external/forward_warping/warping.c:6-33): golden outputs of the reference's own C, the C
restatement, and the reference library compiled from source (oracle/_ref) when it travelled."""
import ctypes
import os

import numpy as np
import pytest
import torch

import synth
from oracle import mpi_oracle as orc

pytestmark = pytest.mark.gpu
PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mpi-flow_b200")


def splat_device(src, idx, idy, z, init=None):
    import _native as nat
    h, w = z.shape
    L = nat.lib()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_src, d_idx, d_idy, d_z = t(src), t(idx), t(idy), t(z)
    warped = torch.zeros((h, w, 5), dtype=torch.uint8, device="cuda") if init is None else t(init)
    nbytes = L.mpiflow_forward_warp_workspace_bytes(h, w)
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
    off = (-ws.data_ptr()) % 256
    nat.check(L.mpiflow_forward_warp_f32(ctypes.c_void_p(d_src.data_ptr()), ctypes.c_void_p(d_idx.data_ptr()),
                                         ctypes.c_void_p(d_idy.data_ptr()), ctypes.c_void_p(d_z.data_ptr()),
                                         ctypes.c_void_p(warped.data_ptr()), h, w,
                                         ctypes.c_void_p(ws.data_ptr() + off), nbytes, nat.stream(warped)))
    torch.cuda.synchronize()
    return warped.cpu().numpy()


@pytest.mark.parametrize("tag", ["dense", "pile", "identity", "row", "kitti_strip"])
def test_device_splat_matches_reference_golden(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "splat.npz"))
    h, w, seed = (int(v) for v in g[tag + "_meta"])
    src, idx, idy, z = synth.splat_inputs(h, w, seed, str(g[tag + "_mode"]))
    np.testing.assert_array_equal(splat_device(src, idx, idy, z), g[tag + "_warped"])


@pytest.mark.parametrize("h,w,mode,seed", [(384, 1280, "dense", 21), (384, 1280, "pile", 22), (480, 640, "dense", 23),
                                           (1080, 1920, "dense", 24), (7, 3, "pile", 25), (1, 1, "identity", 26),
                                           (64, 64, "onepixel", 27)])
def test_device_splat_matches_oracle_at_full_sizes(h, w, mode, seed):
    if mode == "onepixel":  # every source pixel hits ONE target: a single segment of h*w writers
        src, idx, idy, z = synth.splat_inputs(h, w, seed, "identity")
        idx[:] = 5
        idy[:] = 9
    else:
        src, idx, idy, z = synth.splat_inputs(h, w, seed, mode)
    want = orc.forward_warping(src, idx, idy, z)
    if orc.ref_splat_lib() is not None:
        np.testing.assert_array_equal(orc.forward_warping(src, idx, idy, z, use_ref=True), want)
    np.testing.assert_array_equal(splat_device(src, idx, idy, z), want)


@pytest.mark.parametrize("per", [2, 5, 15, 16, 17, 40])
def test_exactly_n_writers_per_target(per):
    """the chain path orders up to 16 writers of a target in registers and hands the call to the radix path beyond:
    both sides of that boundary, writers scattered over the raster (not adjacent), z with ties and 1000.0"""
    h, w = 48, 80
    src, _, _, z = synth.splat_inputs(h, w, 50 + per, "identity")
    rng = np.random.default_rng(per)
    tgt = rng.permutation(h * w) // per                      # every used target gets `per` writers (the last one fewer)
    idx, idy = (tgt % w).reshape(h, w).astype(np.int64), (tgt // w).reshape(h, w).astype(np.int64)
    np.testing.assert_array_equal(splat_device(src, idx, idy, z), orc.forward_warping(src, idx, idy, z))


def test_out_of_range_targets_are_dropped():
    """the reference would write out of bounds (undefined); both device paths drop such sources"""
    h, w = 24, 40
    src, idx, idy, z = synth.splat_inputs(h, w, 61, "dense")
    bad = np.random.default_rng(3).random((h, w)) < 0.1
    idx2, idy2 = idx.copy(), idy.copy()
    idx2[bad] = w + 5
    keep = ~bad
    want = np.zeros((h, w, 5), np.uint8)
    # the oracle on the kept sources only, in raster order: emulate by sending dropped sources nowhere
    dl = np.full((h, w), 1000.0, np.float32)
    for i in range(h):
        for j in range(w):
            if not keep[i, j]:
                continue
            x, y = int(idx[i, j]), int(idy[i, j])
            if z[i, j] < dl[y, x]:
                want[y, x, :3] = src[i, j]
            want[y, x, 3] = 1
            want[y, x, 4] = 0 if dl[y, x] != 1000.0 else 1
            dl[y, x] = z[i, j]
    np.testing.assert_array_equal(splat_device(src, idx2, idy2, z), want)


def test_untouched_bytes_keep_caller_content():
    h, w = 20, 30
    src, idx, idy, z = synth.splat_inputs(h, w, 31, "pile")
    init = np.random.default_rng(0).integers(0, 256, (h, w, 5), dtype=np.uint8)
    want = init.copy()
    orc.lib().orc_forward_warping(src.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p),
                                  idy.ctypes.data_as(ctypes.c_void_p), z.ctypes.data_as(ctypes.c_void_p),
                                  want.ctypes.data_as(ctypes.c_void_p), h, w)
    np.testing.assert_array_equal(splat_device(src, idx, idy, z, init=init), want)


def test_legacy_symbol_via_ctypes_like_moving_obj(monkeypatch):
    """moving_obj.py:12-13,127-129: cdll.LoadLibrary('external/forward_warping/libwarping.so') relative to
    the working directory, c_void_p(ndarray.ctypes.data) arguments."""
    monkeypatch.chdir(PKG)
    lib = ctypes.cdll.LoadLibrary("external/forward_warping/libwarping.so")
    warp = lib.forward_warping
    h, w = 48, 64
    src, idx, idy, z = synth.splat_inputs(h, w, 41, "dense")
    warped_arr = np.zeros(h * w * 5).astype(np.uint8)
    warp(ctypes.c_void_p(src.ctypes.data), ctypes.c_void_p(idx.ctypes.data), ctypes.c_void_p(idy.ctypes.data),
         ctypes.c_void_p(z.ctypes.data), ctypes.c_void_p(warped_arr.ctypes.data), ctypes.c_int(h), ctypes.c_int(w))
    np.testing.assert_array_equal(warped_arr.reshape(h, w, 5), orc.forward_warping(src, idx, idy, z))
