"""CPU-side checks: the C-ABI library loads and exports everything include/mpiflow.h declares, the
Python layer mirrors the reference's interface, and host-side helpers agree with the reference."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mpiflow.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpiflow_\w+|forward_warping)\s*\(", text)))


def test_header_symbols_are_exported():
    import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mpiflow.h but not exported"
    assert sorted(_native.EXPORTS) == syms
    assert _native.lib().mpiflow_version() == 100
    assert b"nan" in _native.lib().mpiflow_error_string(-3)


def test_dropin_libwarping_exports_legacy_symbol():
    p = os.path.join(ROOT, "mpi-flow_b200", "external", "forward_warping", "libwarping.so")
    assert hasattr(ctypes.CDLL(p), "forward_warping")


def test_argument_errors_do_not_need_a_gpu():
    import _native
    L = _native.lib()
    assert L.mpiflow_render_workspace_bytes(2, 64) >= 2 * 64 * 128
    assert L.mpiflow_render_workspace_bytes(0, 64) == 0
    assert L.mpiflow_forward_warp_workspace_bytes(384, 1280) > 5 * 4 * 384 * 1280
    assert L.mpiflow_homography_flow_f32(None, 1, 4, 4, None, None) == -1
    assert L.mpiflow_forward_warp_host(None, None, None, None, None, 0, 5) == 0  # empty image: no-op
    z = np.zeros(64, np.float32)
    zp = ctypes.c_void_p(z.ctypes.data)
    # MPIFLOW_POSES_PER_MPI is a device-entry flag: the host entry rejects it before touching the device
    assert L.mpiflow_render_fused_host(zp, None, zp, zp, zp, zp, 2, 1, 2, 4, (2 - 1) << 8, zp, zp, zp, zp, None) == -1
    assert L.mpiflow_blend_src_f32(None, 0, 0, 0, None, 0, 0, None, None, None, 1, 1, 2, 4, None, None) == -1
    assert L.mpiflow_assemble_sample_f32(zp, zp, zp, zp, zp, zp, zp, 0.99, 0, 4, zp, zp, zp, None) == -1


def test_python_interface_mirrors_reference_signatures():
    from utils.mpi import mpi_rendering as R, rendering_utils as U
    from utils.mpi.homography_sampler import HomographySample, inverse  # noqa: F401
    import geometry
    want = {
        R.render: ["rgb_BS3HW", "sigma_BS1HW", "xyz_BS3HW", "src_sigma_BS1HW", "src_flow_BS2HW", "src_xyz_BS3HW",
                   "use_alpha", "is_bg_depth_inf", "hard_flow", "obj_mask"],
        R.alpha_composition: ["alpha_BK1HW", "value_BKCHW"],
        R.plane_volume_rendering: ["rgb_BS3HW", "sigma_BS1HW", "xyz_BS3HW", "src_sigma_BS1HW", "src_flow_BS2HW",
                                   "src_xyz_BS3HW", "is_bg_depth_inf", "hard_flow", "obj_mask"],
        R.plane_volume_rendering_flow: ["src_sigma_BS1HW", "src_flow_BS2HW", "src_xyz_BS3HW", "is_bg_depth_inf", "hard_flow"],
        R.weighted_sum_mpi: ["rgb_BS3HW", "xyz_BS3HW", "weights", "is_bg_depth_inf"],
        R.get_xyz_from_depth: ["meshgrid_homo", "depth", "K_inv"],
        R.disparity_consistency_src_to_tgt: ["meshgrid_homo", "K_src_inv", "disparity_src", "G_tgt_src", "K_tgt", "disparity_tgt"],
        R.get_src_xyz_from_plane_disparity: ["meshgrid_src_homo", "mpi_disparity_src", "K_src_inv"],
        R.get_tgt_xyz_from_plane_disparity: ["xyz_src_BS3HW", "G_tgt_src"],
        R.render_tgt_rgb_depth: ["H_sampler", "mpi_rgb_src", "mpi_sigma_src", "mpi_disparity_src", "xyz_tgt_BS3HW",
                                 "xyz_src_BS3HW", "G_tgt_src", "K_src_inv", "K_tgt", "mpi_flow_src", "use_alpha",
                                 "is_bg_depth_inf", "hard_flow", "obj_mask"],
        HomographySample.sample: ["self", "src_BCHW", "d_src_B", "G_tgt_src", "K_src_inv", "K_tgt"],
        HomographySample.sample_inverse: ["self", "src_BCHW", "d_src_B", "G_tgt_src", "K_src_inv", "K_tgt"],
        HomographySample.__init__: ["self", "H_tgt", "W_tgt", "device"],
        U.transform_G_xyz: ["G", "xyz", "is_return_homo"],
        U.gather_pixel_by_pxpy: ["img", "pxpy"],
        U.uniformly_sample_disparity_from_bins: ["batch_size", "disparity_np", "device"],
        U.uniformly_sample_disparity_from_linspace_bins: ["batch_size", "num_bins", "start", "end", "device"],
        U.sample_pdf: ["values", "weights", "N_samples"],
        geometry.transformation_from_parameters: ["axisangle", "translation", "invert"],
        geometry.get_translation_matrix: ["translation_vector"],
        geometry.rot_from_axisangle: ["vec"],
        geometry.Project3D.forward: ["self", "points", "K", "T", "T2"],
        geometry.BackprojectDepth.forward: ["self", "depth", "inv_K"],
    }
    for fn, names in want.items():
        assert list(inspect.signature(fn).parameters) == names, fn
    assert geometry.__all__ == ["BackprojectDepth", "Project3D", "transformation_from_parameters"]
    s = HomographySample(6, 9)
    assert s.meshgrid.shape == (3, 6, 9) and s.n.tolist() == [0, 0, 1] and (s.Height_tgt, s.Width_tgt) == (6, 9)
    assert s.meshgrid[0, 2, 5] == 5 and s.meshgrid[1, 2, 5] == 2 and s.meshgrid[2].min() == 1


def test_product_has_no_cpu_fallback():
    from utils.mpi import mpi_rendering as R
    z = torch.zeros(1, 2, 3, 4, 5)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        R.plane_volume_rendering(z, z[:, :, :1], z, None, None, None, False)
    import _native
    src = open(_native.__file__).read() + open(R.__file__).read()
    assert "oracle" not in src, "the product must never import the oracle"


def test_geometry_matches_reference_golden(golden_dir):
    import geometry
    g = np.load(os.path.join(golden_dir, "geometry.npz"))
    T = torch.from_numpy
    M = geometry.transformation_from_parameters(T(g["axisangle"]), T(g["translation"]))
    Mi = geometry.transformation_from_parameters(T(g["axisangle"]), T(g["translation"]), invert=True)
    np.testing.assert_array_equal(M.numpy(), g["M"])          # bit-exact: same op order
    np.testing.assert_array_equal(Mi.numpy(), g["M_invert"])
    h, w = g["depth"].shape[2:]
    cam = geometry.BackprojectDepth(1, h, w)(T(g["depth"]), T(g["K4inv"]))
    pix, z = geometry.Project3D(1, h, w)(cam, T(g["K4"]), T(g["M"][:1]))
    np.testing.assert_array_equal(cam.numpy(), g["cam_points"])
    np.testing.assert_array_equal(pix.numpy(), g["pix_coords"])
    np.testing.assert_array_equal(z.numpy(), g["z"])


def test_thin_torch_helpers():
    from utils.mpi import rendering_utils as U, mpi_rendering as R
    img = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).reshape(2, 3, 4, 5)
    pxpy = torch.tensor([[[0.4, 4.6, -3.0], [0.0, 2.5, 9.0]]]).repeat(2, 1, 1)
    got = U.gather_pixel_by_pxpy(img, pxpy)
    assert got.shape == (2, 3, 3)
    assert got[0, 0].tolist() == [img[0, 0, 0, 0].item(), img[0, 0, 2, 4].item(), img[0, 0, 3, 0].item()]
    d = U.uniformly_sample_disparity_from_linspace_bins(3, 8, 1.0, 0.01, "cpu")
    e = torch.linspace(1.0, 0.01, 9)
    assert d.shape == (3, 8) and (d <= e[:-1]).all() and (d >= e[1:]).all()
    d2 = U.uniformly_sample_disparity_from_bins(2, e.numpy(), "cpu")
    assert (d2 <= e[:-1]).all() and (d2 >= e[1:]).all()
    vals = torch.linspace(1, 2, 6).reshape(1, 1, 1, 6).repeat(2, 1, 3, 1)
    smp = U.sample_pdf(vals, torch.rand(2, 1, 3, 6) + 0.1, 7)
    assert smp.shape == (2, 1, 3, 7) and smp.min() >= 1 and smp.max() <= 2
    w = torch.rand(2, 4, 1, 3, 3)
    rgb, depth = R.weighted_sum_mpi(torch.rand(2, 4, 3, 3, 3), torch.rand(2, 4, 3, 3, 3), w, False)
    assert rgb.shape == (2, 3, 3, 3) and depth.shape == (2, 1, 3, 3)
    mesh = torch.stack(torch.meshgrid(torch.arange(3.), torch.arange(4.), indexing="xy") + (torch.ones(4, 3),), 0)
    xyz = R.get_xyz_from_depth(mesh, torch.full((1, 1, 4, 3), 2.0), torch.eye(3)[None])
    assert torch.equal(xyz[0, 2], torch.full((4, 3), 2.0)) and xyz[0, 0, 1, 2] == 4.0
