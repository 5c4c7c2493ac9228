"""Pins the CPU oracle (oracle/mpi_oracle.c) against outputs of the reference itself
(tests/golden/*.npz, made by tests/golden/make_golden.py from /root/reference on CPU).

Tolerances (fp32): image/flow/obj 1e-5 max-abs in CPU-numerics mode -- ten times tighter than the
1e-4 the north star allows the CUDA kernels; depth 1e-4 relative (it is a quotient of sums of
values up to ~60); the splat and validity masks bit-exact.
"""
import os

import numpy as np
import pytest

import synth
from oracle import mpi_oracle as orc

FRAMES = ["frame_white_random", "frame_smooth_regression", "frame_identity", "frame_object_pose_flags",
          "frame_wide_1280", "frame_c1_256"]


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def frame_inputs_of(g):
    B, S, H, W, seed = (int(v) for v in g["meta"])
    inp = synth.frame_inputs(B, S, H, W, str(g["kind"]), str(g["pose"]), seed)
    if "in_mpi" in g.files:  # small cases also store their inputs: the generator must reproduce them
        for k in inp:
            np.testing.assert_array_equal(inp[k], g["in_" + k], err_msg=k)
    return inp


@pytest.mark.parametrize("name", FRAMES)
def test_render_frame_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    inp = frame_inputs_of(g)
    rgb, depth, mask, flow, obj = orc.render_frame(
        inp["mpi"], inp["disparity"], inp["G"], inp["Kinv"], inp["K"], inp["obj_mask"],
        numerics=orc.NUM_CPU, is_bg_depth_inf=bool(g["is_bg_depth_inf"]), hard_flow=bool(g["hard_flow"]))
    assert np.abs(rgb - g["out_rgb"]).max() <= 1e-5
    assert np.abs(obj - g["out_obj"]).max() <= 1e-5
    assert np.abs(flow - g["out_flow"]).max() <= 1e-5 * max(1.0, np.abs(g["out_flow"]).max() / 16)
    np.testing.assert_array_equal(mask, g["out_mask"])
    np.testing.assert_allclose(depth, g["out_depth"], rtol=1e-4, atol=1e-4)


def test_cuda_numerics_mode_stays_within_north_star_tolerance(golden_dir):
    """The CUDA-library arithmetic variant differs from the CPU one by ulps of the coordinates."""
    g = load(golden_dir, "frame_smooth_regression")  # white-noise sigma amplifies 1-ulp coordinate
    inp = frame_inputs_of(g)                          # differences ~1000x (last plane): DESIGN.md
    rgb, depth, mask, flow, obj = orc.render_frame(
        inp["mpi"], inp["disparity"], inp["G"], inp["Kinv"], inp["K"], inp["obj_mask"], numerics=orc.NUM_CUDA)
    assert np.abs(rgb - g["out_rgb"]).max() <= 1e-4
    assert np.abs(flow - g["out_flow"]).max() <= 1e-4


def test_functions_match_reference(golden_dir):
    g = load(golden_dir, "functions_tiny")
    B, S, H, W, _ = (int(v) for v in g["meta"])
    mpi, disp, G, Kinv, K = (g["in_" + k] for k in ("mpi", "disparity", "G", "Kinv", "K"))
    xyz_src = orc.get_src_xyz_from_plane_disparity(disp, Kinv, H, W)
    np.testing.assert_array_equal(xyz_src, g["xyz_src"])          # a2: bit-exact
    xyz_tgt = orc.get_tgt_xyz_from_plane_disparity(xyz_src, G)
    np.testing.assert_allclose(xyz_tgt, g["xyz_tgt"], rtol=0, atol=2e-5)  # a3 (values up to ~60)
    rep = lambda m: np.repeat(m[:, None], S, 1).reshape(B * S, *m.shape[1:])
    depth = (1.0 / disp.astype(np.float32)).astype(np.float32).reshape(-1)
    Hts, Hst = orc.plane_homographies(rep(G), rep(Kinv), rep(K), depth, orc.NUM_CPU)
    src = np.concatenate([mpi, g["xyz_tgt"]], 2).reshape(B * S, 7, H, W)
    warped, valid, flowB2A = orc.homography_sample(src, Hst, orc.NUM_CPU)
    np.testing.assert_array_equal(valid, g["valid"])                # a4 validity: bit-exact
    assert np.abs(flowB2A - g["flowB2A"]).max() <= 2e-5
    assert np.abs(warped[:, :4] - g["warped"][:, :4]).max() <= 1e-5   # rgb + sigma
    np.testing.assert_allclose(warped[:, 4:], g["warped"][:, 4:], rtol=1e-5, atol=1e-4)  # xyz
    flowA2B = orc.homography_sample_inverse(Hts, H, W)
    assert np.abs(flowA2B - g["flowA2B"]).max() <= 2e-5              # a5
    # a7 on the reference's own warped tensors, so only the compositing is compared
    w5 = g["warped"].reshape(B, S, 7, H, W)
    fl = np.ascontiguousarray(g["flowA2B"].transpose(0, 3, 1, 2)).reshape(B, S, 2, H, W)
    objm = np.ascontiguousarray(np.broadcast_to(g["in_obj_mask"][:, None], (B, S, 1, H, W)))
    rgb, dep, tacc, wts, flow, obj = orc.plane_volume_rendering(
        w5[:, :, 0:3], w5[:, :, 3:4], w5[:, :, 4:7], fl, objm)
    for a, k, tol in ((rgb, "pvr_rgb", 2e-6), (tacc, "pvr_tacc", 2e-6), (wts, "pvr_weights", 2e-6),
                      (obj, "pvr_obj", 2e-6), (flow, "pvr_flow", 2e-5)):
        assert np.abs(a - g[k]).max() <= tol, k
    np.testing.assert_allclose(dep, g["pvr_depth"], rtol=2e-5, atol=2e-5)
    dep_bg = orc.plane_volume_rendering(w5[:, :, 0:3], w5[:, :, 3:4], w5[:, :, 4:7], is_bg_depth_inf=True)[1]
    np.testing.assert_allclose(dep_bg, g["pvr_bg_depth"], rtol=2e-5, atol=1e-3)
    # a8
    soft = orc.plane_volume_rendering_flow(mpi[:, :, 3:4], fl, g["xyz_src"], hard_flow=False)
    hard = orc.plane_volume_rendering_flow(mpi[:, :, 3:4], fl, g["xyz_src"], hard_flow=True)
    assert np.abs(soft - g["flow_soft"]).max() <= 2e-5
    assert np.abs(hard - g["flow_hard"]).max() <= 2e-5
    # a9 alpha mode
    alpha = np.clip(mpi[:, :, 3:4], 0, 1)
    comp, w = orc.alpha_composition(alpha, mpi[:, :, 0:3])
    assert np.abs(comp - g["ac_rgb"]).max() <= 2e-6
    assert np.abs(w - g["ac_w"]).max() <= 2e-6


def test_geometry_matches_reference(golden_dir):
    g = load(golden_dir, "geometry")
    M = orc.transformation_from_parameters(g["axisangle"], g["translation"])
    Mi = orc.transformation_from_parameters(g["axisangle"], g["translation"], invert=True)
    assert np.abs(M - g["M"]).max() <= 1e-6
    assert np.abs(Mi - g["M_invert"]).max() <= 1e-6


@pytest.mark.parametrize("tag", ["dense", "pile", "identity", "row", "kitti_strip"])
def test_splat_restatement_is_bit_exact(golden_dir, tag):
    g = load(golden_dir, "splat")
    h, w, seed = (int(v) for v in g[tag + "_meta"])
    src, idx, idy, z = synth.splat_inputs(h, w, seed, str(g[tag + "_mode"]))
    np.testing.assert_array_equal(orc.forward_warping(src, idx, idy, z), g[tag + "_warped"])
    if orc.ref_splat_lib() is not None:  # the reference's own C, compiled from source
        np.testing.assert_array_equal(orc.forward_warping(src, idx, idy, z, use_ref=True), g[tag + "_warped"])


def test_known_answers():
    """Self-made known-answer cases: identity pose => flow == 0 and image ==
    plain composite; pure x-translation => per-plane flow = fx*tx/depth."""
    inp = synth.frame_inputs(1, 4, 24, 32, "smooth", "identity", 5)
    rgb, depth, mask, flow, obj = orc.render_frame(inp["mpi"], inp["disparity"], inp["G"], inp["Kinv"],
                                                   inp["K"], inp["obj_mask"])
    assert np.abs(flow).max() <= 1e-5
    assert np.all(mask == 4)
    H, W = 24, 32
    xyz = orc.get_src_xyz_from_plane_disparity(inp["disparity"], inp["Kinv"], H, W)
    ref_rgb = orc.plane_volume_rendering(inp["mpi"][:, :, :3], inp["mpi"][:, :, 3:4], xyz)[0]
    assert np.abs(rgb - ref_rgb).max() <= 2e-4  # smooth MPI: see DESIGN.md on the last-plane 1e3 gain
    G = np.eye(4, dtype=np.float32)[None].copy()
    G[0, 0, 3] = 0.1
    depth_s = 1.0 / inp["disparity"].reshape(-1)
    rep = lambda m: np.repeat(m, 4, 0)
    Hts, _ = orc.plane_homographies(rep(G), rep(inp["Kinv"]), rep(inp["K"]), depth_s)
    fl = orc.homography_sample_inverse(Hts, H, W)
    fx = inp["K"][0, 0, 0]
    for s in range(4):
        assert np.abs(fl[s, :, :, 0] - fx * 0.1 / depth_s[s]).max() <= 1e-4
        assert np.abs(fl[s, :, :, 1]).max() <= 1e-5


@pytest.mark.parametrize("name", ["frame_white_random", "frame_smooth_regression", "frame_object_pose_flags"])
def test_torch_ref_matches_reference(golden_dir, name):
    """oracle/torch_ref.py (the library-operator restatement used on the GPU box) against the
    reference's own CPU outputs."""
    import torch
    from oracle import torch_ref
    g = load(golden_dir, name)
    inp = frame_inputs_of(g)
    T = torch.from_numpy
    out = torch_ref.render_frame(T(inp["mpi"]), T(inp["disparity"]), T(inp["G"]), T(inp["Kinv"]), T(inp["K"]),
                                 T(inp["obj_mask"]), is_bg_depth_inf=bool(g["is_bg_depth_inf"]),
                                 hard_flow=bool(g["hard_flow"]))
    rgb, depth, mask, flow, obj = (o.numpy() for o in out)
    assert np.abs(rgb - g["out_rgb"]).max() <= 2e-6
    assert np.abs(flow - g["out_flow"]).max() <= 2e-6
    assert np.abs(obj - g["out_obj"]).max() <= 2e-6
    np.testing.assert_array_equal(mask, g["out_mask"])
    np.testing.assert_allclose(depth, g["out_depth"], rtol=1e-5, atol=1e-5)
