"""moving_obj.py drop-in (reference moving_obj.py:16-175) against
tests/golden/moving_obj.npz: what the UNMODIFIED reference passed to, and got back from, its C splat
(tests/golden/make_golden.py --moving-obj-only)."""
import inspect
import os
import random

import numpy as np
import pytest
import torch

from oracle import mpi_oracle as orc

T = torch.from_numpy


@pytest.fixture(scope="module")
def gold(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "moving_obj.npz")))


def _maps(g, device):
    import moving_obj as M
    random.seed(int(g["meta"][2]))
    motion = M.draw_object_motion()
    after = random.random()
    out = M.splat_index_maps(T(g["in_disp"])[None, None], T(g["in_K"]), T(g["in_Kinv"]), T(g["in_inst"])[None, None],
                             motion, device)
    return motion, after, out


def test_index_maps_are_bit_identical_on_the_same_device_class(gold):
    """a13: same torch operators, same order -> the reference's int64 maps and fp32 depth, exactly (CPU vs CPU)"""
    motion, after, (sx, sy, z, p1) = _maps(gold, "cpu")
    assert after == float(gold["next_random"])                  # the `random` stream is consumed like the reference
    assert 0.05 <= motion[0] <= 0.1 and -0.1 <= motion[1] <= -0.05 and 0.05 <= motion[2] <= 0.1
    assert sx.dtype == torch.int64 and sx.shape == (1,) + gold["safe_x"].shape
    np.testing.assert_array_equal(sx[0].numpy(), gold["safe_x"])
    np.testing.assert_array_equal(sy[0].numpy(), gold["safe_y"])
    np.testing.assert_array_equal(z[0].numpy(), gold["z"])
    assert np.array_equal(gold["img"], gold["in_rgb"])


def test_restated_splat_reproduces_the_reference_call(gold):
    out = orc.forward_warping(gold["img"], gold["safe_x"], gold["safe_y"], gold["z"])
    np.testing.assert_array_equal(out, gold["warped"])


def test_signature_and_no_cpu_fallback(gold):
    import moving_obj as M
    assert list(inspect.signature(M.moveing_object_with_mask).parameters)[:7] == [
        "depth_path", "disp", "rgb", "K", "inv_K", "instance_mask", "i"]
    h, w = gold["z"].shape
    with pytest.raises(RuntimeError, match="CUDA-only"):
        M.forward_warp(T(gold["img"]), T(gold["safe_x"]), T(gold["safe_y"]), T(gold["z"]))
    assert "oracle" not in open(M.__file__).read()


@pytest.mark.gpu
def test_device_splat_of_reference_maps_is_bit_exact(gold):
    import moving_obj as M
    warped = M.forward_warp(T(gold["img"]).cuda(), T(gold["safe_x"]).cuda(), T(gold["safe_y"]).cuda(), T(gold["z"]).cuda())
    np.testing.assert_array_equal(warped.cpu().numpy(), gold["warped"])


@pytest.mark.gpu
def test_moving_object_on_the_device(gold):
    pytest.importorskip("cv2")
    import moving_obj as M
    h, w = gold["z"].shape
    random.seed(int(gold["meta"][2]))
    out = M.moveing_object_with_mask(None, T(gold["in_disp"])[None, None], gold["in_rgb"], T(gold["in_K"]),
                                     T(gold["in_Kinv"]), T(gold["in_inst"])[None, None], 0)
    assert random.random() == float(gold["next_random"])
    sx, sy, z = (out[k][0].cpu().numpy() for k in ("safe_x", "safe_y", "z"))
    # static pixels project onto whole coordinates +- 1 ulp (identity camera), so the truncation flips on many of them
    # between cuBLAS and the host BLAS -- the reference's own CPU and CUDA runs differ the same way; never by more than 1
    assert np.abs(sx - gold["safe_x"]).max() <= 1 and np.abs(sy - gold["safe_y"]).max() <= 1
    assert (sx != gold["safe_x"]).mean() < 0.3 and (sy != gold["safe_y"]).mean() < 0.3
    np.testing.assert_allclose(z, gold["z"], rtol=1e-5, atol=1e-5)
    # whatever the maps are, the device splat equals the serial C loop on the same maps
    want = orc.forward_warping(gold["in_rgb"], sx, sy, z)
    np.testing.assert_array_equal(out["im1_raw"], want[..., :3])
    np.testing.assert_array_equal(out["masks"]["H"], want[..., 3:4])
    assert out["im1"].shape == (h, w, 3) and out["im1"].dtype == np.uint8
    assert out["flow_01"].shape == (1, h, w, 2)
    keep = out["masks"]["H"][..., 0] == 1
    assert np.array_equal(out["im1"][keep], out["im1_raw"][keep])          # inpainting only touches the holes
    assert set(np.unique(out["masks"]["H'"])) <= {0, 1}
