"""Sample assembly around the renderer (reference utils/utils.py:159-288).

CPU tests pin oracle/sample_oracle.py and the host logic (pose draws, signatures) to tests/golden/sample_pair.npz,
which holds outputs of the UNMODIFIED reference (tests/golden/make_golden.py --sample-only).  GPU tests compare the
CUDA kernels with the same fixture: the merge bit-exact, the blend and the renders within the 1e-4 fp32 contract.
"""
import inspect
import os
import random
import types

import numpy as np
import pytest
import torch

from oracle import sample_oracle as so

T = torch.from_numpy


@pytest.fixture(scope="module")
def gold(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "sample_pair.npz")))


def _kinv(g):
    return torch.inverse(T(g["in_K"]).double()).float()


# ---------------------------------------------------------------------------------------------------- CPU
def test_oracle_blend_matches_reference(gold):
    mpi = T(gold["in_mpi"])
    out = so.blend_src(T(gold["in_src"]), mpi[:, :, :3], mpi[:, :, 3:], T(gold["in_disparity"]), _kinv(gold))
    np.testing.assert_allclose(out.numpy(), gold["blended"], rtol=0, atol=1e-6)


def test_oracle_assembly_is_bit_exact(gold):
    flow_mix, frame_mix, fill = so.assemble_sample(gold["frame"][0], gold["frame_dync"][0], gold["mask"],
                                                   gold["mask_dync"], gold["flow"][0], gold["flow_dync"][0],
                                                   gold["in_obj"], float(gold["mask_thresh"]))
    assert np.array_equal(flow_mix, gold["flow_mix"])
    assert np.array_equal(frame_mix, gold["frame_mix"])
    assert np.array_equal(fill, gold["fill_mask"])
    assert 0 < fill.sum() < fill.size and np.array_equal(so.to_u8_bgr(gold["in_src"][0]), gold["src_np"])


def test_pose_draws_follow_the_reference(gold):
    """same `random` stream, same order (object-complement pose first, :206-208) -> the poses of the fixture"""
    from utils import utils as U
    random.seed(int(gold["meta"][4]))
    cam_dyn = U.generate_random_pose(float(gold["ext_cz"]), device="cpu")
    cam_obj = U.generate_random_pose(float(gold["ext_cz"]), base_motions=[0, 0, 0], device="cpu")
    np.testing.assert_allclose(cam_dyn.numpy(), gold["G_cam"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(cam_obj.numpy(), gold["G_obj"], rtol=0, atol=1e-7)
    assert cam_dyn[2, 3] < 0 and abs(float(cam_obj[0, 3])) <= 0.05 + 1e-6     # camera moves forward; object motion halved
    path = U.gen_swing_path()
    assert len(path) == 90 and path[0].shape == (4, 4) and torch.allclose(path[0], torch.eye(4), atol=1e-7)
    assert abs(float(path[22][0, 3]) - 0.14 * np.sin(2 * np.pi * 22 / 89)) < 1e-6


def test_signatures_mirror_reference():
    from utils import utils as U
    want = {
        U.gen_swing_path: ["num_frames", "r_x", "r_y", "r_z"],
        U.render_novel_view: ["mpi_all_rgb_src", "mpi_all_sigma_src", "disparity_all_src", "G_tgt_src", "K_src_inv",
                              "K_tgt", "K_src", "src_pose", "homography_sampler", "hard_flow"],
        U.render_novel_view_dynamic: ["obj_mask", "mpi_all_rgb_src", "mpi_all_sigma_src", "disparity_all_src",
                                      "G_tgt_src", "K_src_inv", "K_tgt", "K_src", "src_pose", "homography_sampler",
                                      "hard_flow"],
        U.render_3dphoto_dynamic: ["opt", "src_imgs", "obj_mask", "disp", "mpi_all_src", "disparity_all_src", "k_src",
                                   "k_tgt", "data_path", "name", "hard_flow", "mask_thresh"],
    }
    for fn, names in want.items():
        assert list(inspect.signature(fn).parameters) == names, fn
    assert list(inspect.signature(U.generate_random_pose).parameters)[:2] == ["ext_cz", "base_motions"]
    with pytest.raises(RuntimeError, match="CUDA-only"):      # no CPU fallback
        z = torch.zeros(1, 2, 4, 4, 8)
        U.blend_src_into_mpi(torch.zeros(1, 3, 4, 8), z[:, :, :3], z[:, :, 3:], torch.ones(1, 2), torch.eye(3)[None])
    assert "oracle" not in open(U.__file__).read().replace("oracle/", "")


# ---------------------------------------------------------------------------------------------------- GPU
def _dev(gold, *names):
    return [T(gold[n]).cuda() for n in names]


@pytest.mark.gpu
def test_blend_kernel_matches_reference(gold):
    from utils import utils as U
    mpi, src, disp = _dev(gold, "in_mpi", "in_src", "in_disparity")
    out = U.blend_src_into_mpi(src, mpi[:, :, :3], mpi[:, :, 3:], disp, _kinv(gold).cuda())
    np.testing.assert_allclose(out.cpu().numpy(), gold["blended"], rtol=0, atol=2e-6)
    # strided views of a larger buffer and a batch of two
    big = torch.zeros(2, 6, 5, 40, 64, device="cuda")
    big[:, :, 1:5] = mpi
    out2 = U.blend_src_into_mpi(src.expand(2, 3, 40, 64), big[:, :, 1:4], big[:, :, 4:5], disp.expand(2, 6),
                                _kinv(gold).cuda().expand(2, 3, 3))
    assert torch.equal(out2[0], out[0]) and torch.equal(out2[1], out[0])


@pytest.mark.gpu
def test_assemble_kernel_is_bit_exact(gold):
    import _native as nat
    frame, frame_d, mask, mask_d, flow, flow_d, obj = _dev(gold, "frame", "frame_dync", "mask", "mask_dync", "flow",
                                                           "flow_dync", "in_obj")
    H, W = frame.shape[-2:]
    flow_mix = torch.empty(H, W, 2, device="cuda")
    frame_mix = torch.empty(H, W, 3, dtype=torch.uint8, device="cuda")
    fill = torch.empty(H, W, dtype=torch.uint8, device="cuda")
    nat.check(nat.lib().mpiflow_assemble_sample_f32(
        nat.ptr(frame), nat.ptr(frame_d), nat.ptr(mask), nat.ptr(mask_d), nat.ptr(flow), nat.ptr(flow_d), nat.ptr(obj),
        float(gold["mask_thresh"]), H, W, nat.ptr(flow_mix), nat.ptr(frame_mix), nat.ptr(fill), nat.stream(frame)))
    assert np.array_equal(flow_mix.cpu().numpy(), gold["flow_mix"])
    assert np.array_equal(frame_mix.cpu().numpy(), gold["frame_mix"])
    assert np.array_equal(fill.cpu().numpy(), gold["fill_mask"])
    # un-clipped flows are clipped inside the kernel (:347); half-way values round to even like np.round
    big, big_d = flow * 1e3, flow_d * 1e3             # named: the pointers must outlive the call
    nat.check(nat.lib().mpiflow_assemble_sample_f32(
        nat.ptr(frame), nat.ptr(frame_d), nat.ptr(mask), nat.ptr(mask_d), nat.ptr(big), nat.ptr(big_d),
        nat.ptr(obj), float(gold["mask_thresh"]), H, W, nat.ptr(flow_mix), nat.ptr(frame_mix), nat.ptr(fill),
        nat.stream(frame)))
    assert np.array_equal(flow_mix.cpu().numpy(), np.clip(gold["flow_mix"] * np.float32(1e3), -200, 200))
    half = torch.tensor([0.5, 1.5, 2.5, 254.5, 300.0, -3.0], device="cuda") / 255
    f3 = half.reshape(1, 1, 6).expand(3, 1, 6).contiguous()
    one, zero2 = torch.ones(1, 6, device="cuda"), torch.zeros(2, 1, 6, device="cuda")
    o_flow = torch.empty(1, 6, 2, device="cuda")
    o_u8 = torch.empty(1, 6, 3, dtype=torch.uint8, device="cuda")
    o_fill = torch.empty(1, 6, dtype=torch.uint8, device="cuda")
    nat.check(nat.lib().mpiflow_assemble_sample_f32(
        nat.ptr(f3), nat.ptr(f3), nat.ptr(one), nat.ptr(one), nat.ptr(zero2), nat.ptr(zero2), nat.ptr(one), 0.99, 1, 6,
        nat.ptr(o_flow), nat.ptr(o_u8), nat.ptr(o_fill), nat.stream(f3)))
    want = so.to_u8_bgr(f3.cpu().numpy())
    assert np.array_equal(o_u8.cpu().numpy(), want) and want[0, :, 0].tolist() == [0, 2, 2, 254, 255, 0]


@pytest.mark.gpu
def test_shared_mpi_batch_equals_materialised_batch():
    """one MPI under several poses (batch stride 0) == the same launch on a repeated copy, bit for bit"""
    import synth
    from utils.mpi import mpi_rendering as R
    for (S, H, W, n) in ((6, 40, 64, 2), (5, 70, 200, 3), (64, 384, 1280, 2)):
        inp = synth.frame_inputs(n, S, H, W, "white", "random", 114531)
        mpi = T(inp["mpi"][:1]).cuda()
        G, Kinv, K, disp = (T(inp[k]).cuda() for k in ("G", "Kinv", "K", "disparity"))
        obj = T(inp["obj_mask"]).cuda()
        args = (disp, G, Kinv, K, obj.reshape(n, 1, 1, H, W))
        rep = mpi.repeat(n, 1, 1, 1, 1)
        a = R.render_fused(mpi[:, :, :3].expand(n, S, 3, H, W), mpi[:, :, 3:].expand(n, S, 1, H, W), *args)
        b = R.render_fused(rep[:, :, :3], rep[:, :, 3:], *args)
        for x, y in zip(a, b):
            assert torch.equal(x, y), (S, H, W)


@pytest.mark.gpu
def test_sample_pair_matches_reference(gold):
    from utils import utils as U
    mpi, src, obj, disp, K, G_obj, G_cam = _dev(gold, "in_mpi", "in_src", "in_obj", "in_disparity", "in_K", "G_obj",
                                                "G_cam")
    flow_mix, frame_mix, fill, (rgb2, obj2, flow2) = U.render_sample_pair(src, obj, mpi, disp, K, K, G_obj, G_cam)
    np.testing.assert_allclose(rgb2[0].cpu().numpy(), gold["frame"][0], rtol=0, atol=1e-4)
    np.testing.assert_allclose(rgb2[1].cpu().numpy(), gold["frame_dync"][0], rtol=0, atol=1e-4)
    np.testing.assert_allclose(obj2[0].cpu().numpy(), gold["mask"][0], rtol=0, atol=1e-4)
    np.testing.assert_allclose(obj2[1].cpu().numpy(), gold["mask_dync"][0], rtol=0, atol=1e-4)
    np.testing.assert_allclose(torch.clip(flow2[0], -200, 200).cpu().numpy(), gold["flow"][0], rtol=0, atol=1e-4)
    np.testing.assert_allclose(flow_mix.cpu().numpy(), gold["flow_mix"], rtol=0, atol=1e-4)
    # thresholds and u8 rounding sit on fp32 values that agree to 1e-4: a few pixels may fall on the other side
    fill_np, frame_np = fill.cpu().numpy(), frame_mix.cpu().numpy()
    assert (fill_np != gold["fill_mask"]).mean() < 0.005
    same_region = (fill_np == gold["fill_mask"])[..., None] & (np.abs(gold["mask"][0, 0] - 0.99) > 1e-3)[..., None]
    diff = np.abs(frame_np.astype(np.int16) - gold["frame_mix"].astype(np.int16))
    assert (diff[np.broadcast_to(same_region, diff.shape)] > 1).mean() < 0.005
    assert (diff == 0).mean() > 0.97
    # the two-launch route of the drop-in functions gives the same frames
    kinv = _kinv(gold).cuda()
    blended = U.blend_src_into_mpi(src, mpi[:, :, :3], mpi[:, :, 3:], disp, kinv)
    one = U.render_novel_view_dynamic(obj, blended, mpi[:, :, 3:], disp, G_obj, kinv, K, K, None, None)
    np.testing.assert_allclose(one[0].cpu().numpy(), rgb2[:1].cpu().numpy(), rtol=0, atol=5e-6)
    np.testing.assert_allclose(one[3].cpu().numpy(), obj2[:1].cpu().numpy(), rtol=0, atol=5e-6)
    plain = U.render_novel_view(blended, mpi[:, :, 3:], disp, G_obj, kinv, K, K, None, None)
    np.testing.assert_allclose(plain[0].cpu().numpy(), one[0].cpu().numpy(), rtol=0, atol=5e-6)
    np.testing.assert_allclose(plain[2].clip(-200, 200).cpu().numpy(), one[2].cpu().numpy(), rtol=0, atol=5e-6)


@pytest.mark.gpu
def test_render_3dphoto_dynamic_end_to_end(gold):
    pytest.importorskip("cv2")
    from utils import utils as U
    mpi, src, obj, disp, K = _dev(gold, "in_mpi", "in_src", "in_obj", "in_disparity", "in_K")
    random.seed(int(gold["meta"][4]))
    opt = types.SimpleNamespace(ext_cz=float(gold["ext_cz"]))
    flow_mix, src_np, inpainted, res = U.render_3dphoto_dynamic(opt, src, obj, None, mpi, disp, K, K, name="sample.png")
    assert res is None and np.array_equal(src_np, gold["src_np"])
    assert flow_mix.dtype == np.float32 and flow_mix.shape == gold["flow_mix"].shape
    assert inpainted.dtype == np.uint8 and inpainted.shape == gold["inpainted"].shape
    # poses are computed on the device here (sin/cos of the axis-angle), on the host in the fixture
    np.testing.assert_allclose(flow_mix, gold["flow_mix"], rtol=0, atol=1e-3)
    keep = gold["fill_mask"] == 0
    assert (np.abs(inpainted.astype(np.int16) - gold["inpainted"].astype(np.int16))[keep] <= 1).mean() > 0.99


@pytest.mark.gpu
def test_render_3dphoto_dynamic_device_and_host_hole_filling_agree(gold, monkeypatch):
    """MPIFLOW_INPAINT=device keeps the hole filling on the GPU (csrc/inpaint.cu): the same bytes as the host call"""
    pytest.importorskip("cv2")
    from utils import utils as U
    mpi, src, obj, disp, K = _dev(gold, "in_mpi", "in_src", "in_obj", "in_disparity", "in_K")
    opt = types.SimpleNamespace(ext_cz=float(gold["ext_cz"]))
    outs = []
    for backend in ("host", "device"):
        monkeypatch.setenv("MPIFLOW_INPAINT", backend)
        random.seed(int(gold["meta"][4]))
        outs.append(U.render_3dphoto_dynamic(opt, src, obj, None, mpi, disp, K, K, name="sample.png"))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][2], outs[1][2])


@pytest.mark.gpu
def test_poses_per_mpi_equals_materialised_batch():
    """MPIFLOW_POSES_PER_MPI(P): frames b*P..b*P+P-1 read MPI b == the same launch on repeat_interleave'd MPIs"""
    import _native as nat
    import synth
    from utils.mpi import mpi_rendering as R
    for (S, H, W, N, P) in ((6, 40, 64, 2, 2), (5, 70, 200, 2, 3), (64, 384, 1280, 2, 2), (7, 48, 128, 4, 2)):
        inp = synth.frame_inputs(N * P, S, H, W, "white", "random", 114532)
        mpi = T(inp["mpi"][:N]).cuda()
        G, Kinv, K, disp = (T(inp[k]).cuda() for k in ("G", "Kinv", "K", "disparity"))
        obj = T(inp["obj_mask"]).cuda().reshape(N * P, 1, 1, H, W)
        a = R.render_fused(mpi[:, :, :3], mpi[:, :, 3:], disp[:N], G, Kinv[:N], K[:N], obj, poses_per_mpi=P)
        rep = mpi.repeat_interleave(P, 0)
        b = R.render_fused(rep[:, :, :3], rep[:, :, 3:], disp, G, Kinv, K, obj)
        for x, y in zip(a, b):
            assert x.shape[0] == N * P and torch.equal(x, y), (S, H, W, N, P)
    with pytest.raises(nat.MpiflowError):        # 3 frames cannot be pairs: MPIFLOW_E_BADARG before any launch
        _OddFrames.outs(3, 40, 64)


class _OddFrames:
    """render_fused with a frame count that is not a multiple of the poses: drive the C entry directly"""

    @staticmethod
    def outs(B, H, W):
        import _native as nat
        new = lambda c: torch.empty((B, c, H, W), device="cuda")
        o = (new(3), new(1), new(1), new(2), None)
        ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
        z = torch.zeros(64, 16, device="cuda")
        nat.check(nat.lib().mpiflow_render_fused_f32(
            nat.ptr(z), 1, 1, 1, nat.ptr(z), 1, 1, None, 0, 0, nat.ptr(z), nat.ptr(z), nat.ptr(z), nat.ptr(z),
            B, 2, H, W, (2 - 1) << 8, nat.ptr(o[0]), nat.ptr(o[1]), nat.ptr(o[2]), nat.ptr(o[3]), None, nat.ptr(ws),
            1 << 20, None, None))
        return o


@pytest.mark.gpu
def test_batched_sample_pairs_equal_single_samples(gold):
    from utils import utils as U
    mpi, src, obj, disp, K, G_obj, G_cam = _dev(gold, "in_mpi", "in_src", "in_obj", "in_disparity", "in_K", "G_obj",
                                                "G_cam")
    kinv = _kinv(gold).cuda()
    N = 3
    mpis = torch.cat([mpi * (1 - 0.1 * n) for n in range(N)], 0)                  # three different MPIs
    poses = torch.stack([torch.stack((G_obj, G_cam)), torch.stack((G_cam, G_obj)), torch.stack((G_obj, G_obj))], 0)
    rep = lambda t: t.expand(N, *t.shape[1:]).contiguous()
    flow_b, frame_b, fill_b = U.render_sample_pairs(rep(src), rep(obj.reshape(1, 1, *obj.shape[-2:])), mpis, rep(disp),
                                                    rep(kinv), rep(K), poses)
    for n in range(N):
        flow_1, frame_1, fill_1, _ = U.render_sample_pair(src, obj, mpis[n:n + 1], disp, K, K, poses[n, 0], poses[n, 1],
                                                          k_src_inv=kinv)
        np.testing.assert_allclose(flow_b[n].cpu().numpy(), flow_1.cpu().numpy(), rtol=0, atol=5e-6)
        assert (fill_b[n] != fill_1).float().mean().item() < 0.002
        assert (frame_b[n] != frame_1).float().mean().item() < 0.005
    # sample 0 is the fixture's sample
    np.testing.assert_allclose(flow_b[0].cpu().numpy(), gold["flow_mix"], rtol=0, atol=1e-4)
