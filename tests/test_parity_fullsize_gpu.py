"""Parity of the launches bench.py times, at BASELINE.json's full sizes, against the UNMODIFIED reference on CUDA.

oracle/ref_modules.py imports the reference's own utils/mpi/*.py (byte copies under oracle/_ref/refsrc/, made by
oracle/Makefile) and runs get_src_xyz_from_plane_disparity -> get_tgt_xyz_from_plane_disparity ->
render_tgt_rgb_depth on the same GPU: cuBLAS matmul, ATen grid_sample / norm / exp / cumprod / sum -- "the
reference PyTorch/grid_sample path" of north_star itself.  oracle/torch_ref.py (a restatement of the same operator
sequence) is cross-checked against it here as well.  Gate: 1e-4 max-abs on image, flow and object mask (fp32), the
validity count exact.

Shapes: the exact bench.py launch (16 frames of 1280x384x64, white noise, random poses, plane-spacing table on);
BASELINE config 4 (640x480x64, batched); config 5 (1920x1080 at 16 / 64 / 128 planes); the shared-MPI pose pairs of
the sample assembly; poses that overflow the TMA box under MPIFLOW_POSES_PER_MPI.
"""
import numpy as np
import pytest
import torch

import synth
from oracle import ref_modules, torch_ref

pytestmark = pytest.mark.gpu
TOL = 1e-4
NAMES = ["rgb", "depth", "mask", "flow", "obj"]


def device_batch(B, S, H, W, seed, pose="random"):
    """white-noise frames generated on the device (SURVEY 8d statistics), like bench.make_device_batch"""
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(seed)
    mpi = torch.empty((B, S, 4, H, W), dtype=torch.float32, device=dev)
    mpi[:, :, 0:3].copy_(torch.rand((B, S, 3, H, W), device=dev, generator=gen))
    mpi[:, :, 3:4].copy_(torch.relu(torch.randn((B, S, 1, H, W), device=dev, generator=gen)) + 1e-4)
    K = synth.intrinsics(H, W, B)
    return dict(mpi=mpi, disparity=torch.from_numpy(synth.plane_disparities(S, B)).to(dev),
                G=torch.from_numpy(synth.poses(B, pose, seed)).to(dev), K=torch.from_numpy(K).to(dev),
                Kinv=torch.from_numpy(synth.intrinsics_inv(K)).to(dev),
                obj_mask=torch.from_numpy(synth.obj_mask(B, H, W, seed)).to(dev))


def fused(d, **kw):
    from utils.mpi import mpi_rendering as R
    mpi = d["mpi"]
    out = R.render_fused(mpi[:, :, 0:3], mpi[:, :, 3:4], d["disparity"], d["G"], d["Kinv"], d["K"], d["obj_mask"], **kw)
    torch.cuda.synchronize()
    return out


def reference_frame(d, b):
    """the unmodified reference on frame b of the batch (one frame at a time: its temporaries are ~10x the MPI)"""
    one = {k: v[b:b + 1] for k, v in d.items()}
    with torch.no_grad():
        out = ref_modules.render_frame(one["mpi"], one["disparity"], one["G"], one["Kinv"], one["K"], one["obj_mask"])
    torch.cuda.synchronize()
    return out


def errors(got, want, b=None):
    sel = (lambda t: t) if b is None else (lambda t: t[b:b + 1])
    return {n: float((sel(g) - w).abs().max()) for n, g, w in zip(NAMES, got, want)}


def check(err, what):
    assert err["rgb"] <= TOL and err["flow"] <= TOL and err["obj"] <= TOL and err["mask"] == 0, (what, err)


needs_ref = pytest.mark.skipif(not ref_modules.available(), reason="oracle/_ref/refsrc is built by `make -C oracle ref`")


@needs_ref
def test_torch_ref_restatement_equals_the_unmodified_reference_on_cuda():
    """the port used by the other GPU tests is the reference's operator sequence bit for bit, on this device too"""
    for case in ((1, 16, 96, 320, "white", "random", 4), (2, 9, 64, 128, "smooth", "object", 5)):
        d = {k: torch.from_numpy(v).cuda() for k, v in synth.frame_inputs(*case).items()}
        a = ref_modules.render_frame(d["mpi"], d["disparity"], d["G"], d["Kinv"], d["K"], d["obj_mask"])
        b = torch_ref.render_frame(d["mpi"], d["disparity"], d["G"], d["Kinv"], d["K"], d["obj_mask"])
        for n, x, y in zip(NAMES, a, b):
            assert float((x - y).abs().max()) <= 2e-6, (case, n)
        print(case, "bit-equal" if all(torch.equal(x, y) for x, y in zip(a, b)) else "equal to rounding")


@needs_ref
def test_bench_launch_shape_against_the_unmodified_reference():
    """bench.py's launch: 16 frames of config 2 in one persistent launch, plane-spacing table on."""
    B, S, H, W = 16, 64, 384, 1280
    d = device_batch(B, S, H, W, synth.SEED)
    got = fused(d)
    worst = {n: 0.0 for n in NAMES}
    big_flow = 0.0
    for b in range(B):
        want = reference_frame(d, b)
        err = errors(got, want, b)
        check(err, f"frame {b}")
        worst = {n: max(worst[n], err[n]) for n in NAMES}
        big_flow = max(big_flow, float(want[3].abs().max()))
        del want
    assert big_flow > 50, big_flow
    print("bench-shape parity, worst over 16 frames:", worst)


@needs_ref
@pytest.mark.parametrize("kind", ["white", "smooth"])
def test_config2_single_frame_against_the_unmodified_reference(kind):
    """the reference's own call shape: B = 1 (64x8 tiles, no table)"""
    d = {k: torch.from_numpy(v).cuda() for k, v in synth.frame_inputs(1, 64, 384, 1280, kind, "random", 77).items()}
    check(errors(fused(d), reference_frame(d, 0)), kind)


@needs_ref
def test_config4_batch_against_the_unmodified_reference():
    """BASELINE config 4 shape (COCO 640x480, 64 planes), 8 frames per launch"""
    B, S, H, W = 8, 64, 480, 640
    d = device_batch(B, S, H, W, synth.SEED + 4)
    got = fused(d)
    for b in range(B):
        check(errors(got, reference_frame(d, b), b), f"c4 frame {b}")


@needs_ref
@pytest.mark.parametrize("S", [16, 64, 128])
def test_config5_plane_sweep_against_the_unmodified_reference(S):
    """BASELINE config 5: 1920x1080 at 16 / 64 / 128 planes (|flow| reaches the 200 px clip); S = 128 runs the
    4-stage ring (the plane constants of 2 x 129 planes take the fifth stage's room)"""
    d = device_batch(1, S, 1080, 1920, synth.SEED + S)
    check(errors(fused(d), reference_frame(d, 0)), f"c5 S={S}")


@needs_ref
def test_sample_pairs_at_full_size_against_two_independent_reference_renders():
    """utils.utils.render_sample_pairs: one launch of 2N frames, each MPI read for both of its poses
    (MPIFLOW_POSES_PER_MPI(2)), against two independent renders of the unmodified reference per sample."""
    from utils import utils as U
    from utils.mpi import mpi_rendering as R
    N, S, H, W = 2, 64, 384, 1280
    d = device_batch(N, S, H, W, synth.SEED + 9)
    obj = d["obj_mask"]
    poses = torch.stack((torch.from_numpy(synth.poses(N, "object", 5)).cuda(), d["G"]), 1)      # N x 2 x 4 x 4
    src = d["mpi"][:, 0, 0:3].contiguous()
    blended = U.blend_src_into_mpi(src, d["mpi"][:, :, 0:3], d["mpi"][:, :, 3:], d["disparity"], d["Kinv"])
    masks = torch.stack((obj, 1 - obj), 1).reshape(2 * N, 1, 1, H, W)
    got = R.render_fused(blended, d["mpi"][:, :, 3:], d["disparity"], poses.reshape(2 * N, 4, 4), d["Kinv"], d["K"], masks,
                         poses_per_mpi=2)
    torch.cuda.synchronize()
    for n in range(N):
        mpi_n = torch.cat((blended[n:n + 1], d["mpi"][n:n + 1, :, 3:]), 2)
        for j in range(2):
            with torch.no_grad():
                want = ref_modules.render_frame(mpi_n, d["disparity"][n:n + 1], poses[n:n + 1, j], d["Kinv"][n:n + 1],
                                                d["K"][n:n + 1], masks[2 * n + j].reshape(1, 1, H, W))
            check(errors(got, want, 2 * n + j), f"sample {n} pose {j}")


@needs_ref
def test_sample_pair_with_an_instance_mask_like_the_generator():
    """the generator's mask is ONE instance's binary mask, bilinearly resized with align_corners=True
    (gen_3dphoto_dynamic_v2.py:101-105): constant over most of the frame, soft only along the object's outline; the
    pair renders it as obj and 1 - obj on a shared MPI (utils/utils.py:210-238)"""
    from utils import utils as U
    S, H, W = 64, 384, 1280
    d = device_batch(1, S, H, W, synth.SEED + 21)
    yy, xx = torch.meshgrid(torch.arange(H // 2, dtype=torch.float32), torch.arange(W // 2, dtype=torch.float32), indexing="ij")
    inst = ((((xx - 0.4 * W / 2) / (0.18 * W / 2)) ** 2 + ((yy - 0.55 * H / 2) / (0.3 * H / 2)) ** 2) < 1).float()
    obj = torch.nn.functional.interpolate(inst[None, None].cuda(), size=(H, W), mode="bilinear", align_corners=True)
    assert 0.05 < float(obj.mean()) < 0.5 and float(((obj > 0) & (obj < 1)).float().mean()) < 0.02
    cam_obj = torch.from_numpy(synth.poses(1, "object", 7)).cuda()[0]
    flow_mix, frame_mix, fill_mask, (rgb2, obj2, flow2) = U.render_sample_pair(
        d["mpi"][:, 0, 0:3].contiguous(), obj, d["mpi"], d["disparity"], d["K"], d["K"], cam_obj, d["G"][0], k_src_inv=d["Kinv"])
    torch.cuda.synchronize()
    blended = U.blend_src_into_mpi(d["mpi"][:, 0, 0:3].contiguous(), d["mpi"][:, :, 0:3], d["mpi"][:, :, 3:], d["disparity"], d["Kinv"])
    mpi_b = torch.cat((blended, d["mpi"][:, :, 3:]), 2)
    for j, (pose, m) in enumerate(((cam_obj, obj), (d["G"][0], 1 - obj))):
        with torch.no_grad():
            want = ref_modules.render_frame(mpi_b, d["disparity"], pose[None], d["Kinv"], d["K"], m)
        assert float((rgb2[j:j + 1] - want[0]).abs().max()) <= TOL, j
        assert float((obj2[j:j + 1] - want[4]).abs().max()) <= TOL, j
        assert float((flow2[j:j + 1] - want[3]).abs().max()) <= TOL, j             # raw renders: the clip of :347 comes with the merge


def test_poses_per_mpi_with_footprints_beyond_the_box():
    """MPIFLOW_POSES_PER_MPI(2) with N >= 2 distinct MPIs and a strong zoom-out pose: some footprints do not fit the
    TMA box, those planes are read from global memory and must come from the frame's own MPI (index b / poses)."""
    from utils.mpi import mpi_rendering as R
    N, S, H, W = 3, 6, 64, 128
    inp = synth.frame_inputs(N, S, H, W, "smooth", "identity", 41)
    d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    G = torch.from_numpy(np.stack([synth.poses(2, "random", 60 + n) for n in range(N)])).cuda()   # N x 2 x 4 x 4
    G[:, 1, 2, 3] = 0.9          # second pose of every MPI: camera far behind, footprint ~2x the tile
    G[:, 1, 0, 3] = 0.3
    obj = d["obj_mask"]
    masks = torch.stack((obj, 1 - obj), 1).reshape(2 * N, 1, 1, H, W)
    got = R.render_fused(d["mpi"][:, :, 0:3], d["mpi"][:, :, 3:4], d["disparity"], G.reshape(2 * N, 4, 4), d["Kinv"], d["K"],
                         masks, poses_per_mpi=2)
    rep = lambda t: t.repeat_interleave(2, 0)
    mpi2 = rep(d["mpi"])
    want = R.render_fused(mpi2[:, :, 0:3], mpi2[:, :, 3:4], rep(d["disparity"]), G.reshape(2 * N, 4, 4), rep(d["Kinv"]),
                          rep(d["K"]), masks)
    torch.cuda.synchronize()
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    ref = torch_ref.render_frame(mpi2, rep(d["disparity"]), G.reshape(2 * N, 4, 4), rep(d["Kinv"]), rep(d["K"]),
                                 masks.reshape(2 * N, 1, H, W))
    check(errors(got, ref), "zoom-out pairs")
