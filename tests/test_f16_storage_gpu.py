"""fp16 STORAGE of the MPI (mpiflow_render_fused_f16, mpiflow_blend_src_f16; the generator's
production dtype, gen_3dphoto_dynamic_v2.py:46,59,82-93).

Parity contract: fp32 arithmetic on the fp16-rounded inputs.  The checker is therefore the oracle (the reference's
operator sequence in fp32, oracle/torch_ref.py on the same GPU) fed with the SAME rounded inputs up-cast to fp32:
image, flow and object mask within 1e-4, the validity count exact.  (The reference's own fp16 mode computes in fp16
and quantises sample coordinates to whole pixels beyond x = 1024 -- a quirk, not a target: SURVEY.md section 7.)
"""
import ctypes

import numpy as np
import pytest
import torch

import synth
from oracle import torch_ref

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _inputs(B, S, H, W, kind, pose, seed):
    d = {k: torch.from_numpy(v).cuda() for k, v in synth.frame_inputs(B, S, H, W, kind, pose, seed).items()}
    d["mpi16"] = d["mpi"].half()
    d["obj16"] = d["obj_mask"].half()
    return d


@pytest.mark.parametrize("case", [(2, 8, 48, 96, "smooth", "random", 3), (1, 16, 96, 320, "white", "random", 4),
                                  (3, 5, 33, 47, "white", "object", 5), (1, 64, 384, 1280, "smooth", "random", 77)])
def test_render_fused_f16_matches_the_oracle_on_the_rounded_inputs(case):
    from utils.mpi import mpi_rendering as R
    d = _inputs(*case)
    m16 = d["mpi16"]
    got = R.render_fused(m16[:, :, 0:3], m16[:, :, 3:4], d["disparity"], d["G"], d["Kinv"], d["K"], d["obj16"])
    up, obj_up = m16.float(), d["obj16"].float()
    want = torch_ref.render_frame(up, d["disparity"], d["G"], d["Kinv"], d["K"], obj_up)
    torch.cuda.synchronize()
    err = {n: float((g - w).abs().max()) for n, g, w in zip(["rgb", "depth", "mask", "flow", "obj"], got, want)}
    assert err["rgb"] <= TOL and err["flow"] <= TOL and err["obj"] <= TOL and err["mask"] == 0, (case, err)
    assert all(o.dtype == torch.float32 for o in got)
    # and it agrees with the fp32 entry point on the same values (different kernels: to rounding)
    f32 = R.render_fused(up[:, :, 0:3], up[:, :, 3:4], d["disparity"], d["G"], d["Kinv"], d["K"], obj_up)
    for a, b in zip(got, f32):
        assert float((a - b).abs().max()) <= 1e-4


def test_f16_entry_is_not_an_upcast_in_disguise():
    """the half tensors go to the library as they are: strided views of the fp16 MPI, no fp32 copy of it is made"""
    import _native as nat
    from utils.mpi import mpi_rendering as R
    d = _inputs(1, 8, 64, 128, "smooth", "random", 9)
    m16 = d["mpi16"]
    seen = {}
    real = nat.lib().mpiflow_render_fused_f16

    class Spy:
        def __call__(self, *a):
            seen["rgb_ptr"] = a[0].value
            return real(*a)
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    L = nat.lib()
    orig = L.mpiflow_render_fused_f16
    try:
        L.mpiflow_render_fused_f16 = Spy()
        R.render_fused(m16[:, :, 0:3], m16[:, :, 3:4], d["disparity"], d["G"], d["Kinv"], d["K"], d["obj16"])
    finally:
        L.mpiflow_render_fused_f16 = orig
    assert seen["rgb_ptr"] == m16.data_ptr()
    assert torch.cuda.max_memory_allocated() - base < m16.numel() * 2      # nothing the size of an fp32 copy


def test_host_f16_entry_equals_device_f16_entry():
    import _native as nat
    from utils.mpi import mpi_rendering as R
    d = _inputs(3, 6, 40, 72, "smooth", "random", 15)
    m16 = d["mpi16"]
    want = R.render_fused(m16[:, :, 0:3], m16[:, :, 3:4], d["disparity"], d["G"], d["Kinv"], d["K"], d["obj16"])
    torch.cuda.synchronize()
    B, S, _, H, W = m16.shape
    h_mpi, h_obj = m16.cpu().numpy(), d["obj16"].cpu().numpy()
    small = {k: d[k].cpu().numpy() for k in ("disparity", "G", "Kinv", "K")}
    outs = [np.empty((B, c, H, W), np.float32) for c in (3, 1, 1, 2, 1)]
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = nat.lib().mpiflow_render_fused_host_f16(P(h_mpi), P(h_obj), P(small["disparity"]), P(small["G"]), P(small["Kinv"]),
                                                 P(small["K"]), B, S, H, W, 0, *[P(o) for o in outs])
    assert rc == 0
    for a, b in zip(outs, want):
        np.testing.assert_array_equal(a, b.cpu().numpy())


def test_blend_src_f16_is_the_f32_blend_of_the_rounded_mpi():
    from utils import utils as U
    d = _inputs(2, 7, 40, 64, "smooth", "random", 21)
    m16 = d["mpi16"]
    src16 = torch.rand((2, 3, 40, 64), device="cuda").half()
    rgb, sig = U.blend_src_mpi(src16, m16[:, :, 0:3], m16[:, :, 3:4], d["disparity"], d["Kinv"])
    assert rgb.dtype == torch.float32 and sig.dtype == torch.float32
    assert rgb.data_ptr() + 3 * 40 * 64 * 4 == sig.data_ptr()              # one fp32 MPI [B,S,4,H,W], two views
    up = m16.float()
    want = U.blend_src_into_mpi(src16.float(), up[:, :, 0:3], up[:, :, 3:4], d["disparity"], d["Kinv"])
    torch.cuda.synchronize()
    assert torch.equal(rgb, want)
    assert torch.equal(sig, up[:, :, 3:4])


def test_sample_pair_from_an_fp16_mpi():
    """render_sample_pair on the generator's dtypes equals the same call on the up-cast tensors"""
    from utils import utils as U
    d = _inputs(1, 8, 48, 96, "smooth", "random", 23)
    m16 = d["mpi16"]
    src16 = torch.rand((1, 3, 48, 96), device="cuda").half()
    G2 = torch.from_numpy(synth.poses(2, "random", 31)).cuda()
    a = U.render_sample_pair(src16, d["obj16"], m16, d["disparity"].half(), d["K"].half(), d["K"].half(), G2[0], G2[1],
                             k_src_inv=d["Kinv"])
    b = U.render_sample_pair(src16.float(), d["obj16"].float(), m16.float(), d["disparity"].half().float(), d["K"].half().float(),
                             d["K"].half().float(), G2[0], G2[1], k_src_inv=d["Kinv"])
    torch.cuda.synchronize()
    for x, y in zip(a[:3], b[:3]):
        assert torch.equal(x, y)
