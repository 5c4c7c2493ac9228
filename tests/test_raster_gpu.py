"""Mesh rasteriser behind warpback.utils.RGBDRenderer.render_mesh (csrc/raster.cu) against oracle/raster_oracle.py,
the numpy restatement of pytorch3d's published naive rasteriser.  PARITY UNPINNED: pytorch3d itself is not available
(SURVEY.md 8c), so these tests pin the CUDA kernels to the restatement, not to the reference's dependency."""
import numpy as np
import pytest
import torch

from oracle import raster_oracle as ro

pytestmark = pytest.mark.gpu


def _mesh(b, h, w, seed, jitter=0.3):
    from warpback.utils import RGBDRenderer
    g = torch.Generator().manual_seed(seed)
    rgbd = torch.rand((b, 4, h, w), generator=g)
    rgbd[:, 3] = 0.2 + 0.6 * rgbd[:, 3]
    rgbd[:, 3, :, w // 2:] *= 0.5                               # depth discontinuity -> occlusions after the warp
    cam_int = torch.tensor([[0.58, 0, 0.5], [0, 0.58, 0.5], [0, 0, 1]]).repeat(b, 1, 1)
    cam_ext = torch.tensor([[1., 0., 0., jitter], [0., 1., 0., -0.1], [0., 0., 1., 0.05]]).repeat(b, 1, 1)
    r = RGBDRenderer(torch.device("cuda"))
    return r, rgbd.cuda(), cam_int.cuda(), cam_ext.cuda()


def _ndc_and_attrs(r, mesh, cam_int, cam_ext):
    """the host maths of render_mesh up to the rasteriser call (reference :31-50)"""
    h, w = mesh["size"]
    world = torch.matmul(cam_ext.unsqueeze(1), r.lift_to_homo(mesh["vertice"])[..., None]).squeeze(-1)
    attrs = torch.cat([mesh["attributes"], world[..., -1:]], dim=-1)
    ndc = torch.matmul(r.get_perspective_from_intrinsic(cam_int).unsqueeze(1), r.lift_to_homo(world)[..., None]).squeeze(-1)
    ndc = ndc[..., :-1] / ndc[..., -1:]
    ndc[..., :-1] *= -1
    ndc[..., 0] *= w / h
    return ndc, attrs


@pytest.mark.parametrize("b,h,w,seed", [(1, 12, 20, 1), (2, 16, 16, 2), (1, 24, 10, 3)])
def test_rasteriser_matches_restatement(b, h, w, seed):
    from warpback.utils import rasterize_and_interpolate
    r, rgbd, cam_int, cam_ext = _mesh(b, h, w, seed)
    mesh = r.construct_mesh(rgbd, cam_int)
    ndc, attrs = _ndc_and_attrs(r, mesh, cam_int, cam_ext)
    out, p2f, zbuf, bary = rasterize_and_interpolate(ndc, mesh["faces"], attrs, h, w, return_aux=True)
    F = mesh["faces"].shape[1]
    for i in range(b):
        want = ro.rasterize(ndc[i].cpu().numpy(), mesh["faces"][i].cpu().numpy(), attrs[i].cpu().numpy(), h, w)
        got_f = p2f[i].cpu().numpy()
        got_f = np.where(got_f >= 0, got_f - i * F, -1)
        same = got_f == want["pix_to_face"]
        assert same.mean() >= 0.995, same.mean()                  # an fp32 tie on a shared edge may pick the neighbour
        np.testing.assert_allclose(zbuf[i].cpu().numpy()[same], want["zbuf"][same], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(bary[i].cpu().numpy()[same], want["bary"][same], rtol=0, atol=2e-4)
        np.testing.assert_allclose(out[i].cpu().numpy()[:, same], want["out"][:, same], rtol=0, atol=2e-4)
        # hits and holes (portrait images cover little: the reference scales x by w/h whatever the orientation, :50)
        assert (want["pix_to_face"] >= 0).mean() > 0.1 and (want["pix_to_face"] < 0).any()


def test_identity_camera_reproduces_the_image():
    r, rgbd, cam_int, _ = _mesh(1, 32, 48, 4)
    ramp = torch.linspace(0.3, 0.6, 48, device="cuda")[None, :].expand(32, 48)
    rgbd[0, 3] = ramp                                            # smooth disparity: the Sobel mask keeps (almost) everything
    cam_ext = torch.eye(4, device="cuda")[None, :3]
    render, disp, mask = r.render_mesh(r.construct_mesh(rgbd, cam_int), cam_int, cam_ext)
    vis = mask[0, 0] > 0.99                                       # away from the cut depth edge
    assert vis.float().mean().item() > 0.8
    assert (render[0] - rgbd[0, :3])[:, vis].abs().max().item() < 1e-3
    assert (disp[0, 0] - rgbd[0, 3])[vis].abs().max().item() < 1e-3
