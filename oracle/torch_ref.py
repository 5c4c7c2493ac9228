"""Plain-torch restatement of the reference's render_tgt_rgb_depth pipeline -- TEST INFRASTRUCTURE ONLY.

Uses the same library operators the reference hits (torch.matmul -> cuBLAS/MKL, F.grid_sample -> ATen,
torch.inverse in fp64 on the CPU, cumprod/exp/norm/sum), in the same order, so that on the B200 box
it reproduces the arithmetic of "the reference PyTorch/grid_sample path" on CUDA, where
/root/reference itself is not available.  Written from the behaviour documented in SURVEY.md
section 8(a) (utils/mpi/homography_sampler.py:80-220, utils/mpi/mpi_rendering.py:62-139, :213-349);
pinned on CPU against the reference's own outputs in tests/test_oracle_golden.py::test_torch_ref_*.
It is also the op-for-op CPU cost model used by bench.py --impl reference.
"""
import torch
import torch.nn.functional as F


def _grid(H, W, device, dtype):
    ys, xs = torch.meshgrid(torch.arange(H, device=device, dtype=dtype),
                            torch.arange(W, device=device, dtype=dtype), indexing="ij")
    return torch.stack((xs, ys, torch.ones_like(xs)), 0)  # 3,H,W


def plane_xyz(disparity, Kinv, G, H, W):
    """a2 + a3 -> (xyz_src, xyz_tgt), both B,S,3,H,W"""
    B, S = disparity.shape
    pix = _grid(H, W, disparity.device, disparity.dtype).reshape(1, 3, H * W)
    rays = torch.matmul(Kinv.repeat_interleave(S, 0), pix.expand(B * S, 3, H * W))
    xyz_src = rays.reshape(B, S, 3, H * W) * torch.reciprocal(disparity)[:, :, None, None]
    homo = torch.cat((xyz_src.reshape(B * S, 3, -1), torch.ones(B * S, 1, H * W, device=pix.device, dtype=pix.dtype)), 1)
    xyz_tgt = torch.matmul(G.repeat_interleave(S, 0), homo)[:, :3]
    return xyz_src.reshape(B, S, 3, H, W), xyz_tgt.reshape(B, S, 3, H, W)


def homographies(G, Kinv, Kt, depth):
    """N-batched H_tgt_src (fp32 chain of matmuls) and its fp64-CPU inverse"""
    R, t = G[:, :3, :3], G[:, :3, 3]
    n = torch.tensor([0.0, 0.0, 1.0], device=G.device, dtype=G.dtype).expand(G.shape[0], 3)
    Rtnd = R - torch.matmul(t.unsqueeze(2), n.unsqueeze(1)) / -depth.reshape(-1, 1, 1).expand(-1, 3, 3)
    Hts = torch.matmul(Kt, torch.matmul(Rtnd, Kinv))
    Hst = torch.inverse(Hts.double().cpu()).to(Hts.device).to(Hts.dtype)
    return Hts, Hst


def warp(src, Hst):
    """a4: src N,C,H,W -> (warped N,C,H,W, valid N,H,W)"""
    N, C, H, W = src.shape
    pix = _grid(H, W, src.device, src.dtype).reshape(1, 3, H * W).expand(N, 3, H * W)
    m = torch.matmul(Hst, pix).reshape(N, 3, H, W).permute(0, 2, 3, 1)
    pos = m[..., 0:2] / m[..., 2:]
    valid = (pos[..., 0] < W) & (pos[..., 0] > -1) & (pos[..., 1] < H) & (pos[..., 1] > -1)
    gx = (pos[..., 0] + 0.5) / (W * 0.5) - 1
    gy = (pos[..., 1] + 0.5) / (H * 0.5) - 1
    out = F.grid_sample(src, torch.stack((gx, gy), -1), padding_mode="border", align_corners=False)
    return out, valid


def plane_flow(Hts, H, W):
    """a5 -> N,2,H,W"""
    N = Hts.shape[0]
    pix = _grid(H, W, Hts.device, Hts.dtype).reshape(1, 3, H * W).expand(N, 3, H * W)
    m = torch.matmul(Hts, pix)
    return (m[:, 0:2] / m[:, 2:] - pix[:, 0:2]).reshape(N, 2, H, W)


def _weights(sigma, xyz):
    B, S, _, H, W = sigma.shape
    dist = torch.norm(xyz[:, 1:] - xyz[:, :-1], dim=2, keepdim=True)
    dist = torch.cat((dist, torch.full((B, 1, 1, H, W), 1e3, device=xyz.device, dtype=xyz.dtype)), 1)
    T = torch.exp(-sigma * dist)
    acc = torch.cumprod(T + 1e-6, dim=1)
    acc = torch.cat((torch.ones_like(acc[:, :1]), acc[:, :-1]), 1)
    return acc * (1 - T)


def render_frame(mpi, disparity, G, Kinv, Kt, obj_mask, is_bg_depth_inf=False, hard_flow=False):
    """mpi B,S,4,H,W; obj_mask B,1,H,W -> (rgb, depth, mask, flowA2B, obj) like render_tgt_rgb_depth"""
    B, S, _, H, W = mpi.shape
    xyz_src, xyz_tgt = plane_xyz(disparity, Kinv, G, H, W)
    rep = lambda m: m.repeat_interleave(S, 0)
    Hts, Hst = homographies(rep(G), rep(Kinv), rep(Kt), torch.reciprocal(disparity).reshape(-1))
    stacked = torch.cat((mpi, xyz_tgt, obj_mask[:, None].expand(B, S, 1, H, W)), 2).reshape(B * S, 8, H, W)
    warped, valid = warp(stacked, Hst)
    warped = warped.reshape(B, S, 8, H, W)
    rgb, sigma, xyz, obj = warped[:, :, 0:3], warped[:, :, 3:4], warped[:, :, 4:7], warped[:, :, 7:8]
    sigma = torch.where(xyz[:, :, 2:3] >= 0, sigma, torch.zeros_like(sigma))
    w = _weights(sigma, xyz)
    wsum = w.sum(1)
    out_rgb = (w * rgb).sum(1)
    z = (w * xyz[:, :, 2:3]).sum(1)
    out_depth = z + (1 - wsum) * 1000 if is_bg_depth_inf else z / (wsum + 1e-5)
    out_obj = (w * obj).sum(1)
    ws = _weights(mpi[:, :, 3:4], xyz_src)
    flow = plane_flow(Hts, H, W).reshape(B, S, 2, H, W)
    if hard_flow:
        ws = torch.zeros_like(ws).scatter_(1, ws.max(1, keepdim=True)[1], 1.0)
    out_flow = (ws * flow).sum(1)
    mask = valid.reshape(B, S, H, W).float().sum(1, keepdim=True)
    return out_rgb, out_depth, mask, out_flow, out_obj
