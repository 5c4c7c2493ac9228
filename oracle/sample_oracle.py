"""CPU restatement of the sample assembly around the renderer (reference utils/utils.py:159-288) -- TEST
INFRASTRUCTURE ONLY; nothing under mpi-flow_b200/ imports it.

  blend_src(...)        :183-204  source-view blend_weights of the MPI, photograph blended into the plane colours
  assemble_sample(...)  :239-283  u8 conversion, thresholds and merges of the two renders, flow selection

Pinned against the reference's own outputs (tests/golden/sample_pair.npz, made by tests/golden/make_golden.py
--sample-only from the unmodified utils/utils.py) in tests/test_oracle_golden.py.
"""
import numpy as np
import torch


def blend_src(src_img, mpi_rgb, mpi_sigma, disparity, K_inv):
    """torch operators in the reference's order: xyz = (K^-1 . pixel) * depth (mpi_rendering.py:213-239), dist = norm
    of neighbouring planes (last 1e3), T = exp(-sigma dist), blend_weights = exclusive cumprod(T + 1e-6) (:62-99),
    then w * src + (1 - w) * rgb (utils/utils.py:195-197)."""
    B, S, _, H, W = mpi_rgb.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    pix = torch.stack((xs, ys, torch.ones_like(xs)), 0).reshape(1, 3, H * W)
    rays = torch.matmul(K_inv.repeat_interleave(S, 0), pix.expand(B * S, 3, H * W))
    xyz = (rays.reshape(B, S, 3, H * W) * torch.reciprocal(disparity)[:, :, None, None]).reshape(B, S, 3, H, W)
    dist = torch.norm(xyz[:, 1:] - xyz[:, :-1], dim=2, keepdim=True)
    dist = torch.cat((dist, torch.full((B, 1, 1, H, W), 1e3)), 1)
    T = torch.exp(-mpi_sigma * dist)
    acc = torch.cumprod(T + 1e-6, dim=1)
    w = torch.cat((torch.ones_like(acc[:, :1]), acc[:, :-1]), 1)
    return w * src_img.unsqueeze(1) + (1 - w) * mpi_rgb


def to_u8_bgr(chw):
    """float CxHxW in [0,1] -> HxWx3 uint8, channels reversed (utils/utils.py:173-177, :239-250)"""
    v = np.asarray(chw, np.float32).transpose(1, 2, 0)
    return np.clip(np.round(v * 255), 0, 255).astype(np.uint8)[:, :, ::-1]


def assemble_sample(frame, frame_dync, mask, mask_dync, flow, flow_dync, obj_mask, mask_thresh=0.99):
    """frame/flow/mask: render with obj_mask; *_dync: render with 1 - obj_mask; planar float32 (3|2|1,H,W), flows
    already clipped to +-200.  -> flow_mix HxWx2 f32, frame_mix HxWx3 u8 BGR, fill_mask HxW u8."""
    t = np.float32(mask_thresh)
    m, md, obj = (np.asarray(a, np.float32).reshape(frame.shape[-2:]) for a in (mask, mask_dync, obj_mask))
    a, b = to_u8_bgr(frame).copy(), to_u8_bgr(frame_dync).copy()
    fa, fb = (np.asarray(f, np.float32).transpose(1, 2, 0).copy() for f in (flow, flow_dync))
    fa[obj < t] = 0
    fb[obj >= t] = 0
    a[m < t] = 255
    b[md < t] = 255
    frame_mix, flow_mix, fill = b.copy(), fb.copy(), md.copy()
    frame_mix[m >= t] = a[m >= t]
    flow_mix[obj >= t] = fa[obj >= t]
    fill[m >= t] = 1
    return flow_mix, frame_mix, (fill < t).astype(np.uint8)
