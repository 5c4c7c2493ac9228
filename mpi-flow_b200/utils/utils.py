"""Sample assembly -- drop-in for the rendering half of MPI-Flow's utils/utils.py (gen_swing_path :55,
render_novel_view :65, generate_random_pose :121, render_3dphoto_dynamic :159, render_novel_view_dynamic :291;
same names, positional orders and return tuples).  SURVEY.md 8(f) row 1.

One dataset sample is, on the device:
  1. mpiflow_blend_src_f32       source-view transmittance + blend of the photograph into the plane colours
                                 (reference: get_src_xyz + render + three BxSx3xHxW temporaries, :183-204)
  2. mpiflow_render_fused_f32    BOTH poses of the sample in one launch: the MPI is passed once with a batch stride
                                 of 0, the poses and the two masks (obj, 1 - obj) are per frame; frames of the same
                                 tile are neighbours in the tile queue so the second one finds its boxes in L2
                                 (reference: two render_novel_view_dynamic calls, :210-238)
  3. mpiflow_assemble_sample_f32 flow clip, thresholds, merges, u8 conversion (reference: numpy on the host after
                                 eight device->host copies, :239-283)
and only flow_mix, frame_mix and fill_mask cross PCIe.  No CPU fallback: tensors must live on a CUDA device.

The module also carries the two loaders and the arrow overlay the generator imports from here
(gen_3dphoto_dynamic_v2.py:10-14; reference :25-52).  `utils/` has no __init__.py, like the reference's: it is a
namespace portion, so with this directory first on sys.path `utils.utils` and `utils.mpi.*` resolve here while
`utils.transform`, `utils.flow_viz`, ... still resolve to the reference's files further down the path.
"""
import math
import os
import random

import numpy as np
import torch

import _native as nat
from geometry import transformation_from_parameters
from utils.mpi import mpi_rendering
from utils.mpi.homography_sampler import HomographySample


def arrowon(img, flow, step=32):
    """Copy of img with one arrow per step x step cell along the flow (debug overlay); reference :25-32."""
    import cv2
    h, w = flow.shape[:2]
    out = img.copy()
    for i in range(step // 2, h, step):
        for j in range(step // 2, w, step):
            tip = (int(j + flow[i, j, 0]), int(i + flow[i, j, 1]))
            cv2.arrowedLine(out, (j, i), tip, (0, 0, 255), 2, 8, 0, 0.2)
    return out


def image_to_tensor(img_path, unsqueeze=True):
    """RGB image file -> float32 [1,]3xHxW in [0,1] (what torchvision's ToTensor gives); reference :35-39."""
    from PIL import Image
    rgb = np.array(Image.open(img_path).convert("RGB"), dtype=np.uint8)
    rgb = torch.from_numpy(rgb).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    return rgb.unsqueeze(0) if unsqueeze else rgb


def disparity_to_tensor(disp_path, unsqueeze=True):
    """8-bit grey disparity file -> float32 [1,]1xHxW in [0,1] (fp64 division, then rounded); reference :42-52."""
    import cv2
    disp = torch.from_numpy(cv2.imread(disp_path, 0) / 255)[None, ...]
    return (disp.unsqueeze(0) if unsqueeze else disp).float()


def gen_swing_path(num_frames=90, r_x=0.14, r_y=0.0, r_z=0.10):
    """Tuple of num_frames 4x4 camera poses on the 3D-photo swing; reference :55-62."""
    phase = 2.0 * math.pi * (torch.arange(num_frames) / (num_frames - 1))
    poses = torch.eye(4).repeat(num_frames, 1, 1)
    poses[:, 0, 3] = r_x * torch.sin(phase)
    poses[:, 1, 3] = r_y * torch.cos(phase)
    poses[:, 2, 3] = r_z * (torch.cos(phase) - 1.0)
    return poses.unbind()


def generate_random_pose(ext_cz, base_motions=[0.1, 0.1, 0.1], device="cuda"):
    """Random camera (base_motions 0.1) or object (base_motions 0) motion as a 4x4 matrix; reference :121-156.
    Draws from the `random` module in the reference's order, so a seeded run picks the same poses."""
    sign = lambda: (-1) ** random.randrange(2)
    sx, sy, sz = sign(), sign(), sign()
    if base_motions[0] == 0.1:
        sz = -1                       # cameras move forward
    else:
        sx, sy, sz = sx * 0.5, sy * 0.5, sz * 0.5
    tx = (random.random() * 0.1 + base_motions[0]) * sx
    ty = (random.random() * 0.1 + base_motions[1]) * sy
    tz = (random.random() * ext_cz + base_motions[2]) * sz
    ax_s, ay_s, az_s = sign(), sign(), sign()
    ax = (random.random() * math.pi / 36.0) * ax_s
    ay = (random.random() * math.pi / 36.0) * ay_s
    az = (random.random() * math.pi / 36.0) * az_s
    axisangle = torch.from_numpy(np.array([[[ax * 0.4, ay * 0.4, az * 0.4]]], dtype=np.float32)).to(device).float()
    translation = torch.from_numpy(np.array([[[tx, ty, tz]]])).to(device).float()
    return transformation_from_parameters(axisangle, translation)[0]


def _batched(m, B, n):
    m = m if m.dim() == 3 else m.unsqueeze(0)
    return m.expand(B, n, n)


def render_novel_view(mpi_all_rgb_src, mpi_all_sigma_src, disparity_all_src, G_tgt_src, K_src_inv, K_tgt, K_src,
                      src_pose, homography_sampler, hard_flow=False):
    """-> (tgt_imgs_syn, tgt_depth_syn, flow_syn); reference :65-118.  K_src, src_pose and the sampler are unused by
    the arithmetic there as well (the per-plane flow volume it builds is discarded by render_tgt_rgb_depth)."""
    B = mpi_all_rgb_src.shape[0]
    rgb, depth, _, flow, _ = mpi_rendering.render_fused(
        mpi_all_rgb_src, mpi_all_sigma_src, disparity_all_src, _batched(G_tgt_src, B, 4), K_src_inv, K_tgt,
        None, False, hard_flow)
    return rgb, depth, flow


def render_novel_view_dynamic(obj_mask, mpi_all_rgb_src, mpi_all_sigma_src, disparity_all_src, G_tgt_src, K_src_inv,
                              K_tgt, K_src, src_pose, homography_sampler, hard_flow=False):
    """-> (tgt_imgs_syn, tgt_depth_syn, flow_syn clipped to +-200, obj_mask warped and composited);
    reference :291-348."""
    B = mpi_all_rgb_src.shape[0]
    rgb, depth, _, flow, obj = mpi_rendering.render_fused(
        mpi_all_rgb_src, mpi_all_sigma_src, disparity_all_src, _batched(G_tgt_src, B, 4), K_src_inv, K_tgt,
        obj_mask, False, hard_flow)
    return rgb, depth, torch.clip(flow, -200, 200), obj


def blend_src_mpi(src_imgs, mpi_all_rgb_src, mpi_all_sigma_src, disparity_all_src, k_src_inv):
    """blend_weights * src + (1 - blend_weights) * rgb with the source-view blend_weights of the MPI; reference
    :183-204.  -> (rgb BxSx3xHxW, sigma BxSx1xHxW), fp32, the MPI the renderer takes.
    An fp16 MPI (the generator's production dtype, gen_3dphoto_dynamic_v2.py:82-93) is read as fp16 by
    mpiflow_blend_src_f16, which writes blended colours and sigma as ONE fp32 tensor BxSx4xHxW: no up-cast pass; the
    two results are views of it.  An fp32 MPI keeps its sigma and gets a new BxSx3xHxW colour tensor."""
    B, S, _, H, W = mpi_all_rgb_src.shape
    half = mpi_all_rgb_src.dtype == torch.float16 and mpi_all_sigma_src.dtype == torch.float16
    rgb, sig = nat.dev_storage(mpi_all_rgb_src, "mpi_all_rgb_src"), nat.dev_storage(mpi_all_sigma_src, "mpi_all_sigma_src")
    if not half:
        rgb, sig = nat.dev(rgb), nat.dev(sig)
    if not nat.inner_contiguous(rgb):
        rgb = rgb.contiguous()
    if not nat.inner_contiguous(sig):
        sig = sig.contiguous()
    src = nat.dev_storage(src_imgs, "src_imgs") if half else nat.dev(src_imgs, "src_imgs")
    src = src.to(rgb.dtype).expand(B, 3, H, W).contiguous()
    disp = nat.dev(disparity_all_src, "disparity_all_src").contiguous()
    kinv = nat.dev(k_src_inv.to(rgb.device), "k_src_inv").expand(B, 3, 3).contiguous()
    out = torch.empty((B, S, 4 if half else 3, H, W), dtype=torch.float32, device=rgb.device)
    fn = nat.lib().mpiflow_blend_src_f16 if half else nat.lib().mpiflow_blend_src_f32
    nat.check(fn(nat.ptr(rgb), rgb.stride(0), rgb.stride(1), rgb.stride(2), nat.ptr(sig), sig.stride(0), sig.stride(1),
                 nat.ptr(src), nat.ptr(disp), nat.ptr(kinv), B, S, H, W, nat.ptr(out), nat.stream(rgb)))
    return (out[:, :, 0:3], out[:, :, 3:4]) if half else (out, sig)


def blend_src_into_mpi(src_imgs, mpi_all_rgb_src, mpi_all_sigma_src, disparity_all_src, k_src_inv):
    """The blended plane colours alone (BxSx3xHxW fp32); see blend_src_mpi."""
    return blend_src_mpi(src_imgs, mpi_all_rgb_src, mpi_all_sigma_src, disparity_all_src, k_src_inv)[0]


def render_sample_pair(src_imgs, obj_mask, mpi_all_src, disparity_all_src, k_src, k_tgt, cam_ext, cam_ext_dynamic,
                       hard_flow=False, mask_thresh=0.99, k_src_inv=None):
    """The device part of render_3dphoto_dynamic for given poses (cam_ext moves the masked object, cam_ext_dynamic
    the rest).  -> device tensors (flow_mix HxWx2 f32, frame_mix HxWx3 u8 BGR, fill_mask HxW u8) plus the two raw
    renders (rgb 2x3xHxW, obj 2x1xHxW, flow 2x2xHxW) for inspection.  k_src_inv: the fp64-inverted intrinsics if the
    caller already has them (the reference inverts on the host for every sample, which costs a device sync)."""
    if mpi_all_src.shape[0] != 1:
        raise ValueError("one sample at a time, like the reference (it reads index 0 of every output)")
    _, S, _, H, W = mpi_all_src.shape
    device = mpi_all_src.device
    if k_src_inv is None:
        k_src_inv = torch.inverse(k_src.to(torch.float64).cpu()).to(device).to(k_src.dtype)  # :186-187
    mpi_rgb, mpi_sigma = mpi_all_src[:, :, 0:3], mpi_all_src[:, :, 3:]
    blended, mpi_sigma = blend_src_mpi(src_imgs, mpi_rgb, mpi_sigma, disparity_all_src, k_src_inv)
    obj = nat.dev(obj_mask, "obj_mask").reshape(1, 1, H, W)
    masks = torch.cat((obj, 1 - obj), 0).reshape(2, 1, 1, H, W)                            # :212, :226
    poses = torch.stack((cam_ext.to(device).float(), cam_ext_dynamic.to(device).float()), 0)
    rgb2, _, _, flow2, obj2 = mpi_rendering.render_fused(
        blended.expand(2, S, 3, H, W), nat.dev(mpi_sigma).expand(2, S, 1, H, W), disparity_all_src.expand(2, S),
        poses, k_src_inv.expand(2, 3, 3), k_tgt.expand(2, 3, 3), masks, False, hard_flow)
    flow_mix = torch.empty((H, W, 2), dtype=torch.float32, device=device)
    frame_mix = torch.empty((H, W, 3), dtype=torch.uint8, device=device)
    fill_mask = torch.empty((H, W), dtype=torch.uint8, device=device)
    obj_c = obj.contiguous()
    nat.check(nat.lib().mpiflow_assemble_sample_f32(
        nat.ptr(rgb2[0]), nat.ptr(rgb2[1]), nat.ptr(obj2[0]), nat.ptr(obj2[1]), nat.ptr(flow2[0]), nat.ptr(flow2[1]),
        nat.ptr(obj_c), float(mask_thresh), H, W, nat.ptr(flow_mix), nat.ptr(frame_mix), nat.ptr(fill_mask),
        nat.stream(rgb2)))
    return flow_mix, frame_mix, fill_mask, (rgb2, obj2, flow2)


def render_sample_pairs(src_imgs, obj_masks, mpi_all_src, disparity_all_src, k_src_inv, k_tgt, poses, hard_flow=False,
                        mask_thresh=0.99):
    """N samples per launch (dataset-scale use; the reference has no batched form): src_imgs Nx3xHxW, obj_masks
    Nx1xHxW, mpi_all_src NxSx4xHxW, disparity NxS, k_src_inv / k_tgt Nx3x3, poses Nx2x4x4 (object pose, camera pose).
    One blend launch, ONE render launch of 2N frames (MPIFLOW_POSES_PER_MPI(2): each MPI is read for both of its
    poses), N merge launches.  -> flow_mix NxHxWx2 f32, frame_mix NxHxWx3 u8 BGR, fill_mask NxHxW u8 (device)."""
    N, S, _, H, W = mpi_all_src.shape
    device = mpi_all_src.device
    mpi_rgb, mpi_sigma = mpi_all_src[:, :, 0:3], mpi_all_src[:, :, 3:]
    blended, mpi_sigma = blend_src_mpi(src_imgs, mpi_rgb, mpi_sigma, disparity_all_src, k_src_inv)
    obj = nat.dev(obj_masks, "obj_masks").reshape(N, 1, H, W).contiguous()
    masks = torch.stack((obj, 1 - obj), 1).reshape(2 * N, 1, 1, H, W)
    rgb2, _, _, flow2, obj2 = mpi_rendering.render_fused(
        blended, nat.dev(mpi_sigma), disparity_all_src, poses.to(device).float().reshape(2 * N, 4, 4), k_src_inv,
        k_tgt, masks, False, hard_flow, poses_per_mpi=2)
    flow_mix = torch.empty((N, H, W, 2), dtype=torch.float32, device=device)
    frame_mix = torch.empty((N, H, W, 3), dtype=torch.uint8, device=device)
    fill_mask = torch.empty((N, H, W), dtype=torch.uint8, device=device)
    L, st = nat.lib(), nat.stream(rgb2)
    for n in range(N):
        a, b = 2 * n, 2 * n + 1
        nat.check(L.mpiflow_assemble_sample_f32(
            nat.ptr(rgb2[a]), nat.ptr(rgb2[b]), nat.ptr(obj2[a]), nat.ptr(obj2[b]), nat.ptr(flow2[a]), nat.ptr(flow2[b]),
            nat.ptr(obj[n]), float(mask_thresh), H, W, nat.ptr(flow_mix[n]), nat.ptr(frame_mix[n]),
            nat.ptr(fill_mask[n]), st))
    return flow_mix, frame_mix, fill_mask


_inpaint_ws = {}     # (device index, stream, n, H, W) -> workspace tensor; reused across samples


def inpaint_ns(frame, fill_mask, radius=3, out=None):
    """cv2.inpaint(frame, fill_mask, radius, cv2.INPAINT_NS) (reference :284-286) on the device, bit for bit:
    frame u8 [H,W,3] or [N,H,W,3] (interleaved, as cv2 takes it), fill_mask u8 [H,W] or [N,H,W] (non-zero = fill).
    -> u8 tensor of frame's shape (out=frame fills in place).  A batch marches its N images on N warps of one
    launch (csrc/inpaint.cu); no host synchronisation."""
    if not (isinstance(frame, torch.Tensor) and frame.is_cuda and isinstance(fill_mask, torch.Tensor) and fill_mask.is_cuda):
        raise RuntimeError("inpaint_ns: frame and fill_mask must be CUDA tensors (no CPU fallback)")
    if frame.dtype != torch.uint8 or fill_mask.dtype != torch.uint8:
        raise TypeError("inpaint_ns: uint8 image and mask, like cv2.inpaint")
    single = frame.dim() == 3
    img = (frame.unsqueeze(0) if single else frame).contiguous()
    msk = (fill_mask.unsqueeze(0) if single else fill_mask).contiguous()
    N, H, W, C = img.shape
    if C != 3 or tuple(msk.shape) != (N, H, W):
        raise ValueError(f"inpaint_ns: image {tuple(frame.shape)} / mask {tuple(fill_mask.shape)}")
    L = nat.lib()
    key = (img.device.index, torch.cuda.current_stream(img.device).cuda_stream, N, H, W)   # one per stream: calls on two streams must not share scratch
    ws = _inpaint_ws.get(key)
    if ws is None:
        if len(_inpaint_ws) >= 8:
            _inpaint_ws.clear()
        ws = torch.empty(L.mpiflow_inpaint_ns_workspace_bytes(N, H, W), dtype=torch.uint8, device=img.device)
        _inpaint_ws[key] = ws
    if out is not None and (out is not frame or img.data_ptr() != frame.data_ptr()):
        raise ValueError("inpaint_ns: out must be None or the (contiguous) frame itself (in place)")
    res = img if out is frame else torch.empty_like(img)
    nat.check(L.mpiflow_inpaint_ns_u8(nat.ptr(img), nat.ptr(msk), N, H, W, int(round(radius)), nat.ptr(res), nat.ptr(ws),
                                      ws.numel(), nat.stream(img)))
    return res[0] if single else res


def render_3dphoto_dynamic(opt, src_imgs, obj_mask, disp, mpi_all_src, disparity_all_src, k_src, k_tgt,
                           data_path=None, name=None, hard_flow=False, mask_thresh=0.99):
    """One training pair from one MPI: -> (flow_mix HxWx2 f32, src_np HxWx3 u8 BGR, inpainted HxWx3 u8 BGR, None);
    reference :159-288.  `disp`, `data_path` and `name` are unused there too (name only has its suffix cut)."""
    if name is not None:
        name = name.split(".")[0]
    src_np = (nat.dev(src_imgs, "src_imgs")[0].permute(1, 2, 0) * 255).round().clamp(0, 255).to(torch.uint8)
    src_np = src_np.flip(-1).contiguous().cpu().numpy()                                     # :173-177
    device = mpi_all_src.device
    cam_ext_dynamic = generate_random_pose(opt.ext_cz, device=device)                       # :206-208, same draw order
    cam_ext = generate_random_pose(opt.ext_cz, base_motions=[0, 0, 0], device=device)
    flow_mix, frame_mix, fill_mask, _ = render_sample_pair(
        src_imgs, obj_mask, mpi_all_src, disparity_all_src, k_src, k_tgt, cam_ext, cam_ext_dynamic, hard_flow,
        mask_thresh)
    # :284-286.  One sample at a time the reference's own host call is the faster of two bit-identical routes (a single
    # image is one serial march: 7-12 ms in OpenCV on a host core, 5-25 ms on one warp of the GPU); batches belong on
    # the device (inpaint_ns on N images, sample_writer.SampleWriter(device_inpaint=True)).  MPIFLOW_INPAINT=device
    # keeps this call on the GPU as well.
    if os.environ.get("MPIFLOW_INPAINT", "host") == "device":
        inpainted = inpaint_ns(frame_mix, fill_mask, 3).cpu().numpy()
    else:
        import cv2
        inpainted = cv2.inpaint(frame_mix.cpu().numpy(), fill_mask.cpu().numpy(), 3, cv2.INPAINT_NS)
    return flow_mix.cpu().numpy(), src_np, inpainted, None
