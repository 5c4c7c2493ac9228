"""Independently moving object by depth-aware forward splatting -- drop-in for MPI-Flow's moving_obj.py
(moveing_object_with_mask, reference :16-175).

The reference projects on the GPU, copies the index maps, the depth and the image to the host and runs the serial C
loop of external/forward_warping/warping.c through ctypes (:123-129).  Here everything up to the five-channel
splat result stays on the device: the index maps come from the same torch operators in the same order (so they
equal the reference's on the same device class, a13), and the splat is mpiflow_forward_warp_f32 on device pointers
(bit-exact with warping.c).  Only the masks' 3x3 dilation and the TELEA hole filling (:145-162, OpenCV, outside the
path) run on the host.  No CPU fallback for the splat.
"""
import math
import random

import numpy as np
import torch

import _native as nat
from geometry import BackprojectDepth, Project3D, transformation_from_parameters


def draw_object_motion():
    """Translation of the object, drawn like the reference (:72-89): three translation draws, then the three angle
    draws (value + sign each) that the reference makes and discards (:90) -- the `random` stream stays aligned."""
    cix = random.random() * 0.05 + 0.05
    ciy = -1 * (random.random() * 0.05 + 0.05)
    ciz = random.random() * 0.05 + 0.05
    for _ in range(3):
        random.random()
        random.randrange(2)
    return [cix, ciy, ciz]


def _homogeneous(m, device):
    out = torch.zeros((1, 4, 4), device=device)
    out[0, -1, -1] = 1.0
    out[:, :3, :3] = m
    return out


def splat_index_maps(disp, K, inv_K, instance_mask, motion, device=None):
    """a12 + a13 (:28-122): depth = min(1 / (disp + 0.005), 100); static camera projection, object pixels replaced by
    their projection under the object motion; pixel units; truncation and clamp to the image.
    disp 1x1xHxW, instance_mask 1x1xHxW (> 0 = object), K / inv_K 3x3 or 1x3x3, motion = (tx, ty, tz).
    -> (safe_x, safe_y int64 1xHxW, z1 float32 1xHxW, p1 float32 1xHxWx2), all on `device`."""
    device = torch.device(device) if device is not None else disp.device
    disp, instance_mask = disp.to(device), instance_mask.to(device)
    h, w = disp.shape[-2:]
    depth = 1.0 / (disp[0] + 0.005)
    depth[depth > 100] = 100
    inst = instance_mask[0]
    inst2 = torch.stack([inst, inst], -1)                                        # 1xHxWx2
    backproject, project = BackprojectDepth(1, h, w).to(device), Project3D(1, h, w).to(device)
    axisangle = torch.zeros((1, 1, 3), dtype=torch.float32, device=device)
    translation = torch.zeros((1, 3), dtype=torch.int64, device=device)
    T1 = transformation_from_parameters(axisangle, translation)
    K4, Kinv4 = _homogeneous(K.to(device), device), _homogeneous(inv_K.to(device), device)
    cam_points = backproject(depth, Kinv4)
    p1, z1 = project(cam_points, K4, T1)
    z1 = z1.reshape(1, h, w)
    ai = torch.zeros((1, 1, 3), dtype=torch.float32, device=device)              # rotation is zeroed (:90)
    tri = torch.from_numpy(np.array([[list(motion)]])).to(device)                # float64, like the reference
    Ti = transformation_from_parameters(axisangle + ai, translation + tri)
    pi, zi = project(cam_points, K4, Ti)
    p1[inst2 > 0] = pi[inst2 > 0]
    zi = zi.reshape(1, h, w)
    z1[inst2[:, :, :, 0] > 0] = zi[inst2[:, :, :, 0] > 0]
    p1 = (p1 + 1) / 2
    p1[:, :, :, 0] *= w - 1
    p1[:, :, :, 1] *= h - 1
    safe_y = torch.clamp(p1[:, :, :, 1].long(), 0, h - 1)                         # trunc -> int64, clamp (:121-122)
    safe_x = torch.clamp(p1[:, :, :, 0].long(), 0, w - 1)
    return safe_x, safe_y, z1, p1


def forward_warp(img_u8, safe_x, safe_y, z):
    """warping.c:6-33 on device buffers: img HxWx3 uint8, safe_x / safe_y HxW int64, z HxW float32 (CUDA tensors)
    -> HxWx5 uint8 (RGB, valid, collision)."""
    for name, t in (("img_u8", img_u8), ("safe_x", safe_x), ("safe_y", safe_y), ("z", z)):
        if not (isinstance(t, torch.Tensor) and t.is_cuda):
            raise RuntimeError(f"{name} must be a CUDA tensor: libmpiflow_sm100 is CUDA-only (no CPU fallback)")
    h, w = z.shape[-2:]
    img = img_u8.to(torch.uint8).contiguous()
    sx, sy = safe_x.to(torch.int64).contiguous(), safe_y.to(torch.int64).contiguous()
    zz = z.to(torch.float32).contiguous()
    warped = torch.zeros((h, w, 5), dtype=torch.uint8, device=zz.device)          # pre-zeroed by the caller (:123)
    L = nat.lib()
    nbytes = L.mpiflow_forward_warp_workspace_bytes(h, w)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=zz.device)
    nat.check(L.mpiflow_forward_warp_f32(nat.ptr(img), nat.ptr(sx), nat.ptr(sy), nat.ptr(zz), nat.ptr(warped), h, w,
                                         nat.ptr(ws), nbytes, nat.stream(zz)))
    return warped


def moveing_object_with_mask(depth_path, disp, rgb, K, inv_K, instance_mask, i, device="cuda"):
    """Same arguments as the reference (:16; depth_path is unused there too).  rgb: HxWx3 array.  The reference returns
    None and writes a visualisation panel to temp/; this returns what the panel is made of:
    dict(im1_raw HxWx3 u8, im1 HxWx3 u8 (TELEA-filled), masks {H, M, M', P, H'}, flow_01 1xHxWx2, safe_x, safe_y, z)."""
    import cv2  # host-side mask dilation and hole filling only

    h, w = rgb.shape[:2]
    motion = draw_object_motion()
    safe_x, safe_y, z1, p1 = splat_index_maps(disp, K, inv_K, instance_mask, motion, device)
    img = torch.from_numpy(np.expand_dims(rgb, 0)).float().to(device).reshape(-1).to(torch.uint8).reshape(h, w, 3)
    warped = forward_warp(img, safe_x[0], safe_y[0], z1[0]).cpu().numpy()
    masks = {"H": warped[:, :, 3:4], "M": warped[:, :, 4:5]}
    masks["M"] = 1 - (masks["M"] == masks["H"]).astype(np.uint8)                  # invalid or colliding (:139)
    masks["M'"] = cv2.dilate(masks["M"], np.ones((3, 3), np.uint8), iterations=1)
    masks["P"] = (np.expand_dims(masks["M'"], -1) == masks["M"]).astype(np.uint8)
    masks["H'"] = masks["H"] * masks["P"]
    yy, xx = np.meshgrid(range(h), range(w), indexing="ij")
    flow_01 = p1.cpu().numpy() - np.stack((xx, yy), -1).astype(np.float32)
    im1_raw = np.ascontiguousarray(warped[:, :, 0:3])
    im1 = cv2.inpaint(im1_raw, 1 - masks["H"], 3, cv2.INPAINT_TELEA)
    return dict(im1_raw=im1_raw, im1=im1, masks=masks, flow_01=flow_01, safe_x=safe_x, safe_y=safe_y, z=z1)
