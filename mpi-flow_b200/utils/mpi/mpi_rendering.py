"""MPI renderer -- drop-in for MPI-Flow's utils/mpi/mpi_rendering.py (same ten functions, same
positional orders and return tuples: reference :7,42,62,102,142,157,180,213,242,259).

Everything heavy runs in libmpiflow_sm100.so.  render_tgt_rgb_depth() -- the generator's hot call,
utils/utils.py:331 -- is ONE fused sm_100a kernel when its xyz arguments are the plane-derived
tensors returned by get_src_xyz_from_plane_disparity / get_tgt_xyz_from_plane_disparity (which they
always are in the reference's callers).  Those two return DEFERRED tensors (_lazy.LazyTensor: right shape,
dtype and device, values produced by their kernel only if something reads them), so the reference's call
sequence utils/utils.py:304-331 costs one fused launch and no xyz traffic; for any other xyz tensors the same
result is produced by the un-fused CUDA operators (warp kernel + compositing kernels).  No torch/CPU fallback.
"""
import torch

import _native as nat
from _lazy import LazyTensor
from utils.mpi.homography_sampler import HomographySample
from utils.mpi.rendering_utils import transform_G_xyz, sample_pdf, gather_pixel_by_pxpy  # noqa: F401

# raise "Matrix inverse contains nan!" like the reference (costs one 4-byte D2H per call, where the
# reference pays a full GPU->CPU->GPU round trip); set False for fully asynchronous operation
CHECK_SINGULAR = True

# render_fused keeps one workspace per (device, stream, shape) between calls so that the plane-spacing table is
# computed once per (K_src^-1, disparities) (MPIFLOW_KEEP_TABLE); False: a fresh workspace per call, as in round 1
KEEP_WORKSPACES = True
_workspaces = {}


def release_workspaces():
    """drop the cached workspaces (one S*H*W-float table each)"""
    _workspaces.clear()


def _flags(is_bg_depth_inf=False, hard_flow=False):
    return (nat.BG_DEPTH_INF if is_bg_depth_inf else 0) | (nat.HARD_FLOW if hard_flow else 0)


def _dense(t, name):
    return nat.dev(t, name).contiguous()


# ---------------------------------------------------------------------------------------------
# compositing
# ---------------------------------------------------------------------------------------------
def render(rgb_BS3HW, sigma_BS1HW, xyz_BS3HW, src_sigma_BS1HW=None, src_flow_BS2HW=None,
           src_xyz_BS3HW=None, use_alpha=False, is_bg_depth_inf=False, hard_flow=False, obj_mask=None):
    """-> (imgs_syn, depth_syn, blend_weights, weights, flowA2B, obj_mask); reference :7-39.

    use_alpha=True raises UnboundLocalError in the reference (:33-39 never bind flowA2B); here it
    returns flowA2B=None and passes obj_mask through."""
    if use_alpha:
        imgs_syn, weights = alpha_composition(sigma_BS1HW, rgb_BS3HW)
        depth_syn, _ = alpha_composition(sigma_BS1HW, xyz_BS3HW[:, :, 2:])
        blend_weights = torch.cumprod(1 - sigma_BS1HW + 1e-6, dim=1)
        return imgs_syn, depth_syn, blend_weights, weights, None, obj_mask
    imgs_syn, depth_syn, blend_weights, weights, _, obj_mask = plane_volume_rendering(
        rgb_BS3HW, sigma_BS1HW, xyz_BS3HW, src_sigma_BS1HW, src_flow_BS2HW, src_xyz_BS3HW,
        is_bg_depth_inf, hard_flow=hard_flow, obj_mask=obj_mask)
    flowA2B = None
    if src_sigma_BS1HW is not None:
        flowA2B = plane_volume_rendering_flow(src_sigma_BS1HW, src_flow_BS2HW, src_xyz_BS3HW,
                                              is_bg_depth_inf, hard_flow=hard_flow)
    return imgs_syn, depth_syn, blend_weights, weights, flowA2B, obj_mask


def alpha_composition(alpha_BK1HW, value_BKCHW):
    """Front-to-back 'over' with given alphas -> (composed BxCxHxW, weights BxKx1xHxW); :42-59."""
    B, K, C, H, W = value_BKCHW.shape
    a, v = _dense(alpha_BK1HW, "alpha_BK1HW"), _dense(value_BKCHW, "value_BKCHW")
    comp = torch.empty((B, C, H, W), dtype=torch.float32, device=a.device)
    wts = torch.empty((B, K, 1, H, W), dtype=torch.float32, device=a.device)
    nat.check(nat.lib().mpiflow_alpha_composition_f32(nat.ptr(a), nat.ptr(v), B, K, C, H, W, nat.ptr(comp),
                                                      nat.ptr(wts), nat.stream(a)))
    return comp.to(value_BKCHW.dtype), wts.to(alpha_BK1HW.dtype)


def plane_volume_rendering(rgb_BS3HW, sigma_BS1HW, xyz_BS3HW, src_sigma_BS1HW, src_flow_BS2HW,
                           src_xyz_BS3HW, is_bg_depth_inf, hard_flow=False, obj_mask=None):
    """sigma -> alpha plane volume rendering in the frame of xyz.

    -> (rgb_out, depth_out, transparency_acc, weights, flow_out, obj_mask); reference :62-99.
    As there, flow_out/obj_mask are produced only when src_sigma_BS1HW is given (and obj_mask must
    then be a tensor: the reference fails on None at :95)."""
    B, S, _, H, W = sigma_BS1HW.shape
    rgb, sig, xyz = _dense(rgb_BS3HW, "rgb"), _dense(sigma_BS1HW, "sigma"), _dense(xyz_BS3HW, "xyz")
    with_flow = src_sigma_BS1HW is not None
    flow = obj = flow_out = obj_out = None
    new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=rgb.device)
    if with_flow:
        if obj_mask is None:
            raise TypeError("obj_mask must be a tensor when src_sigma_BS1HW is given (reference :95)")
        flow, obj = _dense(src_flow_BS2HW, "src_flow"), _dense(obj_mask.expand(B, S, 1, H, W), "obj_mask")
        flow_out, obj_out = new(B, 2, H, W), new(B, 1, H, W)
    rgb_out, depth_out, tacc, wts = new(B, 3, H, W), new(B, 1, H, W), new(B, S, 1, H, W), new(B, S, 1, H, W)
    nat.check(nat.lib().mpiflow_plane_volume_rendering_f32(
        nat.ptr(rgb), nat.ptr(sig), nat.ptr(xyz), nat.ptr(flow), nat.ptr(obj), B, S, H, W,
        _flags(is_bg_depth_inf), nat.ptr(rgb_out), nat.ptr(depth_out), nat.ptr(tacc), nat.ptr(wts),
        nat.ptr(flow_out), nat.ptr(obj_out), nat.stream(rgb)))
    dt = rgb_BS3HW.dtype
    cast = lambda t: None if t is None else t.to(dt)
    return cast(rgb_out), cast(depth_out), cast(tacc), cast(wts), cast(flow_out), (cast(obj_out) if with_flow else obj_mask)


def plane_volume_rendering_flow(src_sigma_BS1HW, src_flow_BS2HW, src_xyz_BS3HW, is_bg_depth_inf,
                                hard_flow=False):
    """Volume-render per-plane flow with the weights of the un-warped MPI -> Bx2xHxW; :102-139."""
    B, S, _, H, W = src_sigma_BS1HW.shape
    sig, flow, xyz = _dense(src_sigma_BS1HW, "src_sigma"), _dense(src_flow_BS2HW, "src_flow"), _dense(src_xyz_BS3HW, "src_xyz")
    out = torch.empty((B, 2, H, W), dtype=torch.float32, device=sig.device)
    nat.check(nat.lib().mpiflow_plane_volume_rendering_flow_f32(
        nat.ptr(sig), nat.ptr(flow), nat.ptr(xyz), B, S, H, W, _flags(hard_flow=hard_flow), nat.ptr(out),
        nat.stream(sig)))
    return out.to(src_flow_BS2HW.dtype)


def weighted_sum_mpi(rgb_BS3HW, xyz_BS3HW, weights, is_bg_depth_inf):
    """Blend planes with given weights -> (rgb Bx3xHxW, depth Bx1xHxW); reference :142-154.
    Only reached through plane_volume_rendering in the reference (fused into that kernel here);
    kept as a standalone helper for API compatibility."""
    wsum = weights.sum(dim=1)
    rgb = (weights * rgb_BS3HW).sum(dim=1)
    z = (weights * xyz_BS3HW[:, :, 2:]).sum(dim=1)
    depth = z + (1 - wsum) * 1000 if is_bg_depth_inf else z / (wsum + 1e-5)
    return rgb, depth


# ---------------------------------------------------------------------------------------------
# geometry of the planes
# ---------------------------------------------------------------------------------------------
def get_xyz_from_depth(meshgrid_homo, depth, K_inv):
    """Back-project a depth map: K^-1 (x,y,1) * depth -> Bx3xHxW; reference :157-177."""
    H, W = meshgrid_homo.shape[1:]
    B, _, H_d, W_d = depth.shape
    assert (H, W) == (H_d, W_d)
    rays = torch.matmul(K_inv, meshgrid_homo.reshape(1, 3, H * W).expand(B, 3, H * W))
    return rays.reshape(B, 3, H, W) * depth


def disparity_consistency_src_to_tgt(meshgrid_homo, K_src_inv, disparity_src, G_tgt_src, K_tgt, disparity_tgt):
    """Mean |1/z_tgt - disparity_tgt(projection)| over pixels projecting inside the image; :180-210."""
    B, _, H, W = disparity_src.shape
    xyz_src = get_xyz_from_depth(meshgrid_homo, torch.reciprocal(disparity_src), K_src_inv).reshape(B, 3, H * W)
    xyz_tgt = transform_G_xyz(G_tgt_src, xyz_src, is_return_homo=False)
    proj = torch.matmul(K_tgt, xyz_tgt)
    pxpy = proj[:, 0:2] / proj[:, 2:]
    inside = (pxpy[:, 0:1] >= 0) & (pxpy[:, 0:1] <= W - 1) & (pxpy[:, 1:2] >= 0) & (pxpy[:, 1:2] <= H - 1)
    diff = torch.abs(torch.reciprocal(xyz_tgt[:, 2:]) - gather_pixel_by_pxpy(disparity_tgt, pxpy))
    return torch.mean(diff[inside])


class _PlaneTag:
    """Provenance of a plane-derived xyz tensor, so render_tgt_rgb_depth can fuse without reading it.  The tag also
    remembers the version counters of the tensor it is attached to and of the tensors it was derived from: an in-place
    edit of any of them after the fact (xyz[...] += ..., disparity.mul_(...)) invalidates it, and the call takes the
    un-fused operators, which read the edited values like the reference would."""
    __slots__ = ("kind", "disparity", "K_inv", "G", "src", "versions")

    def __init__(self, kind, disparity=None, K_inv=None, G=None, src=None):
        self.kind, self.disparity, self.K_inv, self.G, self.src = kind, disparity, K_inv, G, src
        self.versions = ()

    def seal(self, xyz=None):
        """xyz: the materialised tensor the tag is attached to (a deferred one drops its tag itself when edited)"""
        self.versions = tuple((t, t._version) for t in (xyz, self.disparity, self.K_inv, self.G)
                              if t is not None and not isinstance(t, LazyTensor))
        return self

    def intact(self):
        return all(t._version == v for t, v in self.versions)


def _same(a, b):
    """the same tensor (object or view of the same memory) -- or, when only the dtype differs, a converted copy of it:
    the generator hands `G.to(K.dtype)` to get_tgt_xyz_from_plane_disparity and `G` itself to render_tgt_rgb_depth
    (utils/utils.py:308, :331), and in its fp16 mode those are two tensors holding one pose"""
    if a is b:
        return True
    if not (isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor)) or a.shape != b.shape or a.device != b.device:
        return False
    if a.dtype != b.dtype:
        return True
    return a.data_ptr() == b.data_ptr() and a.stride() == b.stride()


def _src_xyz_now(mpi_disparity_src, K_src_inv, H, W):
    B, S = mpi_disparity_src.shape
    disp, kinv = _dense(mpi_disparity_src, "mpi_disparity_src"), _dense(K_src_inv, "K_src_inv")
    xyz = torch.empty((B, S, 3, H, W), dtype=torch.float32, device=disp.device)
    nat.check(nat.lib().mpiflow_src_xyz_f32(nat.ptr(kinv), nat.ptr(disp), B, S, H, W, nat.ptr(xyz), nat.stream(xyz)))
    return xyz.to(K_src_inv.dtype)


def _tgt_xyz_now(xyz_src_BS3HW, G_tgt_src):
    B, S, _, H, W = xyz_src_BS3HW.shape
    G = G_tgt_src if G_tgt_src.dim() == 3 else G_tgt_src.unsqueeze(0)
    src, g = _dense(xyz_src_BS3HW, "xyz_src_BS3HW"), _dense(G.expand(B, 4, 4), "G_tgt_src")
    out = torch.empty_like(src)
    nat.check(nat.lib().mpiflow_tgt_xyz_f32(nat.ptr(src), nat.ptr(g), B, S, H, W, nat.ptr(out), nat.stream(src)))
    return out.to(xyz_src_BS3HW.dtype)


def get_src_xyz_from_plane_disparity(meshgrid_src_homo, mpi_disparity_src, K_src_inv):
    """xyz of every plane pixel in the source camera frame -> BxSx3xHxW; reference :213-239.
    Deferred: the kernel runs when (if) the values are read; render_tgt_rgb_depth does not read them."""
    B, S = mpi_disparity_src.shape
    H, W = meshgrid_src_homo.shape[1:]
    nat.dev(mpi_disparity_src, "mpi_disparity_src"), nat.dev(K_src_inv, "K_src_inv")       # CUDA-only, checked now
    tag = _PlaneTag("src", disparity=mpi_disparity_src, K_inv=K_src_inv).seal()
    return LazyTensor(lambda: _src_xyz_now(mpi_disparity_src, K_src_inv, H, W), (B, S, 3, H, W), K_src_inv.dtype,
                      mpi_disparity_src.device, plane=tag)


def get_tgt_xyz_from_plane_disparity(xyz_src_BS3HW, G_tgt_src):
    """The same points in the target camera frame (G . xyz) -> BxSx3xHxW; reference :242-256.
    Deferred when xyz_src is the (unedited) deferred tensor of get_src_xyz_from_plane_disparity."""
    src_tag = getattr(xyz_src_BS3HW, "_mpiflow_plane", None)
    if isinstance(xyz_src_BS3HW, LazyTensor) and src_tag is not None and src_tag.kind == "src":
        nat.dev(G_tgt_src, "G_tgt_src")
        tag = _PlaneTag("tgt", G=G_tgt_src, src=src_tag).seal()
        return LazyTensor(lambda: _tgt_xyz_now(xyz_src_BS3HW.materialize(), G_tgt_src), xyz_src_BS3HW.shape,
                          xyz_src_BS3HW.dtype, xyz_src_BS3HW.device, plane=tag)
    return _tgt_xyz_now(xyz_src_BS3HW, G_tgt_src)


# ---------------------------------------------------------------------------------------------
# the hot call
# ---------------------------------------------------------------------------------------------
def _fusable(mpi_disparity_src, xyz_tgt, xyz_src, G_tgt_src, K_src_inv):
    ts, tt = getattr(xyz_src, "_mpiflow_plane", None), getattr(xyz_tgt, "_mpiflow_plane", None)
    if getattr(xyz_src, "edited", False) or getattr(xyz_tgt, "edited", False):
        return False                       # deferred tensors whose computed values were modified afterwards
    return (ts is not None and tt is not None and ts.kind == "src" and tt.kind == "tgt" and tt.src is ts
            and ts.intact() and tt.intact()
            and _same(ts.disparity, mpi_disparity_src) and _same(ts.K_inv, K_src_inv) and _same(tt.G, G_tgt_src))


def render_fused(mpi_rgb_src, mpi_sigma_src, mpi_disparity_src, G_tgt_src, K_src_inv, K_tgt, obj_mask=None,
                 is_bg_depth_inf=False, hard_flow=False, out=None, poses_per_mpi=1):
    """The fused kernel on its own inputs: rgb BxSx3xHxW, sigma BxSx1xHxW (views are fine), disparity
    BxS, G Bx4x4, K_src_inv/K_tgt Bx3x3, obj_mask BxSx1xHxW or anything expandable to it.
    poses_per_mpi = P > 1: rgb / sigma hold B/P MPIs, frames b*P .. b*P+P-1 render MPI b (everything else is per frame).

    -> (rgb Bx3xHxW, depth Bx1xHxW, mask Bx1xHxW, flowA2B Bx2xHxW, obj Bx1xHxW or None), fp32."""
    B, S, _, H, W = mpi_rgb_src.shape
    if not 1 <= poses_per_mpi <= 256:
        raise ValueError("poses_per_mpi must be in 1..256")
    B *= poses_per_mpi
    # an fp16 MPI is read as fp16 (mpiflow_render_fused_f16: fp32 arithmetic on the converted values), never up-cast
    half = mpi_rgb_src.dtype == torch.float16 and mpi_sigma_src.dtype == torch.float16 and poses_per_mpi == 1
    if half:
        rgb, sig = nat.dev_storage(mpi_rgb_src, "mpi_rgb_src"), nat.dev_storage(mpi_sigma_src, "mpi_sigma_src")
    else:
        rgb, sig = nat.dev(mpi_rgb_src, "mpi_rgb_src"), nat.dev(mpi_sigma_src, "mpi_sigma_src")
    if not nat.inner_contiguous(rgb):
        rgb = rgb.contiguous()
    if not nat.inner_contiguous(sig):
        sig = sig.contiguous()
    device = rgb.device
    # per-MPI arguments are repeated for the poses of each MPI; per-frame arguments pass through
    per_frame = lambda t: t.repeat_interleave(poses_per_mpi, 0) if (poses_per_mpi > 1 and t.shape[0] * poses_per_mpi == B) else t
    G = G_tgt_src if G_tgt_src.dim() == 3 else G_tgt_src.unsqueeze(0)
    G = _dense(per_frame(G.to(device)).expand(B, 4, 4), "G_tgt_src")
    disp = _dense(per_frame(mpi_disparity_src.to(device)).expand(B, S), "mpi_disparity_src")
    kinv = _dense(per_frame(K_src_inv.to(device)).expand(B, 3, 3), "K_src_inv")
    kt = _dense(per_frame(K_tgt.to(device)).expand(B, 3, 3), "K_tgt")
    obj = None
    obj_sb = obj_ss = 0
    if obj_mask is not None:
        obj = nat.dev(obj_mask.to(device), "obj_mask")
        if half:
            obj = obj.to(torch.float16)       # B x 1 x H x W: the entry point reads one storage type
        while obj.dim() < 5:
            obj = obj.unsqueeze(1)
        obj = obj.expand(B, S, 1, H, W)
        if not nat.inner_contiguous(obj):
            obj = obj.contiguous()
        obj_sb, obj_ss = obj.stride(0), obj.stride(1)
    if out is None:
        new = lambda c: torch.empty((B, c, H, W), dtype=torch.float32, device=device)
        out = (new(3), new(1), new(1), new(2), new(1) if obj is not None else None)
    o_rgb, o_depth, o_mask, o_flow, o_obj = out
    L = nat.lib()
    # One workspace per (device, stream, shape), kept between calls: the plane-spacing table in it is then computed once
    # per (K_src^-1, disparities) instead of once per call (MPIFLOW_KEEP_TABLE; the library checks the key on the
    # device), which is what a generator loop with fixed intrinsics wants -- single frames and pose pairs included.
    keep = 4 if (KEEP_WORKSPACES and W % 4 == 0 and not half) else 0
    if keep:
        key = (device.index, torch.cuda.current_stream(device).cuda_stream, B, S, H, W)
        ws = _workspaces.get(key)
        if ws is None:
            if len(_workspaces) >= 8:
                _workspaces.clear()
            # zeroed once: recycled allocator memory may hold the key of an earlier table whose data has been overwritten
            ws = _workspaces[key] = torch.zeros(L.mpiflow_render_workspace_bytes_hw(B, S, H, W), dtype=torch.uint8, device=device)
        ws_bytes = ws.numel()
    else:
        ws_bytes = L.mpiflow_render_workspace_bytes_hw(B, S, H, W) if B >= 4 else L.mpiflow_render_workspace_bytes(B, S)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    status = torch.zeros(1, dtype=torch.int32, device=device)
    entry = L.mpiflow_render_fused_f16 if half else L.mpiflow_render_fused_f32
    nat.check(entry(
        nat.ptr(rgb), rgb.stride(0), rgb.stride(1), rgb.stride(2), nat.ptr(sig), sig.stride(0), sig.stride(1),
        nat.ptr(obj), obj_sb, obj_ss, nat.ptr(disp), nat.ptr(G), nat.ptr(kinv), nat.ptr(kt), B, S, H, W,
        _flags(is_bg_depth_inf, hard_flow) | ((poses_per_mpi - 1) << 8) | keep, nat.ptr(o_rgb), nat.ptr(o_depth), nat.ptr(o_mask), nat.ptr(o_flow),
        nat.ptr(o_obj), nat.ptr(ws), ws_bytes, nat.ptr(status), nat.stream(rgb)))
    if CHECK_SINGULAR and int(status.item()):
        raise Exception("Matrix inverse contains nan!")
    return o_rgb, o_depth, o_mask, o_flow, o_obj


def render_tgt_rgb_depth(H_sampler: HomographySample, mpi_rgb_src, mpi_sigma_src, mpi_disparity_src,
                         xyz_tgt_BS3HW, xyz_src_BS3HW, G_tgt_src, K_src_inv, K_tgt, mpi_flow_src=None,
                         use_alpha=False, is_bg_depth_inf=False, hard_flow=False, obj_mask=None):
    """Novel view + flow of an MPI -> (tgt_rgb Bx3xHxW, tgt_depth Bx1xHxW, tgt_mask Bx1xHxW (number of
    planes whose warp lands inside the source image), flowA2B Bx2xHxW (source frame), obj_mask
    Bx1xHxW); reference :259-349.  mpi_flow_src is accepted and ignored, as in the reference."""
    B, S, _, H, W = mpi_rgb_src.shape
    dt = mpi_rgb_src.dtype
    if not use_alpha and _fusable(mpi_disparity_src, xyz_tgt_BS3HW, xyz_src_BS3HW, G_tgt_src, K_src_inv):
        outs = render_fused(mpi_rgb_src, mpi_sigma_src, mpi_disparity_src, G_tgt_src, K_src_inv, K_tgt,
                            obj_mask, is_bg_depth_inf, hard_flow)
        return tuple(None if o is None else o.to(dt) for o in outs)

    # arbitrary xyz tensors (or alpha mode): the reference's own sequence on the un-fused operators
    depth = torch.reciprocal(mpi_disparity_src).reshape(B * S)
    planes = [mpi_rgb_src, mpi_sigma_src, xyz_tgt_BS3HW] + ([obj_mask] if obj_mask is not None else [])
    stacked = torch.cat([nat.dev(t).expand(B, S, t.shape[2], H, W) for t in planes], dim=2)
    C = stacked.shape[2]
    G = G_tgt_src if G_tgt_src.dim() == 3 else G_tgt_src.unsqueeze(0)
    per_plane = lambda m: m.unsqueeze(1).expand(B, S, *m.shape[1:]).reshape(B * S, *m.shape[1:])
    args = (depth, per_plane(G.expand(B, 4, 4)), per_plane(K_src_inv), per_plane(K_tgt))
    warped, valid, _ = H_sampler.sample(stacked.reshape(B * S, C, H, W), *args)
    flow = H_sampler.sample_inverse(stacked.reshape(B * S, C, H, W), *args)
    flow = flow.permute(0, 3, 1, 2).reshape(B, S, 2, H, W)
    warped = warped.reshape(B, S, C, H, W)
    t_rgb, t_sigma, t_xyz = warped[:, :, 0:3], warped[:, :, 3:4], warped[:, :, 4:7]
    t_obj = warped[:, :, 7:] if obj_mask is not None else None
    t_sigma = torch.where(t_xyz[:, :, 2:3] >= 0, t_sigma, torch.zeros_like(t_sigma))
    rgb, dep, _, _, flowA2B, obj_syn = render(t_rgb, t_sigma, t_xyz, mpi_sigma_src, flow, xyz_src_BS3HW,
                                              use_alpha=use_alpha, is_bg_depth_inf=is_bg_depth_inf,
                                              hard_flow=hard_flow, obj_mask=t_obj)
    mask = valid.reshape(B, S, H, W).to(torch.float32).sum(dim=1, keepdim=True)
    return rgb, dep, mask, flowA2B, obj_syn
