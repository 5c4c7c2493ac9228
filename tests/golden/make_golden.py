"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference,
CPU, fp32) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py            [--warpback-only | --flow-io-only | --sample-only | --moving-obj-only]

The reference hard-codes CUDA in two places (utils/mpi/homography_sampler.py:16-17); for a CPU run
they are shimmed: torch.cuda.synchronize -> no-op, Tensor.cuda -> identity.  Nothing of the
reference is copied: only its outputs are stored.  The splat fixtures come from the reference's
warping.c compiled from source by oracle/Makefile (oracle/_ref/libwarping_ref.so).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MPIFLOW_REFERENCE", "/root/reference")

torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.cuda = lambda self, *a, **k: self
sys.path.insert(0, REF)
from utils.mpi import mpi_rendering as ref_render          # noqa: E402
from utils.mpi.homography_sampler import HomographySample  # noqa: E402
import geometry as ref_geometry                            # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "mpi-flow_b200"))
sys.path.insert(0, ROOT)
import synth                                               # noqa: E402
from oracle import mpi_oracle                              # noqa: E402

torch.set_num_threads(os.cpu_count())
T = torch.from_numpy


def run_render(inp, keep_inputs, **flags):
    """reference xyz_src + xyz_tgt + render_tgt_rgb_depth on a synth.frame_inputs dict"""
    mpi, disp, G, Kinv, K = (T(inp[k]) for k in ("mpi", "disparity", "G", "Kinv", "K"))
    B, S, _, H, W = mpi.shape
    sampler = HomographySample(H, W, torch.device("cpu"))
    xyz_src = ref_render.get_src_xyz_from_plane_disparity(sampler.meshgrid, disp, Kinv)
    xyz_tgt = ref_render.get_tgt_xyz_from_plane_disparity(xyz_src, G)
    obj = T(inp["obj_mask"]).unsqueeze(1).repeat(1, S, 1, 1, 1)
    rgb, depth, mask, flow, obj_out = ref_render.render_tgt_rgb_depth(
        sampler, mpi[:, :, 0:3], mpi[:, :, 3:], disp, xyz_tgt, xyz_src, G, Kinv, K,
        None, use_alpha=False, obj_mask=obj, **flags)
    out = dict(out_rgb=rgb, out_depth=depth, out_mask=mask, out_flow=flow, out_obj=obj_out)
    out = {k: v.numpy() for k, v in out.items()}
    if keep_inputs:
        out.update({"in_" + k: v for k, v in inp.items()})
    return out, (sampler, xyz_src, xyz_tgt)


def load_reference_warpback():
    """warpback/utils.py imports pytorch3d (absent here) at module level; its mesh-construction helpers are
    plain torch.  Import the unmodified file with empty stand-ins for the three pytorch3d modules."""
    import importlib.util
    import types
    for name, attrs in (("pytorch3d", ()), ("pytorch3d.renderer", ()), ("pytorch3d.renderer.mesh", ("rasterize_meshes",)),
                        ("pytorch3d.structures", ("Meshes",)), ("pytorch3d.ops", ("interpolate_face_attributes",))):
        m = sys.modules.setdefault(name, types.ModuleType(name))
        for a in attrs:
            setattr(m, a, None)
    spec = importlib.util.spec_from_file_location("ref_warpback_utils", os.path.join(REF, "warpback", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def warpback_fixture():
    ref = load_reference_warpback()
    g = torch.Generator().manual_seed(synth.SEED)
    b, h, w = 2, 12, 20
    rgbd = torch.rand((b, 4, h, w), generator=g)
    rgbd[:, 3, :, w // 2:] *= 0.3                     # a depth discontinuity for the visibility mask
    cam_int = torch.tensor([[0.58, 0, 0.5], [0, 0.58, 0.5], [0, 0, 1]]).repeat(b, 1, 1)
    cam_int[1, 0, 0] = 0.7
    r = ref.RGBDRenderer(torch.device("cpu"))
    mesh = r.construct_mesh(rgbd, cam_int)
    aa = torch.tensor([[[0.01, -0.02, 0.005]], [[0.0, 0.03, -0.01]]])
    tt = torch.tensor([[[0.1, -0.05, -0.15]], [[-0.2, 0.0, 0.05]]])
    save("warpback", rgbd=rgbd.numpy(), cam_int=cam_int.numpy(), vertice=mesh["vertice"].numpy(),
         faces=mesh["faces"].numpy(), attributes=mesh["attributes"].numpy(), size=np.array(mesh["size"]),
         pixel=r.get_screen_pixel_coord(h, w).numpy(), persp=r.get_perspective_from_intrinsic(cam_int).numpy(),
         vis=r.get_visible_mask(rgbd.permute(0, 2, 3, 1)[..., -1:]).numpy(), aa=aa.numpy(), tt=tt.numpy(),
         pose=ref.transformation_from_parameters(aa, tt).numpy(),
         pose_inv=ref.transformation_from_parameters(aa, tt, invert=True).numpy())


def flow_io_fixture():
    """bytes written by the reference's write_flow.py for a seeded flow field"""
    import importlib.util
    import tempfile
    spec = importlib.util.spec_from_file_location("ref_write_flow", os.path.join(REF, "write_flow.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(synth.SEED)
    flow = ((rng.random((11, 17, 2)) - 0.5) * 300).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        ref.writeFlow(os.path.join(d, "a.flo"), flow)
        ref.writeFlowKITTI(os.path.join(d, "a.png"), flow)
        flo = np.fromfile(os.path.join(d, "a.flo"), np.uint8)
        png = np.fromfile(os.path.join(d, "a.png"), np.uint8)
        kf, kv = ref.readFlowKITTI(os.path.join(d, "a.png"))
        back = ref.readFlow(os.path.join(d, "a.flo"))
    save("flow_io", flow=flow, flo_bytes=flo, kitti_png_bytes=png, kitti_flow=kf, kitti_valid=kv, flo_back=back)


def load_reference_utils():
    """utils/utils.py imports moving_obj, which loads a prebuilt external/forward_warping/libwarping.so relative to
    the working directory at import time (moving_obj.py:12).  The sample-assembly functions never call it, so the
    unmodified file is imported with an empty stand-in for that one module."""
    import importlib
    import types
    if "moving_obj" not in sys.modules:
        m = types.ModuleType("moving_obj")
        m.moveing_object_with_mask = None
        sys.modules["moving_obj"] = m
    return importlib.import_module("utils.utils")


def sample_fixture():
    """render_3dphoto_dynamic (utils/utils.py:159-288) on one small seeded sample, with the intermediates its two
    render_novel_view_dynamic calls and its cv2.inpaint call see."""
    import random
    import types
    ru = load_reference_utils()
    B, S, H, W, seed = 1, 6, 40, 64, 114530
    inp = synth.frame_inputs(B, S, H, W, "smooth", "random", seed)
    mpi = inp["mpi"].copy()
    mpi[:, :, 3] *= 4.0                       # opaque enough that both threshold branches occur
    rng = np.random.default_rng(seed)
    src = rng.random((B, 3, H, W), dtype=np.float32)
    obj = (inp["obj_mask"] > 0.45).astype(np.float32) * 0.995 + inp["obj_mask"] * 0.005   # mostly hard, soft rim
    calls, inpaint_args = [], []
    real_rnvd, real_inpaint = ru.render_novel_view_dynamic, ru.cv2.inpaint

    def spy_rnvd(obj_mask, rgb, sigma, disp, G, *rest, **kw):
        out = real_rnvd(obj_mask, rgb, sigma, disp, G, *rest, **kw)
        calls.append(dict(G=G.numpy().copy(), blended=rgb.numpy().copy(), frame=out[0].numpy(), flow=out[2].numpy(),
                          mask=out[3].numpy()))
        return out

    def spy_inpaint(img, mask, *a):
        inpaint_args.append((img.copy(), mask.copy()))
        return real_inpaint(img, mask, *a)

    ru.render_novel_view_dynamic = spy_rnvd
    ru.cv2 = types.SimpleNamespace(inpaint=spy_inpaint, INPAINT_NS=ru.cv2.INPAINT_NS)
    random.seed(seed)
    opt = types.SimpleNamespace(ext_cz=0.15)
    flow_mix, src_np, inpainted, _ = ru.render_3dphoto_dynamic(
        opt, T(src), T(obj), None, T(mpi), T(inp["disparity"]), T(inp["K"]), T(inp["K"]), name="sample.png")
    assert len(calls) == 2 and len(inpaint_args) == 1
    save("sample_pair", meta=np.array([B, S, H, W, seed], np.int64), ext_cz=0.15, mask_thresh=0.99,
         in_mpi=mpi, in_src=src, in_obj=obj, in_disparity=inp["disparity"], in_K=inp["K"],
         G_obj=calls[0]["G"], G_cam=calls[1]["G"], blended=calls[0]["blended"],
         frame=calls[0]["frame"], frame_dync=calls[1]["frame"], flow=calls[0]["flow"], flow_dync=calls[1]["flow"],
         mask=calls[0]["mask"], mask_dync=calls[1]["mask"],
         frame_mix=inpaint_args[0][0], fill_mask=inpaint_args[0][1],
         flow_mix=flow_mix, src_np=src_np, inpainted=inpainted)
    print("holes:", int(inpaint_args[0][1].sum()), "of", H * W, " own:", int((calls[0]["mask"] >= 0.99).sum()),
          " obj>=t:", int((obj >= 0.99).sum()))


def moving_obj_fixture():
    """moveing_object_with_mask (moving_obj.py:16-175) on one small seeded case.  The module loads
    external/forward_warping/libwarping.so relative to the working directory at import time (:12); it is imported
    from a scratch directory where that path is the reference's warping.c built from source (oracle/_ref).  Its
    call of the splat is wrapped to record the index maps, depth and result it passes by pointer (:127-129)."""
    import ctypes
    import importlib.util
    import random
    import tempfile
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libwarping_ref.so")
    assert os.path.exists(ref_so), "run `make -C oracle ref` first"
    sys.modules.pop("moving_obj", None)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as scratch:
        os.makedirs(os.path.join(scratch, "external", "forward_warping"))
        os.makedirs(os.path.join(scratch, "temp"))
        os.symlink(ref_so, os.path.join(scratch, "external", "forward_warping", "libwarping.so"))
        os.chdir(scratch)
        try:
            # by path: the drop-in of the same name in mpi-flow_b200/ is ahead of the reference on sys.path here
            spec = importlib.util.spec_from_file_location("moving_obj", os.path.join(REF, "moving_obj.py"))
            mo = importlib.util.module_from_spec(spec)
            assert sys.modules["geometry"] is ref_geometry
            spec.loader.exec_module(mo)
            h, w, seed = 36, 60, 114540
            rng = np.random.default_rng(seed)
            yy, xx = np.mgrid[0:h, 0:w]
            disp = (0.15 + 0.5 * (yy / h) + 0.2 * np.exp(-((xx - 25) ** 2 + (yy - 18) ** 2) / 60.0)).astype(np.float32)
            disp += (rng.random((h, w), dtype=np.float32) - 0.5) * 0.02
            inst = (((xx - 25) ** 2 + (yy - 18) ** 2) < 81).astype(np.float32)
            rgb = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
            K = synth.intrinsics(h, w)
            Kinv = synth.intrinsics_inv(K)
            rec = {}
            real = mo.warp

            def spy(src, idx, idy, z, warped, hh, ww):
                n = hh.value * ww.value
                rec["safe_x"] = np.ctypeslib.as_array((ctypes.c_int64 * n).from_address(idx.value)).copy()
                rec["safe_y"] = np.ctypeslib.as_array((ctypes.c_int64 * n).from_address(idy.value)).copy()
                rec["z"] = np.ctypeslib.as_array((ctypes.c_float * n).from_address(z.value)).copy()
                rec["img"] = np.ctypeslib.as_array((ctypes.c_uint8 * (3 * n)).from_address(src.value)).copy()
                real(src, idx, idy, z, warped, hh, ww)
                rec["warped"] = np.ctypeslib.as_array((ctypes.c_uint8 * (5 * n)).from_address(warped.value)).copy()

            mo.warp = spy
            random.seed(seed)
            mo.moveing_object_with_mask(None, T(disp)[None, None], rgb, T(K), T(Kinv), T(inst)[None, None], 0)
            rec["next_random"] = random.random()      # where the reference leaves the `random` stream
        finally:
            os.chdir(cwd)
            sys.modules.pop("moving_obj", None)
    save("moving_obj", meta=np.array([h, w, seed], np.int64), in_disp=disp, in_inst=inst, in_rgb=rgb, in_K=K, in_Kinv=Kinv,
         safe_x=rec["safe_x"].reshape(h, w), safe_y=rec["safe_y"].reshape(h, w), z=rec["z"].reshape(h, w),
         img=rec["img"].reshape(h, w, 3), warped=rec["warped"].reshape(h, w, 5), next_random=rec["next_random"])
    moved = (rec["safe_x"].reshape(h, w) != xx) | (rec["safe_y"].reshape(h, w) != yy)
    print("moved pixels:", int(moved.sum()), " valid:", int(rec["warped"].reshape(h, w, 5)[..., 3].sum()),
          " collisions flagged:", int(rec["warped"].reshape(h, w, 5)[..., 4].sum()))


def save(name, **arrs):
    p = os.path.join(HERE, name + ".npz")
    np.savez_compressed(p, **arrs)
    print(f"{name}: {os.path.getsize(p) / 1024:.0f} KiB")


def main():
    if "--warpback-only" in sys.argv:
        warpback_fixture()
        return
    if "--moving-obj-only" in sys.argv:
        moving_obj_fixture()
        return
    if "--sample-only" in sys.argv:
        sample_fixture()
        return
    if "--flow-io-only" in sys.argv:
        flow_io_fixture()
        return
    # ---- whole-frame cases (inputs regenerable from synth + the stored meta) ------------------
    cases = {
        # name: (B, S, H, W, kind, pose, seed, flags)
        "frame_white_random": (2, 6, 40, 56, "white", "random", 114514, {}),
        "frame_smooth_regression": (1, 8, 48, 64, "smooth", "regression", 114515, {}),
        "frame_identity": (1, 5, 32, 48, "white", "identity", 114516, {}),
        "frame_object_pose_flags": (1, 6, 36, 52, "smooth", "object", 114517,
                                    dict(is_bg_depth_inf=True, hard_flow=True)),
        "frame_wide_1280": (1, 4, 16, 1280, "white", "random", 114518, {}),
        "frame_c1_256": (1, 32, 256, 256, "smooth", "identity", 114519, {}),
    }
    for name, (B, S, H, W, kind, pose, seed, flags) in cases.items():
        inp = synth.frame_inputs(B, S, H, W, kind, pose, seed)
        small = inp["mpi"].nbytes < (1 << 20)
        out, _ = run_render(inp, keep_inputs=small, **flags)
        meta = np.array([B, S, H, W, seed], np.int64)
        save(name, meta=meta, kind=kind, pose=pose, is_bg_depth_inf=bool(flags.get("is_bg_depth_inf", False)),
             hard_flow=bool(flags.get("hard_flow", False)), **out)

    # ---- per-function intermediates on one tiny case -------------------------------------------
    inp = synth.frame_inputs(2, 5, 24, 40, "white", "random", 114520)
    mpi, disp, G, Kinv, K = (T(inp[k]) for k in ("mpi", "disparity", "G", "Kinv", "K"))
    B, S, _, H, W = mpi.shape
    _, (sampler, xyz_src, xyz_tgt) = run_render(inp, False)
    rep = lambda m: m.unsqueeze(1).repeat(1, S, 1, 1).reshape(B * S, *m.shape[1:])
    depth = torch.reciprocal(disp).view(B * S)
    src = torch.cat((mpi, xyz_tgt), dim=2).view(B * S, 7, H, W)
    warped, valid, flowB2A = sampler.sample(src, depth, rep(G), rep(Kinv), rep(K))
    flowA2B = sampler.sample_inverse(src, depth, rep(G), rep(Kinv), rep(K))
    w5 = warped.view(B, S, 7, H, W)
    pvr = ref_render.plane_volume_rendering(
        w5[:, :, 0:3], w5[:, :, 3:4], w5[:, :, 4:7], mpi[:, :, 3:],
        flowA2B.permute(0, 3, 1, 2).reshape(B, S, 2, H, W), xyz_src, False,
        obj_mask=T(inp["obj_mask"]).unsqueeze(1).repeat(1, S, 1, 1, 1))
    pvr_bg = ref_render.plane_volume_rendering(w5[:, :, 0:3], w5[:, :, 3:4], w5[:, :, 4:7], None, None, None, True)
    flow_soft = ref_render.plane_volume_rendering_flow(
        mpi[:, :, 3:], flowA2B.permute(0, 3, 1, 2).reshape(B, S, 2, H, W), xyz_src, False, hard_flow=False)
    flow_hard = ref_render.plane_volume_rendering_flow(
        mpi[:, :, 3:], flowA2B.permute(0, 3, 1, 2).reshape(B, S, 2, H, W), xyz_src, False, hard_flow=True)
    alpha = torch.clamp(mpi[:, :, 3:], 0, 1)
    ac_rgb, ac_w = ref_render.alpha_composition(alpha, mpi[:, :, 0:3])
    # NOTE: ref_render.render(..., use_alpha=True) raises UnboundLocalError in the reference
    # (utils/mpi/mpi_rendering.py:33-39 never binds flowA2B), so that branch has no golden.
    save("functions_tiny", meta=np.array([B, S, H, W, 114520], np.int64),
         **{"in_" + k: v for k, v in inp.items()},
         meshgrid=sampler.meshgrid.numpy(), xyz_src=xyz_src.numpy(), xyz_tgt=xyz_tgt.numpy(),
         warped=warped.numpy(), valid=valid.numpy(), flowB2A=flowB2A.numpy(), flowA2B=flowA2B.numpy(),
         pvr_rgb=pvr[0].numpy(), pvr_depth=pvr[1].numpy(), pvr_tacc=pvr[2].numpy(), pvr_weights=pvr[3].numpy(),
         pvr_flow=pvr[4].numpy(), pvr_obj=pvr[5].numpy(), pvr_bg_depth=pvr_bg[1].numpy(),
         flow_soft=flow_soft.numpy(), flow_hard=flow_hard.numpy(),
         ac_rgb=ac_rgb.numpy(), ac_w=ac_w.numpy())

    # ---- geometry.py ------------------------------------------------------------------------------
    rng = np.random.default_rng(7)
    aa = (rng.standard_normal((4, 1, 3)) * 0.05).astype(np.float32)
    aa[3] = 0.0                                                   # zero rotation (moving_obj.py:94)
    tr = (rng.standard_normal((4, 1, 3)) * 0.2).astype(np.float32)
    M = ref_geometry.transformation_from_parameters(T(aa), T(tr)).numpy()
    Mi = ref_geometry.transformation_from_parameters(T(aa), T(tr), invert=True).numpy()
    h, w = 20, 28
    K4 = np.eye(4, dtype=np.float32)[None]; K4[0, :3, :3] = synth.intrinsics(h, w)[0]
    K4i = np.eye(4, dtype=np.float32)[None]; K4i[0, :3, :3] = synth.intrinsics_inv(synth.intrinsics(h, w))[0]
    dep = (rng.random((1, 1, h, w)) * 5 + 0.5).astype(np.float32)
    cam = ref_geometry.BackprojectDepth(1, h, w)(T(dep), T(K4i))
    pix, z = ref_geometry.Project3D(1, h, w)(cam, T(K4), T(M[:1]))
    save("geometry", axisangle=aa, translation=tr, M=M, M_invert=Mi, K4=K4, K4inv=K4i, depth=dep,
         cam_points=cam.numpy(), pix_coords=pix.numpy(), z=z.numpy())

    # ---- forward splat (reference warping.c compiled from source) ----------------------------------
    splat = {}
    for tag, (h, w, mode, seed) in {"dense": (48, 64, "dense", 1), "pile": (40, 40, "pile", 2),
                                    "identity": (16, 24, "identity", 3), "row": (1, 97, "dense", 4),
                                    "kitti_strip": (24, 1280, "dense", 5)}.items():
        src, idx, idy, z = synth.splat_inputs(h, w, seed, mode)
        out = mpi_oracle.forward_warping(src, idx, idy, z, use_ref=True)
        splat.update({f"{tag}_meta": np.array([h, w, seed], np.int64), f"{tag}_mode": mode, f"{tag}_warped": out})
    save("splat", **splat)
    warpback_fixture()
    flow_io_fixture()
    sample_fixture()
    moving_obj_fixture()


if __name__ == "__main__":
    main()
