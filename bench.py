#!/usr/bin/env python
"""bench.py -- MPI frames/sec (64 planes, 1280x384) on N B200s; HBM GB/s vs roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5-16|c5-32|c5-64|c5-128]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the fused renderer (render_tgt_rgb_depth-equivalent: homography warp of
every plane + over-composite + flow projection) over one batch of synthetic frames of BASELINE.json
config 2 (1x64x384x1280 per frame, random camera pose).  Prints ONE JSON line (rank 0):
  value    frames/s, inputs resident in HBM, CUDA-event timed on the launching stream, max over ranks
  e2e      frames/s through the C-ABI host entry point (mpiflow_render_fused_host): pinned HOST
           buffers in, HOST buffers out, H2D + kernels + D2H inside the timed region
  roofline algorithmic bytes/launch / measured launch time vs the measured HBM copy bandwidth
  cpu_baseline  the CPU oracle timed on this box's host cores on a bounded sample (rank 0, N=1)
--impl reference times the reference's own CPU implementation of the same path: the UNMODIFIED utils/mpi/*.py of
Sharpiless/MPI-Flow (byte copies under oracle/_ref/refsrc/, made by oracle/Makefile) on torch-CPU with all host
threads, one whole frame per step (kind "reference"); oracle/torch_ref.py, the same operator sequence restated, only
when those files are absent (kind "port").
--config selects another BASELINE.json configuration (default c2, the one the metric is quoted on): c3 adds one
forward splat per frame, c4 is the 256-frame 640x480 batch strong-scaled over the ranks, c5-S the 1920x1080 plane sweep.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "mpi-flow_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import synth  # noqa: E402

METRIC = "MPI frames/sec (64 planes, 1280x384)"
# BASELINE.json configurations: planes, height, width, frames per launch per GPU, description
WORKLOADS = {
    "c2": (64, 384, 1280, 16, "config 2: KITTI-shape 1280x384, 64-plane MPI, random camera pose (synthetic white-noise MPI)"),
    "c3": (64, 384, 1280, 16, "config 3: 1280x384, 64 planes + one forward_warping splat per frame (dynamic path)"),
    "c4": (64, 480, 640, 32, "config 4: batch of 256 COCO-shape 640x480 frames, 64 planes, strong-scaled over the ranks"),
    "c5-16": (16, 1080, 1920, 4, "config 5: 1920x1080, 16 planes"), "c5-32": (32, 1080, 1920, 4, "config 5: 1920x1080, 32 planes"),
    "c5-64": (64, 1080, 1920, 4, "config 5: 1920x1080, 64 planes"), "c5-128": (128, 1080, 1920, 2, "config 5: 1920x1080, 128 planes"),
}
S, H, W = WORKLOADS["c2"][:3]
WORKLOAD = WORKLOADS["c2"][4]
BYTES_PER_FRAME = 4 * H * W * (4 * S + 1 + 8)  # RGBA once + mask once + 8 output channels


def set_workload(name):
    """module-level shape of the run (tools/ import bench for the default, config 2)"""
    global S, H, W, WORKLOAD, BYTES_PER_FRAME
    S, H, W, frames, WORKLOAD = WORKLOADS[name]
    BYTES_PER_FRAME = 4 * H * W * (4 * S + 1 + 8)
    return frames


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (NVML in-process; one sample every few ms)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x1: "gpu_idle", 0x2: "app_clocks"}

    def __init__(self, device_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = torch.cuda.get_device_properties(device_index).uuid
            self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            self.nv = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in {**self.BAD, **self.NOTE}.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + self.err]}
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def make_device_batch(B, seed, device):
    """synthetic frames of config 2, generated on the device (SURVEY 8d statistics, white noise)"""
    gen = torch.Generator(device=device).manual_seed(seed)
    mpi = torch.empty((B, S, 4, H, W), dtype=torch.float32, device=device)
    mpi[:, :, 0:3].copy_(torch.rand((B, S, 3, H, W), device=device, generator=gen))
    mpi[:, :, 3:4].copy_(torch.relu(torch.randn((B, S, 1, H, W), device=device, generator=gen)) + 1e-4)
    K = synth.intrinsics(H, W, B)
    d = dict(mpi=mpi,
             disparity=torch.from_numpy(synth.plane_disparities(S, B)).to(device),
             G=torch.from_numpy(synth.poses(B, "random", seed)).to(device),
             K=torch.from_numpy(K).to(device), Kinv=torch.from_numpy(synth.intrinsics_inv(K)).to(device),
             obj_mask=torch.from_numpy(synth.obj_mask(B, H, W, seed)).to(device))
    return d


class FusedCall:
    """pre-bound C-ABI call (no per-step allocation): mpiflow_render_fused_f32 on the current stream"""

    def __init__(self, d, device, out=None, flags=0):
        import _native as nat
        self.nat, self.L = nat, nat.lib()
        B = d["mpi"].shape[0]
        self.B = B
        new = lambda c: torch.empty((B, c, H, W), dtype=torch.float32, device=device)
        self.out = tuple(out) if out is not None else (new(3), new(1), new(1), new(2), new(1))
        self.ws_bytes = self.L.mpiflow_render_workspace_bytes_hw(B, S, H, W)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.d = d
        mpi = d["mpi"]
        rgb, sig, obj = mpi[:, :, 0:3], mpi[:, :, 3:4], d["obj_mask"]
        p = nat.ptr
        self.args = [p(rgb), rgb.stride(0), rgb.stride(1), rgb.stride(2), p(sig), sig.stride(0), sig.stride(1),
                     p(obj), obj.stride(0), 0, p(d["disparity"]), p(d["G"]), p(d["Kinv"]), p(d["K"]), B, S, H, W, flags,
                     *[p(o) for o in self.out], p(self.ws), self.ws_bytes, p(self.status)]
        self.out_arg0 = 19

    def retarget(self, out):
        """render into other (preallocated) output tensors from now on"""
        self.out = tuple(out)
        self.args[self.out_arg0:self.out_arg0 + 5] = [self.nat.ptr(o) for o in self.out]

    def __call__(self):
        rc = self.L.mpiflow_render_fused_f32(*self.args, self.nat.stream(self.ws))
        if rc:
            self.nat.check(rc)


class SplatCall:
    """pre-bound mpiflow_forward_warp_f32 on one frame's index maps (config 3: moving_obj.py:108-129 per frame)"""

    def __init__(self, device, seed=7):
        import _native as nat
        self.nat, self.L = nat, nat.lib()
        src, idx, idy, z = synth.splat_inputs(H, W, seed, "dense")
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.inp = (t(src), t(idx), t(idy), t(z))
        self.host = (src, idx, idy, z)
        self.warped = torch.zeros((H, W, 5), dtype=torch.uint8, device=device)
        n = self.L.mpiflow_forward_warp_workspace_bytes(H, W)
        self.ws = torch.empty(n + 256, dtype=torch.uint8, device=device)
        off = (-self.ws.data_ptr()) % 256
        self.args = [nat.ptr(x) for x in self.inp] + [nat.ptr(self.warped), H, W, ctypes.c_void_p(self.ws.data_ptr() + off), n]

    def __call__(self):
        self.warped.zero_()                                  # pre-zeroed by the caller (moving_obj.py:123)
        rc = self.L.mpiflow_forward_warp_f32(*self.args, self.nat.stream(self.warped))
        if rc:
            self.nat.check(rc)


def sample_assembly_leg(d, steps=24):
    """SURVEY 8(f) row 1 next to the headline: one dataset sample of the generator (utils/utils.py:159-288) = source
    blend + both poses of the sample in one launch on a shared MPI + merge, device-resident, rotated over the MPIs of
    the batch (503 MB each > L2).  Reported beside the metric, never part of it."""
    try:
        from utils import utils as U
        from utils.mpi import mpi_rendering as R
        check, R.CHECK_SINGULAR = R.CHECK_SINGULAR, False      # no host sync inside the timed loop
        n = d["mpi"].shape[0]
        src = d["mpi"][:, 0, 0:3].contiguous()                   # any Bx3xHxW image

        def one(i):
            k = i % n
            j = (k + 1) % n
            return U.render_sample_pair(src[k:k + 1], d["obj_mask"][k:k + 1], d["mpi"][k:k + 1], d["disparity"][k:k + 1],
                                        d["K"][k:k + 1], d["K"][k:k + 1], d["G"][k], d["G"][j], k_src_inv=d["Kinv"][k:k + 1])
        for i in range(3):
            one(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            one(i)
        e1.record()
        torch.cuda.synchronize()
        R.CHECK_SINGULAR = check
        us = e0.elapsed_time(e1) / steps * 1e3
        return {"value": round(1e6 / us, 1), "unit": "samples/s", "us_per_sample": round(us, 1), "steps": steps,
                "api": "utils.utils.render_sample_pair (blend + 2 poses in one launch on a shared MPI + merge), "
                       "device-resident, CUDA events"}
    except Exception as e:  # an auxiliary leg must never take the headline down
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def hole_filling_leg(device, batch=148, unique=8, steps=3):
    """SURVEY 8(f) row 4 next to the headline: cv2.inpaint(INPAINT_NS) of the generator (utils/utils.py:284-286) on the
    device, a batch of generator-like masks (the disocclusion ring of one moved object + the strips a camera move
    uncovers at the border, 1.7 % holes; `unique` different images repeated to one image per SM).  One image is checked against the C restatement of OpenCV's algorithm
    (oracle/inpaint_oracle.c, pinned against cv2), which is also timed on one host core.  Reported beside the metric."""
    try:
        from utils import utils as U
        g = torch.Generator().manual_seed(synth.SEED)
        img = torch.nn.functional.avg_pool2d(torch.rand((unique, 3, H + 8, W + 8), generator=g), 9, 1)     # smooth, HxW
        img = (img * 255).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
        yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        masks = []
        for n in range(unique):
            dx, dy = 3 + n % 4, (n % 7) - 3
            a = ((xx - 0.4 * W) / (0.18 * W)) ** 2 + ((yy - 0.55 * H) / (0.3 * H)) ** 2 < 1
            b = ((xx - 0.4 * W - dx) / (0.18 * W)) ** 2 + ((yy - 0.55 * H - dy) / (0.3 * H)) ** 2 < 1
            m = a ^ b
            m[:, :4] = True
            m[:3, :] = True
            masks.append(m)
        mask = torch.stack(masks).to(torch.uint8)
        rep = [n % unique for n in range(batch)]
        d_img, d_mask = img.to(device)[rep].contiguous(), mask.to(device)[rep].contiguous()
        out = U.inpaint_ns(d_img, d_mask, 3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            U.inpaint_ns(d_img, d_mask, 3)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        from oracle import inpaint_oracle as io          # CPU leg: the checker, timed as the host baseline
        t0 = time.perf_counter()
        want = io.inpaint_ns(img[0].numpy(), mask[0].numpy(), 3)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        return {"value": round(batch / ms * 1e3, 1), "unit": "images/s", "batch": batch, "ms_per_launch": round(ms, 3),
                "hole_fraction": round(float(mask.float().mean()), 4), "bit_exact_vs_oracle": bool((out[0].cpu().numpy() == want).all()),
                "cpu_one_core_ms_per_image": round(cpu_ms, 2),
                "api": "utils.utils.inpaint_ns -> mpiflow_inpaint_ns_u8 (OpenCV's Navier-Stokes fast-marching inpainting, "
                       "byte-identical), device-resident, CUDA events"}
    except Exception as e:  # an auxiliary leg must never take the headline down
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def api_leg(d, steps=10):
    """The same batch through the reference-shaped Python API, in the reference's own call sequence
    (utils/utils.py:304-331: get_src_xyz_from_plane_disparity -> get_tgt_xyz_from_plane_disparity ->
    render_tgt_rgb_depth), device-resident, CUDA events.  Reported beside the metric."""
    try:
        from utils.mpi import mpi_rendering as R
        from utils.mpi.homography_sampler import HomographySample
        check, R.CHECK_SINGULAR = R.CHECK_SINGULAR, False
        mpi = d["mpi"]
        B = mpi.shape[0]
        sampler = HomographySample(H, W, mpi.device)
        obj = d["obj_mask"].unsqueeze(1).expand(B, S, 1, H, W)

        def one():
            xyz_src = R.get_src_xyz_from_plane_disparity(sampler.meshgrid, d["disparity"], d["Kinv"])
            xyz_tgt = R.get_tgt_xyz_from_plane_disparity(xyz_src, d["G"])
            return R.render_tgt_rgb_depth(sampler, mpi[:, :, 0:3], mpi[:, :, 3:], d["disparity"], xyz_tgt, xyz_src, d["G"],
                                          d["Kinv"], d["K"], None, use_alpha=False, is_bg_depth_inf=False, obj_mask=obj)
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        torch.cuda.synchronize()
        R.CHECK_SINGULAR = check
        ms = e0.elapsed_time(e1) / steps
        return {"value": round(B / ms * 1e3, 1), "unit": "frames/s", "ms_per_step": round(ms, 4), "steps": steps,
                "api": "mpi_rendering.get_src_xyz_from_plane_disparity + get_tgt_xyz_from_plane_disparity + "
                       "render_tgt_rgb_depth (reference call sequence), device-resident, CUDA events"}
    except Exception as e:  # an auxiliary leg must never take the headline down
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    import _native as nat
    import shard
    L = nat.lib()
    B = args.frames_per_step
    strong = args.config == "c4"
    if strong:
        # config 4: a fixed batch of 256 frames split over the ranks; a step renders this rank's whole shard in
        # launches of B frames, cycling over the frames that are resident (up to 64 per GPU, 20 GB)
        lo, hi = shard.shard_range(256, rank, world)
        shard_frames = hi - lo
        launches_per_step = -(-shard_frames // B)
        NBUF = max(2, min(launches_per_step, 64 // B))
        frames_per_step_all = 256
    else:
        launches_per_step, NBUF, frames_per_step_all = 1, 2, world * B
    calls = [FusedCall(make_device_batch(B, synth.SEED + 1000 * rank + i, device), device) for i in range(NBUF)]
    splat = SplatCall(device) if args.config == "c3" else None
    peak, peak_src = measured_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return x if world == 1 else shard.max_over_ranks(x, device)

    def step(i):
        for j in range(launches_per_step):
            calls[(i * launches_per_step + j) % NBUF]()
        if splat is not None:
            for _ in range(B):
                splat()

    # ---- device-resident throughput ----
    for i in range(args.warmup):
        step(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        ev0.record()
        for i in range(args.steps):
            step(i)
        ev1.record()
        barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    assert int(calls[0].status.item()) == 0
    ms_per_step = ms / args.steps
    value = frames_per_step_all * args.steps / (ms * 1e-3)
    frames_per_launch_all = frames_per_step_all / world / launches_per_step     # per GPU and launch
    bytes_per_launch = frames_per_launch_all * BYTES_PER_FRAME
    achieved = bytes_per_launch * launches_per_step / (ms_per_step * 1e-3) / 1e9

    # ---- same, with the 8 output channels of every rank gathered to every rank inside the timed region (N > 1) ----
    # One all_gather_into_tensor per step into preallocated buffers, issued on a side stream while the next step renders
    # into the other slot; the persistent kernel leaves a few SMs to the collective (MPIFLOW_RESERVE_SMS).
    gather = None
    if world > 1 and not strong:
        gsteps = min(args.steps, 50)
        reserve = args.reserve_sms
        g = shard.FrameGatherer(world * B, H, W, device, mode=args.gather)
        gcalls = [FusedCall(c.d, device, out=g.outputs(k % 2), flags=reserve << 16) for k, c in enumerate(calls[:2])]

        def gstep(i):
            slot = i % 2
            g.wait(slot)            # the gather that last read this slot's send buffer is done
            gcalls[slot]()
            g.start(slot)
        for i in range(4):
            gstep(i)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(gsteps):
            gstep(i)
        g.wait(0); g.wait(1)
        g1.record()
        barrier()
        gms = max_over_ranks(g0.elapsed_time(g1))
        gather = {"value": round(world * B * gsteps / (gms * 1e-3), 2), "unit": "frames/s", "steps": gsteps,
                  "bytes_per_step_per_rank": int(g.block * 4), "reserved_sms": reserve, "transport": g.mode,
                  "why_not_p2p": g.why_not_p2p,
                  "collective": ("copy-engine pulls of the packed 8 output channels out of the peers' symmetric memory "
                                 "(one stream per peer, device-side barriers), under the next render" if g.mode == "p2p" else
                                 "copy-engine pushes of the packed 8 output channels into the peers' symmetric receive buffers "
                                 "(one stream per peer, device-side barriers), under the next render" if g.mode == "push" else
                                 "one NCCL all_gather_into_tensor of the packed 8 output channels per step, preallocated, "
                                 "on a side stream overlapping the next render")}
        # the gathered frames are the rendered frames (slot 0 of the last even step, every rank's block)
        torch.cuda.synchronize()
        mine = g.views(0)[0][rank]
        assert torch.equal(mine, gcalls[0].out[0]), "gathered block differs from the rendered one"
        # every rank holds every other rank's frames: compare checksums of what arrived with what each rank rendered
        sums = torch.stack([v.double().sum() for v in g.views(0)[0]])
        own = torch.zeros(world, dtype=torch.float64, device=device)
        own[rank] = gcalls[0].out[0].double().sum()
        dist.all_reduce(own)
        assert torch.equal(sums, own), f"rank {rank}: gathered frames differ from the rendered ones"
        gather["verified"] = "checksums of all ranks' blocks"

    # ---- end to end through the host-buffer C-ABI entry point ----
    Be = args.e2e_frames_per_step
    src = calls[0].d
    pin = lambda t: t.cpu().pin_memory()
    h_in = {k: pin(src[k][:Be]) for k in ("mpi", "obj_mask", "disparity", "G", "Kinv", "K")}
    h_out = [torch.empty((Be, c, H, W), dtype=torch.float32).pin_memory() for c in (3, 1, 1, 2, 1)]
    hp = lambda t: ctypes.c_void_p(t.data_ptr())
    host_args = [hp(h_in["mpi"]), hp(h_in["obj_mask"]), hp(h_in["disparity"]), hp(h_in["G"]), hp(h_in["Kinv"]),
                 hp(h_in["K"]), Be, S, H, W, 0, *[hp(o) for o in h_out]]
    e2e_steps = max(1, min(args.steps, args.e2e_max_steps))
    for _ in range(2):
        nat.check(L.mpiflow_render_fused_host(*host_args))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        nat.check(L.mpiflow_render_fused_host(*host_args))  # synchronous: results are in h_out on return
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * Be * e2e_steps / e2e_s
    # the host path renders frame by frame (smaller tiles per launch), the device path a whole batch: same kernel,
    # but which warps take the lean / border loop differs with the tile shape -> agreement to rounding, not bit for bit
    calls[0]()
    torch.cuda.synchronize()
    dev_rgb = calls[0].out[0][:Be].cpu()
    assert (h_out[0] - dev_rgb).abs().max().item() <= 2e-6, "host-buffer path and device path disagree"
    h2d = sum(h_in[k].numel() * 4 for k in h_in)
    d2h = sum(o.numel() * 4 for o in h_out)

    # ---- the same with the MPI and the mask stored as fp16 on the host (mpiflow_render_fused_host_f16): half the PCIe bytes
    h16 = {k: src[k][:Be].half().cpu().pin_memory() for k in ("mpi", "obj_mask")}
    host16_args = [hp(h16["mpi"]), hp(h16["obj_mask"])] + host_args[2:]
    for _ in range(2):
        nat.check(L.mpiflow_render_fused_host_f16(*host16_args))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        nat.check(L.mpiflow_render_fused_host_f16(*host16_args))
    torch.cuda.synchronize()
    e2e16_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e16 = {"value": round(world * Be * e2e_steps / e2e16_s, 2), "unit": "frames/s",
             "h2d_bytes_per_step": h2d - sum(h_in[k].numel() * 2 for k in ("mpi", "obj_mask")), "d2h_bytes_per_step": d2h,
             "frames_per_step_per_gpu": Be, "steps": e2e_steps,
             "api": "mpiflow_render_fused_host_f16 (fp16 storage of the MPI, fp32 arithmetic and outputs; pinned host buffers)"}

    solo = rank == 0 and world == 1
    sample_leg = sample_assembly_leg(calls[0].d) if (solo and args.config == "c2") else None
    api = api_leg(calls[0].d) if solo else None
    fill_leg = hole_filling_leg(device) if (solo and args.config == "c2" and not args.no_cpu_baseline) else None

    # ---- CPU baseline on a bounded sample (rank 0, single-GPU run only) ----
    cpu = None
    if solo and not args.no_cpu_baseline:
        cpu = cpu_baseline_c_port()

    if rank == 0:
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes/launch of the fused kernel from ncu
        if os.path.exists(tp) and args.config in ("c2", "c3"):
            t = json.load(open(tp))
            traffic = t.get("dram_bytes_per_frame", None)
            traffic = None if traffic is None else traffic * frames_per_launch_all
        launches = (4 if B >= 4 else 3) * launches_per_step + (3 * B if splat is not None else 0)
        line = {
            "metric": METRIC if args.config == "c2" else f"MPI frames/sec ({S} planes, {W}x{H})",
            "value": round(value, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "planes": S, "height": H, "width": W, "frames_per_launch_per_gpu": B,
                       "launches_per_step_per_gpu": launches_per_step,
                       "l2": f"inputs {B * S * 4 * H * W * 4 / 1e9:.2f} GB per launch (> 126 MB L2), rotated over {NBUF} buffers",
                       "sharding": "frames split across ranks, no data-path collective; NCCL only for barrier/max"},
            "clocks": clk.summary(),
            "e2e": {"value": round(e2e_value, 2), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "frames_per_step_per_gpu": Be, "steps": e2e_steps,
                    "api": "mpiflow_render_fused_host (pinned host buffers)"},
            "e2e_f16_storage": e2e16,
            # per launch: plane_consts + batch_flags (+ spacing_table when >= 4 frames share K^-1 and disparities) +
            # render_fused_pk; config 3 adds the 3 kernels of one splat per frame (chain build, resolve, the gated radix launch)
            "gpu_launches": launches * args.steps,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                         "frac_of_datasheet_8TBps": round(achieved / 8000.0, 4),
                         "kernel": "render_fused_pk_kernel (persistent, warp-specialised; step = prologue + this kernel)",
                         "bytes_per_launch": int(bytes_per_launch),
                         "bound_note": "algorithmic HBM bytes vs measured copy bandwidth (DESIGN.md 5.1)"},
            "cpu_baseline": cpu,
        }
        if gather:
            line["with_output_gather"] = gather
        if sample_leg:
            line["sample_assembly"] = sample_leg
        if api:
            line["reference_api"] = api
        if fill_leg:
            line["hole_filling"] = fill_leg
        OUT.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# CPU legs (the only places allowed to execute oracle/)
# ------------------------------------------------------------------------------------------------
def _strip_inputs(rows, seed=synth.SEED):
    """a horizontal strip of the frame: same width, planes, intrinsics and pose statistics"""
    inp = synth.frame_inputs(1, S, rows, W, "white", "random", seed)
    K = synth.intrinsics(H, W, 1)  # intrinsics of the full frame
    inp["K"], inp["Kinv"] = K, synth.intrinsics_inv(K)
    return inp


def cpu_baseline_c_port():
    from oracle import mpi_oracle as orc
    cores = os.cpu_count()
    small = _strip_inputs(16)
    t0 = time.perf_counter()
    orc.render_frame(small["mpi"], small["disparity"], small["G"], small["Kinv"], small["K"], small["obj_mask"], numerics=1)
    per_row = (time.perf_counter() - t0) / 16
    rows = int(min(H, max(16, (15.0 / per_row) // 8 * 8)))
    inp = _strip_inputs(rows)
    t0 = time.perf_counter()
    orc.render_frame(inp["mpi"], inp["disparity"], inp["G"], inp["Kinv"], inp["K"], inp["obj_mask"], numerics=1)
    dt = time.perf_counter() - t0
    return {"value": round((rows / H) / dt, 4), "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"oracle/mpi_oracle.c (C restatement of the path, OpenMP over {cores} threads), one {W}x{rows}x{S} strip "
                      f"({rows}/{H} of a frame) in {dt:.2f} s"}


def run_reference(args, rank):
    """The reference's own CPU path: its unmodified utils/mpi/*.py (oracle/_ref/refsrc, byte copies made by
    oracle/Makefile) on torch-CPU with every host thread, one WHOLE frame of the configuration per step."""
    if rank != 0:
        return
    from oracle import ref_modules, torch_ref
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    T = torch.from_numpy
    real = ref_modules.available()
    render = ref_modules.render_frame if real else torch_ref.render_frame
    inp = synth.frame_inputs(1, S, H, W, "white", "random", synth.SEED)
    t = {k: T(v) for k, v in inp.items()}
    splat_in = synth.splat_inputs(H, W, 7, "dense") if args.config == "c3" else None

    def step():
        t0 = time.perf_counter()
        with torch.no_grad():
            render(t["mpi"], t["disparity"], t["G"], t["Kinv"], t["K"], t["obj_mask"])
        if splat_in is not None:
            from oracle import mpi_oracle as orc
            orc.forward_warping(*splat_in, use_ref=orc.ref_splat_lib() is not None)   # warping.c, serial by construction
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    total = sum(times)
    value = args.steps / total
    kind = "reference" if real else "port"
    what = ("the unmodified utils/mpi/*.py of the reference (oracle/_ref/refsrc): get_src_xyz_from_plane_disparity, "
            "get_tgt_xyz_from_plane_disparity, render_tgt_rgb_depth" if real else
            "oracle/torch_ref.py: the reference's operator sequence restated (oracle/_ref/refsrc absent)")
    sample = f"{what}; torch-CPU, {cores} threads; each step one whole {W}x{H}x{S} frame"
    line = {"impl": "reference", "metric": METRIC if args.config == "c2" else f"MPI frames/sec ({S} planes, {W}x{H})",
            "value": round(value, 4), "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total / args.steps, 2),
            "higher_is_better": True, "scaling": "strong" if args.config == "c4" else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "planes": S, "height": H, "width": W, "frames_per_step": 1},
            "cpu_baseline": {"value": round(value, 4), "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": round(value, 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    OUT.emit(json.dumps(line))


class QuietStdout:
    """The contract is ONE JSON line on stdout.  Libraries write there as well (NCCL prints its version banner at
    communicator creation): while the benchmark runs, file descriptor 1 points at stderr; emit() restores it for
    the single line."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)


OUT = None


def main():
    global OUT
    OUT = QuietStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--frames-per-step", type=int, default=0, help="frames per launch per GPU (0 = the configuration's own)")
    ap.add_argument("--gather", default="auto", choices=["auto", "p2p", "push", "nccl"], help="transport of the output-gather leg")
    ap.add_argument("--reserve-sms", type=int, default=0, help="SMs left to the output gather (with_output_gather leg)")
    ap.add_argument("--e2e-frames-per-step", type=int, default=2)
    ap.add_argument("--e2e-max-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    frames = set_workload(args.config)
    args.frames_per_step = args.frames_per_step or frames
    if args.impl == "reference":
        if args.steps == 200:  # defaults for the slow CPU arm
            args.steps, args.warmup = 3, 1
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    if world != args.gpus and rank == 0:
        print(f"note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
