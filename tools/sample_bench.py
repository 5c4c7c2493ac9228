"""Timing of one dataset sample on the device: the source blend, the two-pose render with a
shared MPI against the same two frames rendered from two distinct MPIs, and the merge; 1280x384x64 by default.

    python tools/sample_bench.py [H W S [steps]]       -> one JSON line (CUDA events, 5 warm-up rounds)

Two sample buffers are rotated so that nothing is served from L2 across iterations.
"""
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mpi-flow_b200")]
import synth                                     # noqa: E402
from utils import utils as U                     # noqa: E402
from utils.mpi import mpi_rendering as R         # noqa: E402

R.CHECK_SINGULAR = False                         # no host sync inside the timed loops


def timed(fn, steps):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3     # us


def main():
    H, W, S = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (384, 1280, 64)))
    steps = int(sys.argv[4]) if len(sys.argv) >= 5 else 30
    T = lambda a: torch.from_numpy(a).cuda()
    samples = []
    for k in range(2):
        inp = synth.frame_inputs(2, S, H, W, "white", "random", 114514 + k)
        g = torch.Generator().manual_seed(k)
        samples.append(types.SimpleNamespace(
            mpi2=T(inp["mpi"]), src=torch.rand((1, 3, H, W), generator=g).cuda(), obj=T(inp["obj_mask"][:1]),
            disp=T(inp["disparity"][:1]), K=T(inp["K"][:1]), Kinv=T(inp["Kinv"][:1]), G=T(inp["G"])))
    out = {"config": {"H": H, "W": W, "S": S, "steps": steps}}

    def blend(i):
        s = samples[i % 2]
        return U.blend_src_into_mpi(s.src, s.mpi2[:1, :, :3], s.mpi2[:1, :, 3:], s.disp, s.Kinv)
    us = timed(blend, steps)
    out["blend_us"] = us
    out["blend_GBps"] = 4 * H * W * (7 * S + 3) / us / 1e3
    if os.environ.get("SAMPLE_BENCH_ONLY") == "blend":
        print(json.dumps(out))
        return

    def masks(s):
        return torch.cat((s.obj, 1 - s.obj), 0).reshape(2, 1, 1, H, W)

    def pair_shared(i):
        s = samples[i % 2]
        return R.render_fused(s.mpi2[:1, :, :3].expand(2, S, 3, H, W), s.mpi2[:1, :, 3:].expand(2, S, 1, H, W),
                              s.disp.expand(2, S), s.G, s.Kinv.expand(2, 3, 3), s.K.expand(2, 3, 3), masks(s))

    def pair_distinct(i):
        s = samples[i % 2]
        return R.render_fused(s.mpi2[:, :, :3], s.mpi2[:, :, 3:], s.disp.expand(2, S), s.G, s.Kinv.expand(2, 3, 3),
                              s.K.expand(2, 3, 3), masks(s))
    out["pair_shared_mpi_us"] = timed(pair_shared, steps)
    out["pair_distinct_mpi_us"] = timed(pair_distinct, steps)

    def sample(i):
        s = samples[i % 2]
        return U.render_sample_pair(s.src, s.obj, s.mpi2[:1], s.disp, s.K, s.K, s.G[0], s.G[1], k_src_inv=s.Kinv)
    out["sample_us"] = timed(sample, steps)
    out["samples_per_s"] = 1e6 / out["sample_us"]

    # N samples per launch (MPIFLOW_POSES_PER_MPI(2)): 2N frames in one render launch
    N = int(os.environ.get("SAMPLES_PER_LAUNCH", 8))
    if N <= 0:
        print(json.dumps(out))
        return
    s = samples[0]
    batches = []
    for k in range(2):
        gen = torch.Generator(device="cuda").manual_seed(k)
        mpis = torch.rand((N, S, 4, H, W), device="cuda", generator=gen)
        mpis[:, :, 3].sub_(0.5).clamp_(min=0).add_(1e-4)
        pose_idx = torch.randint(0, 2, (N, 2), generator=torch.Generator().manual_seed(k))
        batches.append(types.SimpleNamespace(
            mpis=mpis, src=torch.rand((N, 3, H, W), device="cuda", generator=gen), obj=s.obj.expand(N, 1, H, W).contiguous(),
            disp=s.disp.expand(N, S).contiguous(), Kinv=s.Kinv.expand(N, 3, 3).contiguous(),
            K=s.K.expand(N, 3, 3).contiguous(), poses=s.G[pose_idx.reshape(-1)].reshape(N, 2, 4, 4).contiguous()))

    def batch(i):
        b = batches[i % 2]
        return U.render_sample_pairs(b.src, b.obj, b.mpis, b.disp, b.Kinv, b.K, b.poses)
    us = timed(batch, max(4, steps // 4))
    out["batched"] = {"samples_per_launch": N, "us_per_sample": us / N, "samples_per_s": N * 1e6 / us}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
