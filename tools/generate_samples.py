"""Generator loop on synthetic MPIs: render_sample_pair -> SampleWriter (the shape of gen_3dphoto_dynamic_v2.py:98-122
without the AdaMPI network).  Prints one JSON line: samples/s of the device part alone and of the whole loop including
hole filling and file writing by the writer's worker threads.

    python tools/generate_samples.py [n_samples [workers [out_dir [device_inpaint_batch]]]]

device_inpaint_batch > 0: the holes are filled on the GPU in batches of that many samples (SampleWriter(device_inpaint=True)).
"""
import json
import os
import random
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mpi-flow_b200")]
import synth                                     # noqa: E402
from sample_writer import SampleWriter           # noqa: E402
from utils import utils as U                     # noqa: E402
from utils.mpi import mpi_rendering as R         # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    out_dir = sys.argv[3] if len(sys.argv) > 3 else tempfile.mkdtemp(prefix="mpiflow_samples_")
    dev_batch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    S, H, W = 64, 384, 1280
    R.CHECK_SINGULAR = False
    T = lambda a: torch.from_numpy(a).cuda()
    inp = synth.frame_inputs(1, S, H, W, "smooth", "random", 114514)
    mpi = T(inp["mpi"])
    # a surface-like scene, as a predicted MPI is: density concentrated around one plane per pixel (a smooth depth
    # field in the far two thirds of the volume), an opaque last plane, one hard instance mask (gen...v2.py:101);
    # holes are then the disocclusions around the moved object and at the image border (about 5 % of the pixels)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32, device="cuda"),
                            torch.arange(W, dtype=torch.float32, device="cuda"), indexing="ij")
    surface = S * (0.8 + 0.12 * torch.sin(xx / W * 6.0) * torch.cos(yy / H * 4.0))
    planes = torch.arange(S, dtype=torch.float32, device="cuda").view(1, S, 1, 1)
    mpi[:, :, 3] = 50 * torch.exp(-0.5 * (planes - surface) ** 2) + 1e-4
    mpi[:, -1, 3] += 10.0
    disp, K, Kinv = T(inp["disparity"]), T(inp["K"]), T(inp["Kinv"])
    obj = ((((xx - 0.4 * W) / (0.18 * W)) ** 2 + ((yy - 0.55 * H) / (0.3 * H)) ** 2) < 1).float()[None, None]
    src = mpi[:, 0, 0:3].contiguous()
    src_np = (src[0].permute(1, 2, 0) * 255).round().clamp(0, 255).to(torch.uint8).flip(-1).contiguous().cpu().numpy()
    random.seed(114514)
    poses = [(U.generate_random_pose(0.15, base_motions=[0, 0, 0]), U.generate_random_pose(0.15)) for _ in range(n)]

    def device_part(k):
        return U.render_sample_pair(src, obj, mpi, disp, K, K, poses[k][0], poses[k][1], k_src_inv=Kinv)[:3]

    for k in range(3):
        device_part(k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n):
        device_part(k)
    torch.cuda.synchronize()
    t_dev = time.perf_counter() - t0

    t0 = time.perf_counter()
    with SampleWriter(out_dir, workers=workers, depth=2 * workers, device_inpaint=dev_batch > 0, inpaint_batch=max(dev_batch, 1)) as w:
        for k in range(n):
            flow_mix, frame_mix, fill = device_part(k)
            w.submit(f"demo_{k}", flow_mix, frame_mix, fill, src_np)
        torch.cuda.synchronize()
        t_submit = time.perf_counter() - t0
    t_all = time.perf_counter() - t0
    holes = float(fill.float().mean().item())
    size = sum(os.path.getsize(os.path.join(out_dir, d, f)) for d in ("flows", "dst_images", "src_images")
               for f in os.listdir(os.path.join(out_dir, d)))
    print(json.dumps({"samples": n, "workers": workers, "device_only_samples_per_s": n / t_dev,
                      "submit_loop_samples_per_s": n / t_submit, "with_fill_and_files_samples_per_s": n / t_all,
                      "hole_fraction_last_sample": holes, "device_inpaint_batch": dev_batch, "bytes_written": size, "out_dir": out_dir}))


if __name__ == "__main__":
    main()
