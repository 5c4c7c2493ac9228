"""One two-pose launch on a shared MPI and one on two distinct MPIs, for an ncu DRAM-traffic comparison:

    ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:render_fused_pk \
        --csv --log-file gpurun_out/pair_traffic.csv python tools/pair_traffic.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mpi-flow_b200")]
import synth                                     # noqa: E402
from utils.mpi import mpi_rendering as R         # noqa: E402

S, H, W = 64, 384, 1280
inp = synth.frame_inputs(2, S, H, W, "white", "random", 114514)
T = lambda a: torch.from_numpy(a).cuda()
mpi, G, Kinv, K, disp, obj = (T(inp[k]) for k in ("mpi", "G", "Kinv", "K", "disparity", "obj_mask"))
masks = torch.cat((obj[:1], 1 - obj[:1]), 0).reshape(2, 1, 1, H, W)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for shared in (True, False, True, False):
    flush.zero_()
    if shared:
        R.render_fused(mpi[:1, :, :3].expand(2, S, 3, H, W), mpi[:1, :, 3:].expand(2, S, 1, H, W), disp, G, Kinv, K, masks)
    else:
        R.render_fused(mpi[:, :, :3], mpi[:, :, 3:], disp, G, Kinv, K, masks)
torch.cuda.synchronize()
