"""One launch of the source-blend kernel and one of the merge kernel at 1280x384x64 (ncu target):
    ncu --set full --clock-control none --import-source on -k regex:"blend_src|assemble_sample" -o gpurun_out/blend python tools/blend_once.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mpi-flow_b200")]
import synth                                     # noqa: E402
from utils import utils as U                     # noqa: E402

S, H, W = 64, 384, 1280
inp = synth.frame_inputs(2, S, H, W, "white", "random", 114514)
T = lambda a: torch.from_numpy(a).cuda()
mpi, G, Kinv, K, disp, obj = (T(inp[k]) for k in ("mpi", "G", "Kinv", "K", "disparity", "obj_mask"))
src = torch.rand(1, 3, H, W, device="cuda")
for _ in range(2):
    U.render_sample_pair(src, obj[:1], mpi[:1], disp[:1], K[:1], K[:1], G[0], G[1], k_src_inv=Kinv[:1])
torch.cuda.synchronize()
