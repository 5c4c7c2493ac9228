"""Debug aid: full-size (config 2) frame against oracle/torch_ref.py run on the same GPU (the reference's operator
sequence on cuBLAS / ATen).  usage: dbg_fullsize.py kind pose [S H W]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mpi-flow_b200")]
import synth
from oracle import torch_ref
from utils.mpi import mpi_rendering as R
kind, pose = sys.argv[1], sys.argv[2]
S, H, W = (int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (64, 384, 1280)
inp = synth.frame_inputs(1, S, H, W, kind, pose, 77)
d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
got = R.render_fused(d["mpi"][:, :, 0:3], d["mpi"][:, :, 3:4], d["disparity"], d["G"], d["Kinv"], d["K"], d["obj_mask"])
want = torch_ref.render_frame(d["mpi"], d["disparity"], d["G"], d["Kinv"], d["K"], d["obj_mask"])
torch.cuda.synchronize()
for n, g, w in zip(["rgb", "depth", "mask", "flow", "obj"], got, want):
    e = (g - w).abs()
    rel = e / (1 + w.abs()) if n == "depth" else e
    print(f"{n:6s} max err {rel.max().item():.3e}   #>1e-4: {(rel > 1e-4).sum().item()} of {e.numel()}")
e = (got[3] - want[3]).abs()[0]
bad = (e > 1e-4).any(dim=0)
ys, xs = torch.nonzero(bad, as_tuple=True)
if len(xs):
    print("flow >1e-4: x range", xs.min().item(), xs.max().item(), "y range", ys.min().item(), ys.max().item(),
          " |flow| there", got[3][0][:, bad].abs().max().item(), " max |flow| overall", got[3].abs().max().item())
    k = torch.argmax(e.max(dim=0).values.flatten()).item(); yy, xx = divmod(k, e.shape[2])
    print("worst pixel", yy, xx, "got", got[3][0, :, yy, xx].tolist(), "want", want[3][0, :, yy, xx].tolist())
