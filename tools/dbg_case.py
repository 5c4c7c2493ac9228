"""Debug aid: one synthetic case through the fused renderer, compared with the oracle (CUDA numerics);
prints the error per output and where the worst rgb / depth pixels are.  usage: dbg_case.py B S H W kind pose seed"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mpi-flow_b200")]
import synth
from oracle import mpi_oracle as orc
from utils.mpi import mpi_rendering as R
B, S, H, W = (int(v) for v in sys.argv[1:5]); kind, pose, seed = sys.argv[5], sys.argv[6], int(sys.argv[7])
inp = synth.frame_inputs(B, S, H, W, kind, pose, seed)
d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
out = R.render_fused(d["mpi"][:, :, 0:3], d["mpi"][:, :, 3:4], d["disparity"], d["G"], d["Kinv"], d["K"], d["obj_mask"])
torch.cuda.synchronize()
got = [o.cpu().numpy() for o in out]
want = orc.render_frame(inp["mpi"], inp["disparity"], inp["G"], inp["Kinv"], inp["K"], inp["obj_mask"], numerics=orc.NUM_CUDA)
for n, g, w in zip(["rgb", "depth", "mask", "flow", "obj"], got, want):
    e = np.abs(g - w)
    idx = np.unravel_index(np.argmax(e), e.shape)
    print(f"{n:6s} max err {e.max():.3e} at {idx}  got {g[idx]:.6f} want {w[idx]:.6f}   #>1e-4: {(e > 1e-4).sum()}")
e = np.abs(got[0] - want[0]).max(axis=(0, 1))
ys, xs = np.nonzero(e > 1e-4)
if len(ys):
    print("rgb >1e-4 rows", sorted(set(ys.tolist()))[:40], "cols", sorted(set(xs.tolist()))[:60])
np.savez(os.environ.get("DBG_OUT", "/tmp/dbg_out.npz"), *got)
