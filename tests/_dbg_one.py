import sys, numpy as np, torch
sys.path[:0]=['/root/repo','/root/repo/mpi-flow_b200']
import synth
from oracle import mpi_oracle as orc
from utils.mpi import mpi_rendering as R
for case in [(1, 8, 48, 64, "smooth", "regression", 1), (2, 6, 40, 56, "white", "random", 2), (1, 16, 96, 320, "smooth", "random", 4)]:
    B,S,H,W,kind,pose,seed=case
    inp = synth.frame_inputs(B,S,H,W,kind,pose,seed)
    d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    o = R.render_fused(d['mpi'][:, :, 0:3], d['mpi'][:, :, 3:4], d['disparity'], d['G'], d['Kinv'], d['K'], d['obj_mask'])
    torch.cuda.synchronize()
    want = orc.render_frame(inp["mpi"], inp["disparity"], inp["G"], inp["Kinv"], inp["K"], inp["obj_mask"], numerics=1)
    for n,a,b in zip(['rgb','depth','mask','flow','obj'],o,want):
        e=np.abs(a.cpu().numpy()-b); idx=np.unravel_index(e.argmax(),e.shape)
        print(case, n, float(e.max()), idx, float(a.cpu().numpy()[idx]), float(b[idx]))
