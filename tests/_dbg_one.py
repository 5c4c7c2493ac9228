import sys, numpy as np, torch
sys.path[:0]=['/root/repo','/root/repo/mpi-flow_b200']
import synth
from utils.mpi import mpi_rendering as R
inp = synth.frame_inputs(1, 8, 48, 64, 'smooth', 'regression', 1)
d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
o = R.render_fused(d['mpi'][:, :, 0:3], d['mpi'][:, :, 3:4], d['disparity'], d['G'], d['Kinv'], d['K'], d['obj_mask'])
torch.cuda.synchronize()
print('ok', [float(t.abs().max()) for t in o])
