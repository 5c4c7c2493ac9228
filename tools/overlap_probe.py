"""Does the hole-filling launch run UNDER the persistent renderer?  (Its CTAs need 9.6 KB of shared memory and 32 threads;
the renderer leaves ~11 KB and 1400 threads per SM free.)  Times N render launches alone, a batch of inpaint launches
alone on a side stream, and both together.

    python tools/overlap_probe.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mpi-flow_b200"), os.path.join(ROOT, "tools")]
import bench  # noqa: E402
import inpaint_bench as ib  # noqa: E402
from utils import utils as U  # noqa: E402

dev = torch.device("cuda", 0)
calls = [bench.FusedCall(bench.make_device_batch(16, 114514 + i, dev), dev) for i in range(2)]
rng = np.random.default_rng(1)
cases = [ib.ring_case(rng, 384, 1280) for _ in range(8)]
imgs = torch.from_numpy(np.stack([cases[i % 8][0] for i in range(148)])).cuda()
masks = torch.from_numpy(np.stack([cases[i % 8][1] for i in range(148)])).cuda()
side = torch.cuda.Stream()


def renders(n):
    for i in range(n):
        calls[i % 2]()


def fills(n):
    with torch.cuda.stream(side):
        for _ in range(n):
            U.inpaint_ns(imgs, masks, 3)


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


renders(4); fills(1); torch.cuda.synchronize()
with torch.cuda.stream(side):
    U.inpaint_ns(imgs, masks, 3)          # the side stream's own workspace
torch.cuda.synchronize()
a = timed(lambda: renders(60))
b = timed(lambda: fills(4))
c = timed(lambda: (fills(4), renders(60)))
print(json.dumps({"renders_60x16_frames_ms": round(a, 1), "fills_4x148_images_ms": round(b, 1), "both_ms": round(c, 1),
                  "sum_ms": round(a + b, 1), "overlap": round((a + b - c) / min(a, b), 2)}))
