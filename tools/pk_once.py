"""A few launches of the fused renderer on config-2 frames, for profilers (ncu wraps this process).

    python tools/pk_once.py [frames_per_launch=16] [launches=4]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mpi-flow_b200")]
import torch  # noqa: E402

import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
calls = [bench.FusedCall(bench.make_device_batch(B, 114514 + i, dev), dev) for i in range(2)]
for i in range(n):
    calls[i % 2]()
torch.cuda.synchronize()
assert int(calls[0].status.item()) == 0
print("ok")
