"""A/B timing of the fused renderer under different MPIFLOW_* environment settings (development aid).

    python tools/ab.py "MPIFLOW_PK_CFG=0" "MPIFLOW_PK_CFG=1" "MPIFLOW_RENDER_PATH=tma" ...

Each setting runs in its own process (the library reads the variables once): config-2 frames, 4 per launch,
two rotated buffers, CUDA-event time per launch and the implied fraction of the measured HBM peak.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, json, torch
sys.path[:0] = [%r, %r]
import bench
dev = torch.device('cuda', 0)
B = int(sys.argv[1]); steps = int(sys.argv[2])
import os
flags = int(os.environ.get('AB_FLAGS', '0'), 0)   # e.g. 4 = MPIFLOW_KEEP_TABLE
calls = [bench.FusedCall(bench.make_device_batch(B, 114514 + i, dev), dev, flags=flags) for i in range(2)]
for i in range(5): calls[i %% 2]()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps): calls[i %% 2]()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
assert int(calls[0].status.item()) == 0
print(json.dumps({'ms_per_launch': ms, 'us_per_frame': ms * 1e3 / B, 'frames_per_s': B / ms * 1e3,
                  'frac_hbm': B * bench.BYTES_PER_FRAME / (ms * 1e-3) / 1e9 / bench.measured_peak()[0]}))
""" % (ROOT, os.path.join(ROOT, "mpi-flow_b200"))

if __name__ == "__main__":
    B = int(os.environ.get("AB_FRAMES", 4))
    steps = int(os.environ.get("AB_STEPS", 40))
    for setting in sys.argv[1:] or [""]:
        env = dict(os.environ)
        for kv in setting.split():
            k, v = kv.split("=", 1)
            env[k] = v
        r = subprocess.run([sys.executable, "-c", CHILD, str(B), str(steps)], env=env, capture_output=True, text=True)
        out = r.stdout.strip().splitlines()
        print(f"{setting or '(default)':40s} {out[-1] if out else 'FAILED: ' + r.stderr[-400:]}", flush=True)
