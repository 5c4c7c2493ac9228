"""Throughput of the fused renderer over the other BASELINE.json configurations (device-resident inputs):
c1 256x256x32, c2 1280x384x64, c4 640x480x64 (32 frames = one GPU's share of batch 256), c5 1920x1080 x {16,32,64,128}.
One JSON line per configuration: frames/s, algorithmic GB/s, fraction of the measured HBM peak."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mpi-flow_b200")]
import bench  # noqa: E402
import synth  # noqa: E402

CONFIGS = [("c1", 16, 32, 256, 256), ("c2", 16, 64, 384, 1280), ("c4", 32, 64, 480, 640),
           ("c5-16", 4, 16, 1080, 1920), ("c5-32", 4, 32, 1080, 1920), ("c5-64", 4, 64, 1080, 1920), ("c5-128", 2, 128, 1080, 1920)]


def main():
    dev = torch.device("cuda", 0)
    peak, _ = bench.measured_peak()
    for name, B, S, H, W in CONFIGS:
        bench.S, bench.H, bench.W = S, H, W
        calls = [bench.FusedCall(bench.make_device_batch(B, synth.SEED + i, dev), dev) for i in range(2)]
        for i in range(3):
            calls[i % 2]()
        torch.cuda.synchronize()
        steps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            calls[i % 2]()
        e1.record()
        torch.cuda.synchronize()
        assert int(calls[0].status.item()) == 0
        ms = e0.elapsed_time(e1) / steps
        byts = 4 * H * W * (4 * S + 1 + 8) * B
        print(json.dumps({"config": name, "frames_per_launch": B, "planes": S, "height": H, "width": W,
                          "ms_per_launch": round(ms, 4), "frames_per_s": round(B / ms * 1e3, 1),
                          "algorithmic_GBps": round(byts / ms / 1e6, 1), "frac_hbm_peak": round(byts / ms / 1e6 / peak, 4)}), flush=True)
        del calls
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
