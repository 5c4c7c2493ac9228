"""Timing of the device hole filling (csrc/inpaint.cu) against cv2.inpaint on the host, generator frame size.

    python tools/inpaint_bench.py [batch ...]      -> JSON lines
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mpi-flow_b200")]
from utils import utils as U  # noqa: E402


def ring_case(rng, H, W):
    """what the generator's samples look like: the disocclusion ring around one moved object plus the strips a camera
    move uncovers at the image border -- ONE connected front, nothing to cut"""
    import cv2
    img = cv2.GaussianBlur((rng.random((H, W, 3)) * 255).astype(np.uint8), (0, 0), 2.0)
    a, b = np.zeros((H, W), np.uint8), np.zeros((H, W), np.uint8)
    cv2.ellipse(a, (int(0.4 * W), int(0.55 * H)), (int(0.18 * W), int(0.3 * H)), 0, 0, 360, 1, -1)
    cv2.ellipse(b, (int(0.4 * W) + int(rng.integers(3, 7)), int(0.55 * H) + int(rng.integers(-3, 4))),
                (int(0.18 * W), int(0.3 * H)), 0, 0, 360, 1, -1)
    mask = (a ^ b)
    mask[:, :4] = 1
    mask[:3, :] = 1
    return img, mask


def case(rng, H, W, n_lines):
    import cv2
    img = cv2.GaussianBlur((rng.random((H, W, 3)) * 255).astype(np.uint8), (0, 0), 2.0)
    mask = np.zeros((H, W), np.uint8)
    for _ in range(n_lines):
        x, y = int(rng.integers(0, W)), int(rng.integers(0, H))
        cv2.line(mask, (x, y), (int(np.clip(x + rng.integers(-40, 40), 0, W - 1)),
                                int(np.clip(y + rng.integers(-120, 120), 0, H - 1))), 1, int(rng.integers(1, 5)))
    mask[:, :2] = 1
    return img, mask


if __name__ == "__main__":
    import cv2
    H, W = 384, 1280
    rng = np.random.default_rng(1)
    batches = [int(a) for a in sys.argv[1:]] or [1, 8, 32, 148, 592]
    nmax = max(batches)
    ring = os.environ.get("INPAINT_BENCH_CASE", "slivers") == "ring"
    cases = [ring_case(rng, H, W) if ring else case(rng, H, W, 30) for _ in range(min(nmax, 32))]
    imgs = np.stack([cases[i % len(cases)][0] for i in range(nmax)])
    masks = np.stack([cases[i % len(cases)][1] for i in range(nmax)])
    frac = float(masks.astype(bool).mean())
    t0 = time.perf_counter()
    for i in range(8):
        ref = cv2.inpaint(imgs[i % len(cases)], masks[i % len(cases)], 3, cv2.INPAINT_NS)
    cpu_ms = (time.perf_counter() - t0) / 8 * 1e3
    dimgs, dmasks = torch.from_numpy(imgs).cuda(), torch.from_numpy(masks).cuda()
    for n in batches:
        a, m = dimgs[:n].contiguous(), dmasks[:n].contiguous()
        out = U.inpaint_ns(a, m, 3)
        torch.cuda.synchronize()
        ok = bool(np.array_equal(out[0].cpu().numpy(), cv2.inpaint(imgs[0], masks[0], 3, cv2.INPAINT_NS)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            U.inpaint_ns(a, m, 3)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"case": "ring" if ring else "slivers", "batch": n, "hole_fraction": round(frac, 4), "ms_per_launch": round(ms, 3),
                          "images_per_s": round(n / ms * 1e3, 1), "cv2_ms_per_image_one_core": round(cpu_ms, 2),
                          "equals_cv2": ok}), flush=True)
