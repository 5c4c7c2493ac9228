"""Forward splat (BASELINE config 3 adds it to the render): device entry point, legacy host-pointer symbol and the
reference's serial C loop (oracle/_ref, or the C restatement) on the same index maps.  One JSON line per size."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mpi-flow_b200")]
import _native as nat  # noqa: E402
import synth  # noqa: E402
from oracle import mpi_oracle as orc  # noqa: E402


def main():
    L = nat.lib()
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    for h, w in ((384, 1280), (480, 640), (1080, 1920)):
        src, idx, idy, z = synth.splat_inputs(h, w, 7, "dense")
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        d_src, d_idx, d_idy, d_z = t(src), t(idx), t(idy), t(z)
        warped = torch.zeros((h, w, 5), dtype=torch.uint8, device="cuda")
        n = L.mpiflow_forward_warp_workspace_bytes(h, w)
        ws = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
        off = (-ws.data_ptr()) % 256
        args = [ctypes.c_void_p(x.data_ptr()) for x in (d_src, d_idx, d_idy, d_z, warped)] + \
               [h, w, ctypes.c_void_p(ws.data_ptr() + off), n, nat.stream(warped)]
        for _ in range(3):
            warped.zero_()
            nat.check(L.mpiflow_forward_warp_f32(*args))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps):
            warped.zero_()
            nat.check(L.mpiflow_forward_warp_f32(*args))
        e1.record()
        torch.cuda.synchronize()
        dev_us = e0.elapsed_time(e1) / reps * 1e3
        got = warped.cpu().numpy()
        host = np.zeros((h, w, 5), np.uint8)
        L.forward_warping(P(src), P(idx), P(idy), P(z), P(host), h, w)
        t0 = time.perf_counter()
        for _ in range(10):
            host[:] = 0
            L.forward_warping(P(src), P(idx), P(idy), P(z), P(host), h, w)
        host_us = (time.perf_counter() - t0) / 10 * 1e6
        use_ref = orc.ref_splat_lib() is not None
        want = orc.forward_warping(src, idx, idy, z, use_ref=use_ref)
        t0 = time.perf_counter()
        for _ in range(10):
            orc.forward_warping(src, idx, idy, z, use_ref=use_ref)
        cpu_us = (time.perf_counter() - t0) / 10 * 1e6
        assert np.array_equal(got, want) and np.array_equal(host, want)
        print(json.dumps({"h": h, "w": w, "device_us": round(dev_us, 1), "io_bytes": 28 * h * w,
                          "device_GBps": round(28 * h * w / dev_us / 1e3, 1), "legacy_host_symbol_us": round(host_us, 1),
                          "cpu_serial_loop_us": round(cpu_us, 1), "cpu_kind": "reference warping.c" if use_ref else "C restatement",
                          "bit_exact": True}), flush=True)


if __name__ == "__main__":
    main()
