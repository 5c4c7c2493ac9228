"""Asynchronous sample writer -- the step after the path (reference
gen_3dphoto_dynamic_v2.py:107-122 and utils/utils.py:284-286).

The reference finishes every sample synchronously: device->host copies, cv2.inpaint, writeFlow and two cv2.imwrite
calls sit between two renders, so the GPU idles while the disk and OpenCV work.  Here submit() only enqueues three
device->pinned-host copies on a side stream (flow_mix, frame_mix, fill_mask -- 12.3 + 1.5 + 0.5 MB at 1280x384) and
returns; worker threads wait for the copy event, fill the holes (cv2.inpaint, INPAINT_NS, radius 3, as :284-286) and
write  flows/<name>.flo, dst_images/<name>.png, src_images/<name>.png  exactly as the reference names them.
A bounded ring of pinned slots gives back-pressure instead of unbounded host memory.

device_inpaint=True moves the hole filling to the GPU as well (utils.utils.inpaint_ns: cv2.inpaint(INPAINT_NS) bit for
bit, csrc/inpaint.cu): submit() copies the sample into a device staging batch; every `inpaint_batch` samples one launch
on the side stream marches all their holes (it fits next to the persistent renderer on every SM, so the next renders
run under it), then the filled frames and the flows go to the pinned slots -- the fill masks never cross PCIe and the
workers only encode and write.
"""
import os
import queue
import threading

import numpy as np
import torch

from write_flow import writeFlow


class SampleWriter:
    def __init__(self, out_dir, workers=2, depth=4, device_inpaint=False, inpaint_batch=8):
        import cv2  # PNG encoding; hole filling too unless device_inpaint
        self._cv2 = cv2
        self._device_inpaint = bool(device_inpaint)
        self._batch = max(1, int(inpaint_batch))
        self._stages, self._cur = None, 0       # two device staging batches (device_inpaint)
        self.out_dir = out_dir
        for sub in ("flows", "dst_images", "src_images"):
            os.makedirs(os.path.join(out_dir, sub), exist_ok=True)
        self._depth = max(depth, 2 * self._batch) if self._device_inpaint else depth   # a whole batch takes slots at once
        self._free = queue.Queue()          # pinned slots, created on first use (shapes are not known before)
        self._slots_made = 0
        self._jobs = queue.Queue()
        self._errors = []
        self._copy_stream = None
        self._threads = [threading.Thread(target=self._work, daemon=True) for _ in range(workers)]
        for t in self._threads:
            t.start()

    # ------------------------------------------------------------------------------------------------
    def _slot(self, shapes):
        """a set of pinned host buffers matching `shapes`; blocks when all `depth` slots are in flight"""
        while True:
            if self._slots_made < self._depth and self._free.empty():
                self._slots_made += 1
                pin = torch.cuda.is_available()
                return [torch.empty(s, dtype=d, pin_memory=pin) for s, d in shapes]
            slot = self._free.get()
            if [(tuple(t.shape), t.dtype) for t in slot] == shapes:
                return slot
            self._slots_made -= 1           # shape changed: drop the old slot, make a new one

    def submit(self, name, flow_mix, frame_mix, fill_mask, src_np):
        """flow_mix HxWx2 f32, frame_mix HxWx3 u8 (BGR, holes 255), fill_mask HxW u8 (1 = hole): CUDA or CPU tensors /
        arrays, e.g. the outputs of utils.utils.render_sample_pair; src_np HxWx3 u8 BGR (host).  Returns at once."""
        if self._errors:
            raise self._errors[0]
        tensors = [torch.as_tensor(t) for t in (flow_mix, frame_mix, fill_mask)]
        if self._device_inpaint:
            if not all(t.is_cuda for t in tensors):
                raise RuntimeError("SampleWriter(device_inpaint=True) takes CUDA tensors (no CPU fallback of the device path)")
            return self._submit_staged(name, tensors, np.ascontiguousarray(src_np))
        shapes = [(tuple(t.shape), t.dtype) for t in tensors]
        slot = self._slot(shapes)
        event = None
        if any(t.is_cuda for t in tensors):
            dev = next(t.device for t in tensors if t.is_cuda)
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            self._copy_stream.wait_stream(torch.cuda.current_stream(dev))     # after the kernels that produced them
            with torch.cuda.stream(self._copy_stream):
                for h, t in zip(slot, tensors):
                    h.copy_(t, non_blocking=True)
                    if t.is_cuda:
                        t.record_stream(self._copy_stream)
                event = torch.cuda.Event()
                event.record(self._copy_stream)
        else:
            for h, t in zip(slot, tensors):
                h.copy_(t)
        self._jobs.put((name, slot, event, np.ascontiguousarray(src_np), False))

    # ---- device_inpaint: stage on the device, fill a batch of samples with one launch ----------------------------
    def _submit_staged(self, name, tensors, src_np):
        flow, frame, fill = tensors
        dev = frame.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        B = self._batch
        if self._stages is None or tuple(self._stages[0]["frame"].shape[1:]) != tuple(frame.shape):
            self.flush()
            self._stages = [{"flow": torch.empty((B,) + tuple(flow.shape), dtype=flow.dtype, device=dev),
                             "frame": torch.empty((B,) + tuple(frame.shape), dtype=torch.uint8, device=dev),
                             "fill": torch.empty((B,) + tuple(fill.shape), dtype=torch.uint8, device=dev),
                             "meta": [], "free": None} for _ in range(2)]
        st = self._stages[self._cur]
        cur = torch.cuda.current_stream(dev)
        if not st["meta"] and st["free"] is not None:
            cur.wait_event(st["free"])              # the batch that used this staging set has left the device
        k = len(st["meta"])
        st["flow"][k].copy_(flow)
        st["frame"][k].copy_(frame)
        st["fill"][k].copy_(fill)
        st["meta"].append((name, src_np))
        if len(st["meta"]) == B:
            self.flush()

    def flush(self):
        """device_inpaint: fill and hand over the samples staged so far (a partial batch)"""
        if not self._stages or not self._stages[self._cur]["meta"]:
            return
        from utils import utils as U
        st = self._stages[self._cur]
        n = len(st["meta"])
        dev = st["frame"].device
        self._copy_stream.wait_stream(torch.cuda.current_stream(dev))
        slots = [self._slot([(tuple(st["flow"].shape[1:]), st["flow"].dtype), (tuple(st["frame"].shape[1:]), torch.uint8)])
                 for _ in range(n)]
        with torch.cuda.stream(self._copy_stream):
            frames = st["frame"][:n]
            U.inpaint_ns(frames, st["fill"][:n], 3, out=frames)                     # utils/utils.py:284-286, in place
            for k, slot in enumerate(slots):
                slot[0].copy_(st["flow"][k], non_blocking=True)
                slot[1].copy_(frames[k], non_blocking=True)
            event = torch.cuda.Event()
            event.record(self._copy_stream)
        st["free"] = event
        for (name, src_np), slot in zip(st["meta"], slots):
            self._jobs.put((name, slot, event, src_np, True))
        st["meta"] = []
        self._cur ^= 1

    def _work(self):
        cv2 = self._cv2
        while True:
            job = self._jobs.get()
            if job is None:
                return
            slot = None
            try:
                name, slot, event, src_np, filled = job
                if event is not None:
                    event.synchronize()
                if filled:                                                                   # done on the device
                    flow, inpainted = (t.numpy() for t in slot)
                else:
                    flow, frame, fill = (t.numpy() for t in slot)
                    inpainted = cv2.inpaint(frame, fill, 3, cv2.INPAINT_NS)                  # utils/utils.py:284-286
                writeFlow(os.path.join(self.out_dir, "flows", f"{name}.flo"), flow)           # gen...v2.py:120
                # imwrite reports a full disk / a missing directory / an encoder failure by returning False
                for sub, img in (("dst_images", inpainted), ("src_images", src_np)):
                    path = os.path.join(self.out_dir, sub, f"{name}.png")
                    if not cv2.imwrite(path, img):
                        raise OSError(f"cv2.imwrite failed for {path}")
            except Exception as e:      # surfaced by the next submit() / close()
                self._errors.append(e)
            finally:
                if slot is not None:
                    self._free.put(slot)
                self._jobs.task_done()

    def close(self):
        """wait for everything submitted, stop the workers, re-raise the first error"""
        self.flush()
        self._jobs.join()
        for _ in self._threads:
            self._jobs.put(None)
        for t in self._threads:
            t.join()
        if self._errors:
            raise self._errors[0]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
