"""Batch sharding of the renderer over the GPUs of one node.

Frames are independent (no halo, no exchange step): rank r renders a contiguous slice of the batch
on its own B200; the only collective is the optional gather of the 8 output channels
(32*H*W bytes per frame) and one max-reduce of the elapsed time.  The reference has no in-process
equivalent: it runs one generator process per CUDA_VISIBLE_DEVICES (scripts/gen_train_kitti15_v2.sh:1).

Output gather.  `FrameGatherer` owns, per slot, ONE flat send buffer that the renderer's five outputs are views of
(rgb | depth | mask | flow | obj, each dense) and one receive buffer of `world` such blocks, so that the render of
step i+1 (which writes the other slot) overlaps the gather of step i.  Nothing is allocated, padded or concatenated
per step; `views(slot)` hands out the gathered frames as per-rank views.  Two transports:

  * "p2p" (CUDA, peers reachable over NVLink): the send buffers live in symmetric memory
    (torch.distributed._symmetric_memory); after a device-side barrier every rank PULLS the other ranks' blocks straight
    out of their memory with copy-engine transfers, one side stream per peer.  No SM is used, so the gather runs under
    the persistent renderer, which occupies every SM for the whole launch.
  * "push" (same requirements): the RECEIVE buffers live in symmetric memory instead and every rank writes its block
    into the peers' receive buffers (posted NVLink writes instead of read round trips), between the same two barriers.
  * "nccl" (and gloo in the CPU tests): a single `all_gather_into_tensor` on a side stream.  NCCL's kernel needs SMs;
    pass MPIFLOW_RESERVE_SMS(n) to the renderer to leave it some.
"""
import torch
import torch.distributed as dist

CHANNELS = (3, 1, 1, 2, 1)     # rgb, depth, mask (validity count), flowA2B, object mask


def shard_range(n_frames, rank, world):
    """Contiguous, balanced [lo, hi) slice of range(n_frames) for `rank` of `world`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class FrameGatherer:
    """Preallocated, overlapped all-gather of the renderer's outputs.

        g = FrameGatherer(n_frames, H, W, device)          # n_frames = whole batch; this rank renders its shard
        out = g.outputs(slot)                              # 5 dense tensors [n_local, c, H, W] to render into
        ... render into out on the current stream ...
        g.start(slot)                                      # single collective on the side stream
        ... render the next step into the other slot ...
        g.wait(slot); per_rank = g.views(slot)             # per channel: list over ranks of [n_r, c, H, W] views
    """

    def __init__(self, n_frames, H, W, device, slots=2, group=None, channels=CHANNELS, dtype=torch.float32, mode="auto"):
        self.group, self.device = group, torch.device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_frames, self.H, self.W, self.channels = n_frames, H, W, tuple(channels)
        self.spans = [shard_range(n_frames, r, self.world) for r in range(self.world)]
        self.longest = max(hi - lo for lo, hi in self.spans)
        self.n_local = self.spans[self.rank][1] - self.spans[self.rank][0]
        self.block = self.longest * sum(self.channels) * H * W            # elements per rank (ragged shards padded)
        self.cuda = self.device.type == "cuda"
        self.side = torch.cuda.Stream(self.device) if self.cuda else None
        self.done = [None] * slots
        self.mode, self.why_not_p2p = "nccl" if self.cuda else dist.get_backend(group), None
        self.send = None
        self.recv = None
        if self.cuda and mode in ("auto", "p2p", "push") and self.world > 1:
            try:
                if mode == "push":
                    self._setup_push(slots, dtype)
                else:
                    self._setup_p2p(slots, dtype)
                self.mode = "push" if mode == "push" else "p2p"
            except Exception as e:  # noqa: BLE001 -- no symmetric memory / no peer access: the collective still works
                if mode in ("p2p", "push"):
                    raise
                self.why_not_p2p = f"{type(e).__name__}: {e}"[:200]
        if self.send is None:
            self.send = [torch.zeros(self.block, dtype=dtype, device=self.device) for _ in range(slots)]
        if self.recv is None:
            self.recv = [torch.empty(self.world * self.block, dtype=dtype, device=self.device) for _ in range(slots)]

    def _setup_push(self, slots, dtype):
        import torch.distributed._symmetric_memory as symm_mem
        grp = self.group if self.group is not None else dist.group.WORLD
        recv, handles = [], []
        for _ in range(slots):
            t = symm_mem.empty(self.world * self.block, dtype=dtype, device=self.device)
            t.zero_()
            handles.append(symm_mem.rendezvous(t, grp))
            recv.append(t)
        # every peer's receive buffer of every slot, as tensors in this process's address space (NVLink stores)
        self.peer_recv = [[h.get_buffer(r, (self.world * self.block,), dtype) for r in range(self.world)] for h in handles]
        self.handles = handles
        self.pull_streams = [torch.cuda.Stream(self.device) for _ in range(self.world - 1)]
        torch.cuda.synchronize(self.device)
        handles[0].barrier()
        torch.cuda.synchronize(self.device)
        self.recv = recv
        self.send = [torch.zeros(self.block, dtype=dtype, device=self.device) for _ in range(slots)]

    def _setup_p2p(self, slots, dtype):
        import torch.distributed._symmetric_memory as symm_mem
        grp = self.group if self.group is not None else dist.group.WORLD
        send, handles = [], []
        for _ in range(slots):
            t = symm_mem.empty(self.block, dtype=dtype, device=self.device)
            t.zero_()
            handles.append(symm_mem.rendezvous(t, grp))
            send.append(t)
        # every peer's block of every slot, as tensors in this process's address space (NVLink loads / copy-engine reads)
        self.peer = [[h.get_buffer(r, (self.block,), dtype) for r in range(self.world)] for h in handles]
        self.handles = handles
        self.pull_streams = [torch.cuda.Stream(self.device) for _ in range(self.world - 1)]
        torch.cuda.synchronize(self.device)
        handles[0].barrier()
        torch.cuda.synchronize(self.device)
        self.send = send

    def _split(self, flat, n):
        """the five dense per-channel tensors inside one rank's flat block (block layout: rgb | depth | ...)"""
        out, off = [], 0
        for c in self.channels:
            size = self.longest * c * self.H * self.W
            out.append(flat[off: off + size].view(self.longest, c, self.H, self.W)[:n])
            off += size
        return out

    def outputs(self, slot):
        return self._split(self.send[slot], self.n_local)

    def start(self, slot):
        """gather slot's send buffer into every rank's receive buffer; returns immediately on CUDA"""
        if not self.cuda:
            dist.all_gather_into_tensor(self.recv[slot], self.send[slot], group=self.group)
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        if self.mode == "push":
            h, mine = self.handles[slot], self.send[slot]
            with torch.cuda.stream(self.side):
                self.side.wait_event(ready)
                h.barrier()                      # every rank is done with the previous content of its receive buffer
                fence = torch.cuda.Event()
                fence.record(self.side)
            pushed = []
            for k, st in enumerate(self.pull_streams):      # one copy-engine stream per peer
                r = (self.rank + 1 + k) % self.world
                with torch.cuda.stream(st):
                    st.wait_event(fence)
                    self.peer_recv[slot][r][self.rank * self.block: (self.rank + 1) * self.block].copy_(mine, non_blocking=True)
                    e = torch.cuda.Event()
                    e.record(st)
                    pushed.append(e)
            with torch.cuda.stream(self.side):
                for e in pushed:
                    self.side.wait_event(e)
                h.barrier()                      # every rank's writes have landed: the receive buffers are complete
                self.done[slot] = torch.cuda.Event()
                self.done[slot].record(self.side)
            return
        if self.mode == "p2p":
            h, recv = self.handles[slot], self.recv[slot]
            with torch.cuda.stream(self.side):
                self.side.wait_event(ready)
                h.barrier()                      # every rank's block of this slot is rendered
                fence = torch.cuda.Event()
                fence.record(self.side)
            pulled = []
            for k, st in enumerate(self.pull_streams):      # one copy-engine stream per peer, nearest neighbour first
                r = (self.rank + 1 + k) % self.world
                with torch.cuda.stream(st):
                    st.wait_event(fence)
                    recv[r * self.block: (r + 1) * self.block].copy_(self.peer[slot][r], non_blocking=True)
                    e = torch.cuda.Event()
                    e.record(st)
                    pulled.append(e)
            with torch.cuda.stream(self.side):          # (this rank's own block stays where it was rendered: views())
                for e in pulled:
                    self.side.wait_event(e)
                h.barrier()                      # nobody still reads this rank's block: it may be rendered into again
                self.done[slot] = torch.cuda.Event()
                self.done[slot].record(self.side)
            return
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            dist.all_gather_into_tensor(self.recv[slot], self.send[slot], group=self.group)
            self.done[slot] = torch.cuda.Event()
            self.done[slot].record(self.side)

    def wait(self, slot):
        """the current stream waits for slot's gather (before its buffers are reused or its views are read)"""
        if self.cuda and self.done[slot] is not None:
            torch.cuda.current_stream(self.device).wait_event(self.done[slot])
            self.done[slot] = None

    def views(self, slot):
        """per channel, the list over ranks of that rank's frames [n_r, c, H, W] (views of the receive buffer)"""
        own = self.send[slot] if self.mode in ("p2p", "push") else None     # peer memory moves the peers' blocks only
        per_rank = [self._split(own if (own is not None and r == self.rank) else self.recv[slot][r * self.block: (r + 1) * self.block],
                                hi - lo) for r, (lo, hi) in enumerate(self.spans)]
        return [[per_rank[r][c] for r in range(self.world)] for c in range(len(self.channels))]


def gather_frames(local, n_frames, group=None):
    """all_gather per-rank output slices [n_local, C, H, W] back into batch order [n_frames, C, H, W] (one
    `all_gather_into_tensor`; ragged shards are padded to the largest).  Convenience form for one tensor; the
    benchmark and anything per-step use FrameGatherer, which allocates nothing per call."""
    world = dist.get_world_size(group)
    spans = [shard_range(n_frames, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in spans)
    tail = tuple(local.shape[1:])
    send = local.new_zeros((longest,) + tail)
    send[: local.shape[0]] = local
    recv = local.new_empty((world * longest,) + tail)
    dist.all_gather_into_tensor(recv, send, group=group)
    if all(hi - lo == longest for lo, hi in spans):
        return recv
    return torch.cat([recv[r * longest: r * longest + hi - lo] for r, (lo, hi) in enumerate(spans)], dim=0)


def max_over_ranks(value, device, group=None):
    """Elapsed time of the slowest rank (device timing)."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
