"""N>1 host logic on CPU: two gloo ranks shard a batch, 'render' their slice, gather it back."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_the_batch():
    import shard
    for n in (0, 1, 5, 8, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def _worker(rank, world, port, n_frames, out_q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "mpi-flow_b200")]
    import shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.shard_range(n_frames, rank, world)
        # stand-in for the renderer: frame f -> a [2, 3, 4] block filled with f (+ channel index)
        local = torch.stack([torch.full((2, 3, 4), float(f)) + torch.arange(2.).view(2, 1, 1) for f in range(lo, hi)]) \
            if hi > lo else torch.zeros((0, 2, 3, 4))
        whole = shard.gather_frames(local, n_frames)
        slowest = shard.max_over_ranks(1.0 + rank, torch.device("cpu"))
        if rank == 0:
            out_q.put((whole, slowest))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [4, 5])
def test_two_ranks_gather_in_batch_order(n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + n_frames
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    whole, slowest = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert whole.shape == (n_frames, 2, 3, 4)
    for f in range(n_frames):
        assert torch.equal(whole[f, 0], torch.full((3, 4), float(f)))
        assert torch.equal(whole[f, 1], torch.full((3, 4), float(f) + 1))
    assert slowest == 2.0
