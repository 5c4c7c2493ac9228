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
        # the preallocated gatherer: two steps through two slots, five packed channels, one collective per step
        g = shard.FrameGatherer(n_frames, 3, 4, "cpu")
        packed = []
        for step in range(2):
            for c, o in enumerate(g.outputs(step)):
                for i, f in enumerate(range(lo, hi)):
                    o[i] = 100 * step + 10 * c + f
            g.start(step)
        for step in range(2):
            g.wait(step)
            packed.append([torch.cat(per_rank, 0).clone() for per_rank in g.views(step)])
        if rank == 0:
            out_q.put((whole, slowest, packed))
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
    whole, slowest, packed = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert whole.shape == (n_frames, 2, 3, 4)
    for f in range(n_frames):
        assert torch.equal(whole[f, 0], torch.full((3, 4), float(f)))
        assert torch.equal(whole[f, 1], torch.full((3, 4), float(f) + 1))
    assert slowest == 2.0
    for step in range(2):
        for c, width in enumerate((3, 1, 1, 2, 1)):
            t = packed[step][c]
            assert t.shape == (n_frames, width, 3, 4)
            for f in range(n_frames):
                assert torch.equal(t[f], torch.full((width, 3, 4), float(100 * step + 10 * c + f)))
    assert slowest == 2.0


def _writer_worker(rank, world, port, n_samples, out_dir, golden):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "mpi-flow_b200")]
    import numpy as np
    import shard
    from sample_writer import SampleWriter
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = np.load(golden)
        lo, hi = shard.shard_range(n_samples, rank, world)
        with SampleWriter(out_dir, workers=2, depth=2) as w:      # every rank writes its own samples, no exchange
            for k in range(lo, hi):
                w.submit(f"demo_{k}", g["flow_mix"] + np.float32(k), g["frame_mix"], g["fill_mask"], g["src_np"])
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_write_disjoint_samples(tmp_path, golden_dir):
    """dataset generation at N>1: samples are sharded like frames, each rank's writer produces its own files"""
    pytest.importorskip("cv2")
    import numpy as np
    sys.path[:0] = [os.path.join(ROOT, "mpi-flow_b200")]
    import write_flow
    golden = os.path.join(golden_dir, "sample_pair.npz")
    n = 5
    ctx = mp.get_context("spawn")
    port = 29500 + (os.getpid() % 400) + 17
    procs = [ctx.Process(target=_writer_worker, args=(r, 2, port, n, str(tmp_path), golden)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    g = np.load(golden)
    assert sorted(os.listdir(tmp_path / "flows")) == sorted(f"demo_{k}.flo" for k in range(n))
    assert sorted(os.listdir(tmp_path / "dst_images")) == sorted(f"demo_{k}.png" for k in range(n))
    for k in range(n):
        assert np.array_equal(write_flow.readFlow(str(tmp_path / "flows" / f"demo_{k}.flo")), g["flow_mix"] + np.float32(k))
