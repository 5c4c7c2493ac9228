"""Batch sharding of the renderer over the GPUs of one node.

Frames are independent (no halo, no exchange step): rank r renders a contiguous slice of the batch
on its own B200; the only collective is the optional gather of the 8 output channels
(32*H*W bytes per frame) and one max-reduce of the elapsed time.  The reference has no in-process
equivalent: it runs one generator process per CUDA_VISIBLE_DEVICES (scripts/gen_train_kitti15_v2.sh:1).
"""
import torch
import torch.distributed as dist


def shard_range(n_frames, rank, world):
    """Contiguous, balanced [lo, hi) slice of range(n_frames) for `rank` of `world`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_frames(local, n_frames, group=None):
    """all_gather per-rank output slices [n_local, C, H, W] back into batch order [n_frames, C, H, W].

    Works for ragged shards (n_frames not divisible by the world size) by padding to the largest
    shard; backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
    world = dist.get_world_size(group)
    sizes = [shard_range(n_frames, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


def max_over_ranks(value, device, group=None):
    """Elapsed time of the slowest rank (device timing)."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
