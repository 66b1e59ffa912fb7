"""Batch sharding of crops across ranks (one process per GPU) and the one terminal gather.

The pose path has no exchange step: rank r of W processes the contiguous
crop range [r*ceil(N/W), min(N, (r+1)*ceil(N/W))), and the per-crop results are gathered once.
Works with any torch.distributed backend (NCCL on GPUs; gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous range of crop indices owned by `rank` (may be empty for trailing ranks)."""
    per = -(-n // world) if n > 0 else 0
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def shard_frames(n_frames, rank, world):
    """Frame-level sharding: all boxes of a frame stay with the frame (no frame is copied twice)."""
    return shard_range(n_frames, rank, world)


def gather_rows(local, n_total, group=None):
    """all_gather of per-crop rows; returns the (n_total, ...) tensor in original crop order on every rank.

    `local` holds this rank's rows (shard_range order).  Shards are padded to equal length so a single
    all_gather_into_tensor suffices; pad rows are dropped after the gather.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    per = -(-n_total // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:n_total]


def run_sharded(fn, n_total, micro_batch, group=None, join=None):
    """Run fn(lo, hi) -> (hi-lo, ...) tensor over this rank's range in micro-batches and gather everything.
    `join` (optional) is called once all micro-batches are queued and before their results are concatenated - e.g.
    EnginePool.join when fn submits its work to side streams."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_range(n_total, rank, world)
    parts = [fn(s, min(hi, s + micro_batch)) for s in range(lo, hi, micro_batch)]
    if join is not None:
        join()
    local = torch.cat(parts) if parts else None
    if local is None:                       # empty shard: still take part in the collective
        probe = fn(0, 0)
        local = probe[:0]
    return gather_rows(local, n_total, group)
