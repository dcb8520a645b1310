"""Multi-GPU: frames are independent, so a batch shards across ranks with NO collective inside
forward / decode / NMS.  The only exchange is the gather of the
variable-length detection rows ``[D_r, 8]`` to rank 0, concatenated in rank order with the
image column shifted by the rank's first frame -- bit-identical to ``write_results`` on the
whole batch.  The reference's counterpart is ``nn.DataParallel`` (detect.py:177-183).

One process per GPU (torchrun); ``torch.distributed`` with NCCL over NVLink on the GPU box,
gloo in the CPU tests (the logic here is backend-agnostic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int):
    """Frames [lo, hi) of rank ``rank``: contiguous, sizes differ by at most one."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_detections(local_rows, first_frame: int, group=None, dst: int = 0):
    """Gather per-rank ``write_results`` outputs (tensor ``[D_r, 8]`` or int 0) to ``dst``.

    Returns, on ``dst``, the concatenated ``[D, 8]`` tensor (image index made global) or 0 when no
    rank has a detection; on other ranks returns None.  Two collectives: an all_gather of the
    counts (one int64 per rank) and a padded gather of the rows.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    if isinstance(local_rows, int):
        rows = torch.zeros(0, 8, dtype=torch.float32, device=dev)
    else:
        rows = local_rows.to(dev, torch.float32).clone()
        rows[:, 0] += float(first_frame)
    count = torch.tensor([rows.size(0)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(counts)
    if cap == 0:
        return 0 if rank == dst else None
    padded = torch.zeros(cap, 8, dtype=torch.float32, device=dev)
    padded[:rows.size(0)] = rows
    if rank == dst:
        bucket = [torch.empty_like(padded) for _ in range(world)]
        dist.gather(padded, bucket, dst=dst, group=group)
        return torch.cat([b[:c] for b, c in zip(bucket, counts)], 0)
    dist.gather(padded, None, dst=dst, group=group)
    return None


def detect_sharded(model, frames: torch.Tensor, num_class: int, confidence: float, nms_conf: float,
                   group=None):
    """Run forward + write_results on this rank's frames of the GLOBAL batch ``frames`` (every
    rank passes the same tensor or at least its own slice semantics via shard_bounds) and gather
    the detections to rank 0."""
    from .util import write_results
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(frames.size(0), world, rank)
    local = 0
    if hi > lo:
        pred = model(frames[lo:hi])
        local = write_results(pred, num_class, confidence, nms_conf)
    return gather_detections(local, lo, group=group)
