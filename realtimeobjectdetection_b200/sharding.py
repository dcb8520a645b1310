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
    counts (one int64 per rank) and a padded gather of the rows -- exact for any size, but the counts
    are read on the host; a streaming loop uses ``gather_detections_async`` (one fixed-capacity
    collective, no host synchronisation) instead.
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


class PendingGather:
    """Handle of an enqueued fixed-capacity gather (``gather_detections_async``)."""

    def __init__(self, rank, dst, world, capacity, bucket_host, done, work):
        self._rank, self._dst, self._world, self._cap = rank, dst, world, capacity
        self._bucket_host, self._done, self._work = bucket_host, done, work
        self.bucket_bytes = 0 if bucket_host is None else bucket_host.numel() * 4      # device -> host bytes on dst

    def result(self):
        """On ``dst``: the concatenated ``[D, 8]`` rows (image index global) or 0; elsewhere None.  Waits only for
        the event recorded behind this gather's own copy -- later work on the compute stream keeps running."""
        if self._work is not None:                           # CPU (gloo) path
            self._work.wait()
        if self._done is not None:
            # every rank waits for ITS part of the collective (one step late in a streaming loop): a rank that never
            # collects anything would run its host arbitrarily far ahead of its GPU, the caching allocator could not
            # recycle the per-step buffers and its cudaMalloc / cudaFree calls stalled the whole job for 40-120 ms
            self._done.synchronize()
        if self._rank != self._dst:
            return None
        parts = []
        for r in range(self._world):
            n = int(self._bucket_host[r, 0, 0])
            if n > self._cap:
                raise RuntimeError("rank %d produced %d detections, gather capacity is %d rows per rank: "
                                   "raise `capacity`" % (r, n, self._cap))
            if n:
                parts.append(self._bucket_host[r, 1:1 + n])
        out = torch.cat(parts, 0).clone() if parts else 0
        if self._done is not None and self._bucket_host is not None:     # the pinned bucket goes back to the pool
            _PINNED_BUCKETS.setdefault(tuple(self._bucket_host.shape), []).append(self._bucket_host)
            self._bucket_host = None
        return out


_GATHER_STREAMS = {}
# pinned host buckets by shape, reused from step to step: a fresh page-locked allocation per step costs a cudaHostAlloc
# whenever the caching host allocator has no retired block yet -- 15-120 ms stalls of every rank (they wait in the
# collective) in the first steps of a streaming loop on an 8-GPU box
_PINNED_BUCKETS = {}


def gather_detections_async(rows, count, first_frame: int, capacity: int, group=None, dst: int = 0) -> PendingGather:
    """One fixed-capacity collective per step, no host synchronisation on the way in.

    ``rows``: this rank's ``[>= D_r, 8]`` detection buffer (``PendingDetections.rows_device``: only the first
    ``count`` rows are meaningful), ``count``: int32 tensor holding ``D_r`` (device-resident: it is never read
    on the host here).  Every rank contributes ``capacity + 1`` rows: row 0 carries the count in-band, rows
    1.. the detections with the image column shifted by ``first_frame``.  With NCCL the gather and the copy of
    the bucket to pinned host memory run on a side stream behind an event, so the compute stream goes straight on
    to the next batch; ``.result()`` is meant to be called one step late."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nccl = dist.get_backend(group) == "nccl"
    dev = rows.device if nccl else torch.device("cpu")
    n_rows = min(capacity, rows.size(0))
    if nccl and rows.is_cuda and rows.dtype == torch.float32 and rows.is_contiguous() and count.is_cuda and \
            count.dtype == torch.int32 and rows.data_ptr() % 16 == 0:
        # one launch (rtod_pack_detections) instead of the eight small tensor operations below, every step on every rank
        from . import _lib
        payload = torch.empty(capacity + 1, 8, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().rtod_pack_detections(rows.data_ptr(), n_rows, count.data_ptr(), float(first_frame), capacity,
                                                        payload.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    else:
        payload = torch.zeros(capacity + 1, 8, dtype=torch.float32, device=dev)
        cnt = count.to(dev).reshape(-1)[:1]
        payload[1:1 + n_rows] = rows[:n_rows].to(dev)
        payload[1:, 0] += float(first_frame)
        valid = torch.arange(capacity, device=dev).unsqueeze(1) < cnt     # rows beyond the count: stale buffer content
        payload[1:] *= valid
        payload[0, 0] = cnt.to(torch.float32)[0]                           # exact up to 2^24 rows
    bucket = torch.empty(world, capacity + 1, 8, dtype=torch.float32, device=dev) if rank == dst else None
    if not nccl:
        work = dist.gather(payload, list(bucket.unbind(0)) if rank == dst else None, dst=dst, group=group, async_op=True)
        return PendingGather(rank, dst, world, capacity, bucket, None, work)
    key = dev.index
    if key not in _GATHER_STREAMS:
        _GATHER_STREAMS[key] = torch.cuda.Stream(dev)
    side = _GATHER_STREAMS[key]
    ready = torch.cuda.Event()
    ready.record(torch.cuda.current_stream(dev))
    done, host = None, None
    with torch.cuda.stream(side):
        side.wait_event(ready)
        dist.gather(payload, list(bucket.unbind(0)) if rank == dst else None, dst=dst, group=group)
        payload.record_stream(side)
        if rank == dst:
            bucket.record_stream(side)
            pool = _PINNED_BUCKETS.get(tuple(bucket.shape))
            host = pool.pop() if pool else torch.empty(bucket.shape, dtype=torch.float32, pin_memory=True)
            host.copy_(bucket, non_blocking=True)
        done = torch.cuda.Event()
        done.record(side)
    return PendingGather(rank, dst, world, capacity, host, done, None)


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
