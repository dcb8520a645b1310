"""Detection utilities: drop-in mirror of the reference's ``src/util.py`` hot-path functions.

Same names, argument order, defaults and return conventions as the reference
(``predict_transform`` src/util.py:175-239, ``write_results`` :242-346, ``bbox_iou``
:120-153, ``confidence_mask`` :106-117); the arithmetic runs in the sm_100a kernels of
``librtod.so`` through its C ABI.  Inputs are never modified, outputs are fresh tensors.

Tensors that live on the host are staged to the current CUDA device (pinned, asynchronous
where possible) and the result is returned on the device the input came from -- there is
no CPU implementation here.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_WORKSPACES: dict = {}


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("realtimeobjectdetection_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _to_device(t: torch.Tensor):
    """fp32 contiguous CUDA view/copy of ``t`` plus the device results should return to."""
    home = t.device
    if t.is_cuda:
        dev = t.device
    else:
        _require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
    t = t.detach()
    if t.dtype != torch.float32 or not t.is_cuda:
        t = t.to(device=dev, dtype=torch.float32, non_blocking=True)
    return t.contiguous(), dev, home


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = (device.type, device.index)
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = buf
    return buf


def confidence_mask(tensor: torch.Tensor, confidence: float) -> torch.Tensor:
    """``tensor * (tensor[:, :, 4] > confidence)`` -- src/util.py:106-117."""
    lib = _lib.load()
    x, dev, home = _to_device(tensor)
    out = torch.empty_like(x)
    rows = x.numel() // x.shape[-1] if x.numel() else 0
    with torch.cuda.device(dev):
        _lib.check(lib.rtod_confidence_mask(x.data_ptr(), rows, x.shape[-1], float(confidence),
                                            out.data_ptr(), _stream_ptr(dev)))
    return out.to(home)


def bbox_iou(box1: torch.Tensor, box2: torch.Tensor) -> torch.Tensor:
    """Broadcasting IoU with the +1 pixel convention -- src/util.py:120-153."""
    lib = _lib.load()
    home = box1.device
    a, dev, _ = _to_device(box1)
    b = box2.detach().to(device=dev, dtype=torch.float32)
    lead = torch.broadcast_shapes(a.shape[:-1], b.shape[:-1])
    n_out = 1
    for s in lead:
        n_out *= s
    out = torch.empty(lead, dtype=torch.float32, device=dev)
    if n_out == 0:
        return out.to(home)

    def flat(t):
        n = t.numel() // t.shape[-1]
        if n == 1 or n == n_out:
            return t.contiguous().view(n, t.shape[-1]), n
        t = t.expand(*lead, t.shape[-1]).contiguous()
        return t.view(n_out, t.shape[-1]), n_out

    a2, n1 = flat(a)
    b2, n2 = flat(b)
    with torch.cuda.device(dev):
        _lib.check(lib.rtod_bbox_iou(a2.data_ptr(), n1, a2.shape[1], b2.data_ptr(), n2, b2.shape[1],
                                     out.data_ptr(), _stream_ptr(dev)))
    return out.to(home)


def predict_transform(prediction, inp_dim, anchors, num_class, CUDA, TRAIN=False) -> torch.Tensor:
    """One YOLO head, NCHW ``[B, A*(5+C), G, G]`` -> ``[B, G*G*A, 5+C]`` -- src/util.py:175-239.

    ``CUDA`` is accepted for signature compatibility; the computation is always on the GPU.
    """
    lib = _lib.load()
    x, dev, home = _to_device(prediction)
    if x.dim() != 4 or x.size(2) != x.size(3):
        raise ValueError("predict_transform expects [B, A*(5+C), G, G], got %s" % (tuple(x.shape),))
    batch, grid = x.size(0), x.size(2)
    n_anchor, attrs = len(anchors), 5 + int(num_class)
    if x.size(1) != n_anchor * attrs:
        raise RuntimeError("shape '[%d, %d, %d]' is invalid for input of size %d"
                           % (batch, attrs * n_anchor, grid * grid, x.numel()))
    flat = (ctypes.c_float * (2 * n_anchor))(*[float(v) for pair in anchors for v in pair[:2]])
    out = torch.empty(batch, grid * grid * n_anchor, attrs, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.rtod_yolo_decode(x.data_ptr(), batch, grid, n_anchor, int(num_class), int(inp_dim),
                                        flat, int(bool(TRAIN)), out.data_ptr(), _stream_ptr(dev)))
    return out.to(home)


class PendingDetections:
    """Handle of an enqueued ``write_results`` (see ``write_results_async``): the kernels and the copy of the
    detection count to pinned host memory are in the stream, nothing has been synchronised yet."""

    def __init__(self, rows, count_host, event, home, cap, device):
        self._rows, self._count_host, self._event = rows, count_host, event
        self._home, self._cap, self._device = home, cap, device
        self._d = None
        self.rows_device = rows              # [cap, 8]: the first count_device[0] rows are the detections
        self._host_rows = None               # pinned [K, 8]: speculative copy of the first K rows (write_results_async)
        self.count_device = None             # int32 [1] on the device (set by write_results_async; None: count on the host only)
        self._count_np = None

    def __del__(self):
        # dropped without result(): the kernels may still be about to write the pinned count word -- park it until
        # its event has completed instead of letting the host allocator hand it to somebody else
        try:
            if self._count_host is not None and len(_ORPHANS) < 1024:
                _ORPHANS.append((self._event, self._count_host, self._count_np))
        except Exception:
            pass

    def result(self, to_host: bool = False):
        """``[D, 8]`` rows or the int ``0`` (the reference's convention).  Waits only for the event recorded
        after this call's kernels -- work enqueued later on the stream (the next batch) keeps running.
        ``to_host``: copy the rows to the host on a side stream instead of the compute stream."""
        if self._count_host is not None:
            self._event.synchronize()
            self._d = int(self._count_np[0])                  # (numpy view: indexing the tensor costs microseconds)
            if len(_PINNED_COUNTS) < 64:
                _PINNED_COUNTS.append((self._count_host, self._count_np))    # recycled: no host allocation per call
            self._count_host = None
        d = self._d
        spec_out = None
        if self._host_rows is not None:
            # the first rows were copied to pinned memory in the stream, behind the kernels and before the event: they
            # are on the host already -- no second copy, no second synchronisation when the count fits
            buf, self._host_rows = self._host_rows, None
            if 0 < d <= buf.size(0):
                spec_out = buf[:d].clone()
            pool = _PINNED_ROWS.setdefault(buf.size(0), [])
            if len(pool) < 8:
                pool.append(buf)
        if d < 0:
            raise RuntimeError("write_results: device-side failure (a bounded wait in the NMS kernels timed out)")
        if d == 0:
            return 0
        if d > self._cap:
            raise RuntimeError("write_results: %d detections exceed capacity %d" % (d, self._cap))
        if to_host or self._home.type == "cpu":
            if spec_out is not None:
                return spec_out
            side = _side_stream(self._device)
            out = torch.empty(d, 8, dtype=torch.float32, pin_memory=True)
            with torch.cuda.stream(side):
                out.copy_(self._rows[:d], non_blocking=True)      # rows are final since `event`
            side.synchronize()
            return out
        # the row buffer is this call's own allocation: for small batches hand out a view of it (one launch less on the
        # batch-1 latency path); for large ones copy the few detections out and let the [B*N, 8] buffer go
        if self._rows.numel() * 4 <= (1 << 20):
            return self._rows[:d]
        return self._rows[:d].clone()


_SIDE_STREAMS = {}
_PINNED_COUNTS = []                                     # recycled (1-element pinned int32 tensor, its numpy view) pairs
_ORPHANS = []                                           # (event, pinned count, view) of handles dropped before result()
_PINNED_ROWS = {}                                       # K -> recycled pinned [K, 8] buffers of the speculative row copy


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[key]


def write_results_async(prediction, num_class, confidence=0.6, nms_conf=0.4, host_rows: int = 0,
                        device_count: bool = True) -> PendingDetections:
    """``write_results`` split in two: this call enqueues scan + per-image NMS + emit and returns at once;
    ``.result()`` yields what ``write_results`` returns.  A streaming loop enqueues the next batch's forward
    before collecting, so the GPU never waits for the host (``DetectionPipeline``)."""
    lib = _lib.load()
    x, dev, home = _to_device(prediction)
    if x.dim() != 3 or x.size(2) < 5 + int(num_class):
        raise ValueError("write_results expects [B, N, >=5+num_class], got %s" % (tuple(x.shape),))
    if x.size(2) != 5 + int(num_class):
        x = x[:, :, :5 + int(num_class)].contiguous()      # reference slices 5:5+num_class (:279)
    B, N, C = x.size(0), x.size(1), int(num_class)
    cap = max(B * N, 1)
    rows = torch.empty(cap, 8, dtype=torch.float32, device=dev)
    # (rtod_write_results always stores the count -- the last image's CTA writes it unconditionally -- so only the
    # empty-tensor path, which makes no call, needs a zero: one fill launch less on the batch-1 latency path)
    while _ORPHANS and _ORPHANS[0][0].query():             # parked count words whose kernels have finished
        _, ch, cn = _ORPHANS.pop(0)
        if len(_PINNED_COUNTS) < 64:
            _PINNED_COUNTS.append((ch, cn))
    if _PINNED_COUNTS:
        count_host, count_np = _PINNED_COUNTS.pop()
    else:
        count_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        count_np = count_host.numpy()
    # device_count=False (nobody reads the count on the device: no gather): the kernels store it straight into the
    # pinned word -- page-locked memory is device-addressable under unified addressing, like the plans' failure word --
    # instead of a device int that a copy node then brings over: one node less at the end of every step
    if device_count:
        count = torch.empty(1, dtype=torch.int32, device=dev) if B > 0 and N > 0 else torch.zeros(1, dtype=torch.int32, device=dev)
        count_ptr = count.data_ptr()
    else:
        count, count_ptr = None, count_host.data_ptr()
        count_np[0] = 0
    event = torch.cuda.Event()
    if B > 0 and N > 0:
        nbytes = lib.rtod_write_results_workspace_bytes(B, N, C)
        ws = _workspace(dev, nbytes + 256)
        ws_ptr = (ws.data_ptr() + 255) // 256 * 256
        with torch.cuda.device(dev):
            _lib.check(lib.rtod_write_results(x.data_ptr(), B, N, C, float(confidence), float(nms_conf),
                                              rows.data_ptr(), cap, count_ptr, ws_ptr, nbytes,
                                              _stream_ptr(dev)))
    if count is not None:
        count_host.copy_(count, non_blocking=True)
    spec = None
    if host_rows > 0 and B > 0 and N > 0:
        # the caller will want the rows on the host (``result(to_host=True)``): copy the first `host_rows` of them to
        # pinned memory now, in the stream -- when the count turns out to fit, the detections arrive with it
        k = min(int(host_rows), cap)
        pool = _PINNED_ROWS.get(k)
        spec = pool.pop() if pool else torch.empty(k, 8, dtype=torch.float32).pin_memory()
        spec.copy_(rows[:k], non_blocking=True)
    event.record(torch.cuda.current_stream(dev))
    pending = PendingDetections(rows, count_host, event, home, cap, dev)
    pending._count_np = count_np
    pending._host_rows = spec
    pending._keep = (x, count)                                 # alive until the kernels have run
    pending.count_device = count
    return pending


def write_results(prediction, num_class, confidence=0.6, nms_conf=0.4):
    """Threshold + per-image per-class greedy NMS -- src/util.py:242-346.

    Returns ``[D, 8]`` fp32 rows ``[img, x1, y1, x2, y2, obj, cls_conf, cls]`` ordered image,
    class ascending, objectness descending -- or the int ``0`` when nothing survives (callers
    test ``type(x) == int``).  Rows of one (image, class) with bit-equal objectness are ordered
    by row index (the reference's ``torch.sort`` leaves that order unspecified).
    """
    return write_results_async(prediction, num_class, confidence, nms_conf, device_count=False).result()


# ---------------------------------------------------------------------------------------------------
# callers either side of the hot path: pre-processing, box rescale, IoU matrix
# ---------------------------------------------------------------------------------------------------
RESIZE_FLOAT, RESIZE_OPENCV = 0, 1


def letterbox_geometry(img_w: int, img_h: int, inp_dim: int):
    """``(new_w, new_h, left, top)`` of ``letterbox_image`` (src/util.py:360-369) on a square canvas."""
    lib = _lib.load()
    out = [ctypes.c_int() for _ in range(4)]
    _lib.check(lib.rtod_letterbox_geometry(int(img_w), int(img_h), int(inp_dim), *[ctypes.byref(v) for v in out]))
    return tuple(v.value for v in out)


def _as_u8_frames(img):
    """uint8 ``[B, H, W, 3]`` tensor view of a numpy / torch HWC image or batch, and whether a batch axis was added."""
    t = torch.from_numpy(img) if not isinstance(img, torch.Tensor) else img
    if t.dtype != torch.uint8:
        raise TypeError("expected uint8 HWC frames (what cv2.imread returns), got %s" % (t.dtype,))
    single = t.dim() == 3
    if single:
        t = t.unsqueeze(0)
    if t.dim() != 4 or t.size(3) != 3:
        raise ValueError("expected [H, W, 3] or [B, H, W, 3], got %s" % (tuple(t.shape),))
    return t, single


def prep_frames(frames, inp_dim: int, mode: str = "BGR", resize: int = RESIZE_FLOAT, out: torch.Tensor = None,
                as_uint8: bool = False) -> torch.Tensor:
    """Batched, device-resident ``prep_image``: uint8 ``[B, H, W, 3]`` frames (host or device) ->
    ``[B, 3, inp_dim, inp_dim]`` fp32 (or uint8 planes) ON THE GPU, one kernel (src/util.py:349-397)."""
    assert mode in ("BGR", "RGB")
    lib = _lib.load()
    t, _ = _as_u8_frames(frames)
    if t.is_cuda:
        dev = t.device
    else:
        _require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        t = t.to(dev, non_blocking=True)
    t = t.contiguous()
    B, H, W = t.size(0), t.size(1), t.size(2)
    dtype = torch.uint8 if as_uint8 else torch.float32
    if out is None:
        out = torch.empty(B, 3, int(inp_dim), int(inp_dim), dtype=dtype, device=dev)
    elif out.shape != (B, 3, int(inp_dim), int(inp_dim)) or out.dtype != dtype or not out.is_contiguous():
        raise ValueError("prep_frames: `out` must be a contiguous %s [%d, 3, %d, %d] tensor" % (dtype, B, inp_dim, inp_dim))
    with torch.cuda.device(dev):
        _lib.check(lib.rtod_prep_image(t.data_ptr(), B, H, W, int(inp_dim), int(mode == "RGB"), int(resize),
                                       int(as_uint8), out.data_ptr(), _stream_ptr(dev)))
    out._rtod_keep = t                                            # the staged frames live until the kernel ran
    return out


def prep_image(img, inp_dim, mode="BGR", resize: int = RESIZE_FLOAT) -> torch.Tensor:
    """``prep_image(img, inp_dim, mode='BGR')`` -- src/util.py:375-397: letterbox (cubic resize, 128 fill),
    BGR->RGB, HWC->CHW, ``/255``; returns ``[1, 3, inp_dim, inp_dim]`` fp32 on the device the image came from
    (the host for a numpy image, like the reference)."""
    t, single = _as_u8_frames(img)
    if not single:
        raise ValueError("prep_image takes one [H, W, 3] image; use prep_frames for a batch")
    out = prep_frames(t, inp_dim, mode, resize)
    return out if t.is_cuda else out.cpu()


def letterbox_image(img, inp_dim, resize: int = RESIZE_FLOAT):
    """``letterbox_image(img, (w, h))`` -- src/util.py:349-372 for square canvases; returns the ``[h, w, 3]``
    canvas (channel order untouched) as an int64 numpy array like the reference."""
    w, h = inp_dim
    if w != h:
        raise ValueError("only square canvases are supported (the reference always passes (inp_dim, inp_dim))")
    t, single = _as_u8_frames(img)
    planes = prep_frames(t, w, "RGB", resize, as_uint8=True)            # "RGB" = keep the channel order
    canvas = planes[0].permute(1, 2, 0).contiguous().cpu().numpy().astype("int64")
    return canvas


def rescale_boxes(output: torch.Tensor, im_dim_list: torch.Tensor, inp_dim: int, ref_dim: int = 416):
    """detect.py:120-136 on the device: map the boxes of ``write_results`` rows from letterbox to source-image
    pixels (scale ``min(ref_dim / w, ref_dim / h)``: the reference hard-codes 416 at detect.py:130 -- pass
    ``ref_dim=inp_dim`` for other resolutions) and clamp them to the image.  ``im_dim_list``: ``[n_img, 4]``
    rows ``(w, h, w, h)``.  Returns ``(rows, im_dim_rows)``; inputs are not modified."""
    lib = _lib.load()
    x, dev, home = _to_device(output)
    dims = im_dim_list.detach().to(device=dev, dtype=torch.float32).contiguous()
    if x.dim() != 2 or x.size(1) != 8 or dims.dim() != 2 or dims.size(1) != 4:
        raise ValueError("rescale_boxes expects [D, 8] rows and [n_img, 4] sizes")
    out = torch.empty_like(x)
    out_dims = torch.empty(x.size(0), 4, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.rtod_rescale_boxes(x.data_ptr(), x.size(0), dims.data_ptr(), dims.size(0), int(inp_dim),
                                          int(ref_dim), out.data_ptr(), out_dims.data_ptr(), _stream_ptr(dev)))
    return out.to(home), out_dims.to(home)


def metrics_rows(prediction):
    """What the reference stores in ``metrics.json`` per image name (detect.py:107, 155, 164): the
    ``write_results`` rows as nested lists, or the int 0."""
    return prediction if isinstance(prediction, int) else prediction.detach().cpu().tolist()


def bbox_iou_matrix(pred: torch.Tensor, target: torch.Tensor, threshold=None) -> torch.Tensor:
    """test.py:139-151 as one kernel: ``[P, T]`` matrix of ``bbox_iou(pred[i, 1:5], target[j, 0:4])``;
    with ``threshold`` the entries that are not ``> threshold`` are zeroed like the validator does."""
    lib = _lib.load()
    p, dev, home = _to_device(pred)
    t = target.detach().to(device=dev, dtype=torch.float32).contiguous()
    if p.dim() != 2 or t.dim() != 2 or p.size(1) < 5 or t.size(1) < 4:
        raise ValueError("bbox_iou_matrix expects pred [P, >=5] (boxes in columns 1:5) and target [T, >=4]")
    out = torch.empty(p.size(0), t.size(0), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.rtod_bbox_iou_matrix(p.data_ptr() + 4, p.size(0), p.size(1), t.data_ptr(), t.size(0), t.size(1),
                                            int(threshold is not None), float(threshold or 0.0), out.data_ptr(),
                                            _stream_ptr(dev)))
    return out.to(home)
