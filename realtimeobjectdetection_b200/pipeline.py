"""Streaming detection over host frames: pinned, asynchronous host->device copies on a side
stream overlap the previous batch's forward + decode + NMS (BASELINE.json configs[4]; the
reference's loop is detect.py:57-80 -- synchronous pageable ``.cuda()`` per image).

    pipe = DetectionPipeline(model, num_class=80, confidence=0.5, nms_conf=0.4)
    for det in pipe.run(host_batches):        # det: [D, 8] tensor on the host, or int 0
        ...

``host_batches`` yields either ``[B, 3, H, W]`` fp32 tensors (what ``prep_image`` returns) or **uint8
``[B, h, w, 3]`` BGR frames** as a camera / ``cv2.imread`` delivers them: those are copied as bytes (a
quarter of the fp32 traffic at network resolution) and letterboxed, channel-swapped and scaled by one kernel
on the device (``util.prep_frames``, src/util.py:349-397).
"""
from __future__ import annotations

import torch

from .util import prep_frames, write_results_async


class DetectionPipeline:
    def __init__(self, model, num_class: int, confidence: float = 0.6, nms_conf: float = 0.4,
                 device=None, depth: int = 2, collect_lag: int = 1, resize: int = 0, gather: dict = None):
        if not torch.cuda.is_available():
            raise RuntimeError("DetectionPipeline needs a CUDA device; there is no CPU fallback")
        self.model, self.num_class = model, int(num_class)
        self.confidence, self.nms_conf = float(confidence), float(nms_conf)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.depth = max(2, int(depth))
        # 1 (throughput): a batch's detections are awaited after the NEXT batch is in the stream, the GPU never
        # idles on the host.  0 (live video, per-frame latency): every batch is collected before the next frame
        # is requested from the source.
        self.collect_lag = 1 if collect_lag else 0
        self.resize = int(resize)                              # util.RESIZE_FLOAT / RESIZE_OPENCV for uint8 frames
        # multi-GPU (one process per GPU, frames sharded): {"first_frame": global index of this rank's frame 0 in a
        # batch, "capacity": rows per rank and step, "group": None, "dst": 0} -> every batch's detections are gathered
        # to `dst` with one fixed-capacity collective on a side stream (sharding.gather_detections_async); run() then
        # yields the global rows on `dst` and None on the other ranks
        self.gather = gather
        self.copy_stream = torch.cuda.Stream(self.device)
        self._slots = []
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._spec_rows = 256                                  # rows copied to the host speculatively with the count

    def _slot(self, k, like):
        while len(self._slots) <= k:
            self._slots.append({"buf": None, "x": None, "ready": torch.cuda.Event(), "free": torch.cuda.Event()})
        s = self._slots[k]
        if s["buf"] is None or s["buf"].shape != like.shape or s["buf"].dtype != like.dtype:
            s["buf"] = torch.empty(like.shape, dtype=like.dtype, device=self.device)
            s["x"] = None                                      # fp32 network input of uint8 slots, made on demand
            s["free"].record(torch.cuda.current_stream(self.device))
        return s

    def _stage(self, k, host):
        """enqueue the H2D copy of one host batch into ring slot k on the copy stream"""
        if host.dtype not in (torch.float32, torch.uint8):
            raise TypeError("DetectionPipeline takes fp32 [B,3,H,W] or uint8 [B,h,w,3] batches, got %s" % (host.dtype,))
        s = self._slot(k % self.depth, host)
        if not host.is_pinned():
            host = host.pin_memory()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(s["free"])           # previous user of the slot is done
            s["buf"].copy_(host, non_blocking=True)
            s["ready"].record(self.copy_stream)
        self.h2d_bytes += host.numel() * host.element_size()
        return s

    def _network_input(self, slot):
        """the slot's frames as the tensor the network reads (a stable buffer per slot, so the forward replays one
        CUDA graph per slot): fp32 [B, 3, D, D] as is; uint8 frames are letterboxed / channel-swapped on the device into
        uint8 planes, which the stem consumes directly (value / 255 folded into its weights)"""
        if slot["buf"].dtype == torch.float32:
            return slot["buf"]
        dim = int(self.model.net_info["height"])
        if slot["x"] is None or slot["x"].size(2) != dim:
            slot["x"] = torch.empty(slot["buf"].size(0), 3, dim, dim, dtype=torch.uint8, device=self.device)
        return prep_frames(slot["buf"], dim, "BGR", self.resize, out=slot["x"], as_uint8=True)

    def run(self, host_batches):
        """host_batches: iterable of host tensors (pinned memory avoids a staging copy).  Yields
        write_results() of every batch, on the host."""
        it = iter(host_batches)
        compute = torch.cuda.current_stream(self.device)
        try:
            pending = self._stage(0, next(it))
        except StopIteration:
            return
        k = 0
        prev = None                                            # detections of the previous batch, not yet collected
        borrow = getattr(self.model, "borrow_output", False)
        self.model.borrow_output = True                        # predictions are consumed in-stream: no copies around the graph
        try:
            while pending is not None:
                nxt = None
                try:
                    nxt = self._stage(k + 1, next(it))        # overlaps this batch's compute
                except StopIteration:
                    pass
                compute.wait_event(pending["ready"])
                pred = self.model(self._network_input(pending))
                # (no gather: the rows go to this host -- copy as many as the previous batch needed, rounded up, together
                # with the count, so that collecting them costs one synchronisation instead of two)
                handle = write_results_async(pred, self.num_class, self.confidence, self.nms_conf,
                                             host_rows=0 if self.gather is not None else self._spec_rows,
                                             device_count=self.gather is not None)
                pending["free"].record(compute)
                if self.gather is not None:
                    from .sharding import gather_detections_async
                    g = self.gather
                    handle = gather_detections_async(handle.rows_device, handle.count_device, g["first_frame"],
                                                     g["capacity"], g.get("group"), g.get("dst", 0))
                if self.collect_lag == 0:
                    yield self._collect(handle)
                else:
                    # this batch is in the stream: only now wait for the previous one (the GPU keeps working)
                    if prev is not None:
                        yield self._collect(prev)
                    prev = handle
                pending, k = nxt, k + 1
            if prev is not None:
                yield self._collect(prev)
        finally:
            self.model.borrow_output = borrow

    def _collect(self, handle):
        if self.gather is not None:
            det = handle.result()                              # PendingGather: host rows on dst, None elsewhere
            if det is not None:
                self.d2h_bytes += handle.bucket_bytes
            return det
        spec = getattr(handle, "_host_rows", None)
        spec_rows = 0 if spec is None else spec.size(0)          # rows copied speculatively (all of them cross the bus)
        det = handle.result(to_host=True)
        self.d2h_bytes += spec_rows * 32
        if not isinstance(det, int):
            if det.size(0) > spec_rows:
                self.d2h_bytes += det.numel() * 4                 # did not fit: the exact-size copy on top
            want = ((det.size(0) * 5 // 4 + 255) // 256) * 256          # 25 % head room, in steps of 256 rows
            if want > self._spec_rows or want * 4 < self._spec_rows:
                self._spec_rows = max(256, want)
        self.d2h_bytes += 4                                    # the detection count
        return det
