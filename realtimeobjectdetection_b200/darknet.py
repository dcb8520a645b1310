"""``Darknet``: drop-in mirror of the reference's ``src/darknet.py`` model class.

Public surface kept identical (src/darknet.py:138-603): ``Darknet(cfg_file_path, CUDA)``,
``forward(x)`` -> ``[B, N, 5+C]`` fp32, ``load_weights(path)``, ``net_info`` (mutable; the
decode reads ``net_info["height"]`` at every call, :258), ``blocks``, ``module_list`` (same
module names, hence the same ``state_dict`` keys), ``header``, ``seen``, ``CUDA``, ``TRAIN``,
``train_mode()``, ``anchors`` / ``num_classes`` (set by forward), ``parse_cfg`` & friends.

What differs is the inside of ``forward``: instead of a Python loop of ATen modules it
replays a pre-planned launch sequence of hand-written sm_100a kernels (librtod.so, C ABI in
include/rtod.h).  The ``nn.Module`` parameters are the fp32 master copy; they are folded
(BatchNorm, inference statistics) and re-laid out on the device whenever they change.

Deliberate semantic choice: BatchNorm always uses its running statistics.  The reference's
scripts never call ``.eval()`` and therefore normalise with batch statistics at inference
time (SURVEY.md "fact 3"); folding BN into the convolution is only defined for running
statistics, and that is what a detector should do.  The parity oracle is ``.eval()``.
"""
from __future__ import annotations

import ctypes
import os
import warnings
from contextlib import contextmanager

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .cfg import parse_cfg as _parse_cfg_file


class MaxPoolStride1(nn.Module):
    """Same-size max-pool of tiny-YOLO (src/darknet.py:17-46); structural placeholder."""

    def __init__(self, kernel_size):
        super().__init__()
        self.kernel_size = kernel_size
        self.pad = kernel_size - 1

    def forward(self, x):
        x = F.pad(x, (0, self.pad, 0, self.pad), mode="replicate")
        return F.max_pool2d(x, self.kernel_size, self.pad)


class EmptyLayer(nn.Module):
    """Placeholder for route / shortcut blocks (src/darknet.py:49-54)."""


class DetectionLayer(nn.Module):
    """Holds the anchors of a yolo block (src/darknet.py:57-97)."""

    def __init__(self, anchors, CUDA=False):
        super().__init__()
        self.anchors = anchors
        self.CUDA = CUDA

    def forward(self, x, inp_dim, num_classes):
        from .util import predict_transform
        return predict_transform(x.data, inp_dim, self.anchors, num_classes, CUDA=self.CUDA)


_CUDA_SEEN = False
_EAGER_WEIGHT_CHECK = os.environ.get("RTOD_EAGER_WEIGHT_CHECK") == "1"     # A/B switch: version counters read before the launch


class _Plan:
    """A bound librtod execution plan for one input shape on one device."""

    def __init__(self, lib, descs, key, device, flags):
        batch, in_c, in_h, in_w, inp_dim = key
        self.lib, self.key, self.device = lib, key, device
        self.handle = ctypes.c_void_p()
        arr = (_lib.RtodLayerDesc * len(descs))(*descs)
        _lib.check(lib.rtod_plan_create(arr, len(descs), batch, in_c, in_h, in_w, inp_dim, flags,
                                        ctypes.byref(self.handle)))
        self.n_rows = lib.rtod_plan_num_rows(self.handle)
        self.n_attrs = lib.rtod_plan_num_attrs(self.handle)
        self.launches = lib.rtod_plan_launch_count(self.handle)
        self.conv_flops = lib.rtod_plan_conv_flops(self.handle)
        self.workspace = torch.empty(lib.rtod_plan_workspace_bytes(self.handle) + 256,
                                     dtype=torch.uint8, device=device)
        self.weights = torch.empty(lib.rtod_plan_weight_bytes(self.handle) + 256,
                                   dtype=torch.uint8, device=device)
        ws = (self.workspace.data_ptr() + 255) // 256 * 256
        wa = (self.weights.data_ptr() + 255) // 256 * 256
        with torch.cuda.device(device):
            _lib.check(lib.rtod_plan_bind(self.handle, ws, lib.rtod_plan_workspace_bytes(self.handle),
                                          wa, lib.rtod_plan_weight_bytes(self.handle)))
        self.is_f16 = bool(lib.rtod_plan_is_f16(self.handle))
        first = descs[0]
        # uint8 frames go straight into the tcgen05 stem (stem_tc.cu) when the plan has one
        self.takes_u8 = bool(self.is_f16 and in_c == 3 and first.type == _lib.LAYER_CONV and first.size == 3 and
                             first.stride == 1 and first.pad == 1 and first.filters in (16, 32, 64) and
                             in_w >= 160 and in_w % 16 == 0)
        # failure reporting without a sync: kernels store a code in this pinned, device-mapped int on time-out
        self.err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.err_np = self.err_host.numpy()                # (indexing a tensor costs microseconds on the batch-1 path)
        with torch.cuda.device(device):
            _lib.check(lib.rtod_plan_set_error_sink(self.handle, self.err_host.data_ptr(),
                                                    torch.cuda.current_stream(device).cuda_stream))
        self.weight_version = None                         # _weights_signature() of the last weight sync
        self.weight_epoch = None                           # ... and its cheap part (Darknet._weights_epoch)
        # CUDA graphs of the launch sequence, keyed by (input pointer, train): {"graph", "x", "pred"}.
        # Key (dtype, train) = the staging graph of that input type (input copied into its own buffer first).
        self.graphs = {}
        self.last_ptr = None                               # input pointer of the previous graph-path forward
        self.calls = 0
        self.last_used = 0

    def forward_fn(self, x):
        return self.lib.rtod_plan_forward_u8 if x.dtype == torch.uint8 else self.lib.rtod_plan_forward

    def raise_if_failed(self):
        """Raise (and re-arm the plan) if a kernel of an earlier forward reported a device-side failure."""
        code = int(self.err_np[0])
        if code == 0:
            return
        self.err_np[0] = 0
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            stream.synchronize()
            _lib.check(self.lib.rtod_plan_reset_errors(self.handle, stream.cuda_stream))
        self.graphs.clear()
        raise _lib.RtodError(-6, "device-side failure %d in an earlier forward (tcgen05/TMA pipeline time-out); "
                                 "its output is invalid; the plan has been re-armed" % code)

    def close(self):
        if self.handle:
            self.lib.rtod_plan_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Darknet(nn.Module):
    """YOLO Darknet model with the reference's interface (src/darknet.py:138-603)."""

    def __init__(self, cfg_file_path, CUDA):
        super().__init__()
        self.blocks = self.parse_cfg(cfg_file_path)
        self.net_info, self.module_list = self.create_modules(self.blocks)
        self.header = torch.IntTensor([0, 0, 0, 0])
        self.seen = 0
        self.CUDA = CUDA
        self.TRAIN = False
        # -- B200 runtime state (not part of the reference surface) --
        self._plans = {}
        self._plan_clock = 0
        self._weights_epoch = 0                    # bumped whenever parameters may have changed
        # _lib.PLAN_* flags: PLAN_KEEP_ALL / PLAN_CONV_SIMT for validation, PLAN_BF16 for bf16 storage
        # (default fp16: same tensor-core rate, 8x finer rounding), PLAN_NO_WSPLIT
        self.plan_flags = _lib.PLAN_BF16 if os.environ.get("RTOD_DTYPE", "fp16").lower() == "bf16" else 0
        self.use_cuda_graph = os.environ.get("RTOD_CUDA_GRAPH", "1") != "0"
        # True: forward() returns the replayed graph's own output buffer instead of a fresh copy -- valid until
        # the next forward with the SAME input buffer (DetectionPipeline / bench: the consumer is already
        # enqueued behind it on the stream).  False (default): reference semantics, every call a fresh tensor.
        self.borrow_output = False
        self._warned_train_bn = False
        self._warned_grad = False
        self._yolo_side = None                             # (anchors, classes) the yolo branch publishes, cached

    # ------------------------------------------------------------------ reference accessors
    def get_blocks(self) -> list:
        return self.blocks

    def get_module_list(self) -> nn.ModuleList:
        return self.module_list

    @contextmanager
    def train_mode(self):
        """``with model.train_mode():`` -> decode stops after the sigmoids (src/darknet.py:305-314)."""
        try:
            self.TRAIN = True
            yield
        finally:
            self.TRAIN = False

    # ------------------------------------------------------------------ cfg -> modules
    @staticmethod
    def parse_cfg(cfg_file_path):
        """cfg file -> list of dict blocks (src/darknet.py:412-447)."""
        return _parse_cfg_file(cfg_file_path)

    @staticmethod
    def create_modules(blocks):
        """blocks -> (net_info, nn.ModuleList) with the reference's module names
        (src/darknet.py:449-603), so ``state_dict()`` keys are interchangeable."""
        net_info = blocks[0]
        module_list = nn.ModuleList()
        prev_filters, filters = 3, 3
        output_filters = []
        for index, block in enumerate(blocks[1:]):
            module = nn.Sequential()
            kind = block["type"]
            if kind == "convolutional":
                try:
                    batch_normalize = int(block["batch_normalize"])
                    bias = False
                except (ValueError, KeyError):
                    batch_normalize, bias = 0, True
                filters = int(block["filters"])
                kernel_size = int(block["size"])
                pad = (kernel_size - 1) // 2 if int(block["pad"]) else 0
                module.add_module("conv_%d" % index,
                                  nn.Conv2d(prev_filters, filters, kernel_size, int(block["stride"]),
                                            pad, bias=bias))
                if batch_normalize:
                    module.add_module("batch_norm_%d" % index, nn.BatchNorm2d(filters))
                if block["activation"] == "leaky":
                    module.add_module("leaky_%d" % index, nn.LeakyReLU(0.1, inplace=True))
            elif kind == "upsample":
                int(block["stride"])                                  # parsed but ignored (:589)
                module.add_module("upsample_%d" % index,
                                  nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False))
            elif kind == "route":
                block["layers"] = block["layers"].split(",")        # the reference mutates too (:564)
                start = int(block["layers"][0])
                end = int(block["layers"][1]) if len(block["layers"]) > 1 else 0
                if start > 0:
                    start -= index
                if end > 0:
                    end -= index
                module.add_module("route_%d" % index, EmptyLayer())
                filters = output_filters[index + start] + (output_filters[index + end] if end < 0 else 0)
            elif kind == "shortcut":
                module.add_module("shortcut_%d" % index, EmptyLayer())
            elif kind == "maxpool":
                stride, size = int(block["stride"]), int(block["size"])
                module.add_module("maxpool_%d" % index,
                                  nn.MaxPool2d(size, stride) if stride != 1 else MaxPoolStride1(size))
            elif kind == "yolo":
                mask = [int(v) for v in block["mask"].split(",")]
                flat = [int(v) for v in block["anchors"].split(",")]
                pairs = [(flat[k], flat[k + 1]) for k in range(0, len(flat), 2)]
                module.add_module("Detection_%d" % index, DetectionLayer([pairs[m] for m in mask]))
            else:
                print("Unknown block error: A unknown block is provided")
                assert False
            module_list.append(module)
            prev_filters = filters
            output_filters.append(filters)
        return net_info, module_list

    # ------------------------------------------------------------------ weights
    def load_weights(self, weight_file_path: str):
        """Darknet binary ``.weights``: int32[5] header, then per conv block
        ``[bn.bias, bn.weight, bn.running_mean, bn.running_var]`` or ``[conv.bias]`` followed by
        the conv weight ``[Cout, Cin, k, k]`` (src/darknet.py:316-410)."""
        with open(weight_file_path, "rb") as fh:
            header = np.fromfile(fh, dtype=np.int32, count=5)
            stream = np.fromfile(fh, dtype=np.float32)
        self.header = torch.from_numpy(header)
        self.seen = self.header[3]
        pos = 0

        def fill(dst: torch.Tensor):
            nonlocal pos
            n = dst.numel()
            if pos + n > stream.size:
                raise RuntimeError("weights file %s is too short" % weight_file_path)
            dst.data.copy_(torch.from_numpy(stream[pos:pos + n]).view_as(dst))
            pos += n

        with torch.no_grad():
            for i, module in enumerate(self.module_list):
                if self.blocks[i + 1]["type"] != "convolutional":
                    continue
                conv = module[0]
                try:
                    batch_normalize = int(self.blocks[i + 1]["batch_normalize"])
                except KeyError:
                    batch_normalize = 0
                if batch_normalize:
                    bn = module[1]
                    fill(bn.bias)
                    fill(bn.weight)
                    fill(bn.running_mean)
                    fill(bn.running_var)
                else:
                    fill(conv.bias)
                fill(conv.weight)
        self._weights_epoch += 1

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._weights_epoch += 1
        return out

    def _apply(self, fn, *args, **kwargs):                   # .cuda() / .to() / .float() ...
        out = super()._apply(fn, *args, **kwargs)
        self._weights_epoch = getattr(self, "_weights_epoch", 0) + 1
        return out

    def refresh_weights(self):
        """Call after modifying parameters through ``.data`` (which bypasses version counters)."""
        self._weights_epoch += 1

    def _weights_signature(self):
        total = self._weights_epoch
        for t in self._weight_tensors():
            total += t._version
        return total

    def _weight_tensors(self):
        cached = getattr(self, "_weight_tensor_cache", None)
        if cached is None or cached[0] != self._weights_epoch:
            tensors = list(self.parameters()) + [b for b in self.buffers() if b.is_floating_point()]
            cached = (self._weights_epoch, tensors)
            self._weight_tensor_cache = cached
        return cached[1]

    # ------------------------------------------------------------------ plan
    def _layer_descs(self):
        descs = []
        for i, block in enumerate(self.blocks[1:]):
            d = _lib.RtodLayerDesc()
            d.src0 = d.src1 = -1
            kind = block["type"]
            if kind == "convolutional":
                conv = self.module_list[i][0]
                d.type = _lib.LAYER_CONV
                d.filters, d.size = conv.out_channels, conv.kernel_size[0]
                d.stride, d.pad = conv.stride[0], conv.padding[0]
                d.batch_normalize = int(len(self.module_list[i]) > 1 and
                                        isinstance(self.module_list[i][1], nn.BatchNorm2d))
                d.leaky = int(block["activation"] == "leaky")
            elif kind == "shortcut":
                d.type = _lib.LAYER_SHORTCUT
                d.src0, d.src1 = i - 1, i + int(block["from"])
            elif kind == "route":
                refs = [int(v) for v in block["layers"]]
                refs = [r if r > 0 else i + r for r in refs]
                if len(refs) > 2:
                    raise _lib.RtodError(-2, "route with more than two layers is not supported")
                d.type = _lib.LAYER_ROUTE
                d.src0 = refs[0]
                d.src1 = refs[1] if len(refs) > 1 else -1
            elif kind == "upsample":
                d.type = _lib.LAYER_UPSAMPLE
            elif kind == "maxpool":
                d.type = _lib.LAYER_MAXPOOL
                d.size, d.stride = int(block["size"]), int(block["stride"])
            elif kind == "yolo":
                anchors = self.module_list[i][0].anchors
                if len(anchors) > _lib.RTOD_MAX_ANCHORS:
                    raise _lib.RtodError(-2, "more than %d anchors per yolo layer" % _lib.RTOD_MAX_ANCHORS)
                d.type = _lib.LAYER_YOLO
                d.num_anchors, d.classes = len(anchors), int(block["classes"])
                for k, (w, h) in enumerate(anchors):
                    d.anchors[2 * k], d.anchors[2 * k + 1] = float(w), float(h)
            descs.append(d)
        return descs

    def _get_plan(self, key, device) -> _Plan:
        full_key = key + (device.index, self.plan_flags)
        plan = self._plans.get(full_key)
        if plan is None:
            if len(self._plans) >= 8:                       # bound arena memory: drop the least recently used shape
                victim = min(self._plans, key=lambda k: self._plans[k].last_used)
                self._plans.pop(victim).close()
            flags = self.plan_flags
            if torch.cuda.is_current_stream_capturing():   # bind-time autotuning synchronises: not inside a capture
                flags |= _lib.PLAN_NO_AUTOTUNE
            plan = _Plan(_lib.load(), self._layer_descs(), key, device, flags)
            self._plans[full_key] = plan
        self._plan_clock += 1
        plan.last_used = self._plan_clock
        return plan

    def _sync_weights(self, plan: _Plan, stream: int, sig=None):
        if sig is None:
            sig = self._weights_signature()
        if plan.weight_version == sig:
            return
        lib = plan.lib
        keep = []
        dev = plan.device

        def dptr(t):
            if t is None:
                return None
            t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        for i, block in enumerate(self.blocks[1:]):
            if block["type"] != "convolutional":
                continue
            module = self.module_list[i]
            conv = module[0]
            bn = module[1] if len(module) > 1 and isinstance(module[1], nn.BatchNorm2d) else None
            _lib.check(lib.rtod_plan_set_conv_weights(
                plan.handle, i, dptr(conv.weight), dptr(conv.bias),
                dptr(bn.weight) if bn is not None else None, dptr(bn.bias) if bn is not None else None,
                dptr(bn.running_mean) if bn is not None else None,
                dptr(bn.running_var) if bn is not None else None,
                float(bn.eps) if bn is not None else 0.0, stream))
        torch.cuda.current_stream(dev).synchronize()       # staging copies in `keep` may now die
        plan.weight_version = sig
        plan.weight_epoch = self._weights_epoch
        plan.graphs.clear()                                 # weights live in the same arena: the graphs stay
        # valid, but re-capture keeps the contract simple

    # ------------------------------------------------------------------ forward
    def forward(self, x, CUDA=None):
        """[B, 3, H, W] fp32 -> [B, N, 5+C] fp32 (src/darknet.py:199-253); ``[]`` without yolo layers.

        ``CUDA`` is tolerated for callers written against the north-star wording; the reference
        signature is ``forward(x)``.  Inference only: the result carries no ``grad_fn`` (the
        reference trainer, train.py:412-425, is out of scope -- a warning is issued once if a
        gradient is expected).
        """
        global _CUDA_SEEN
        lib = _lib.load()
        if not _CUDA_SEEN:                                   # (asked once: the query costs microseconds per call)
            if not torch.cuda.is_available():
                raise RuntimeError("Darknet.forward needs a CUDA device (B200, sm_100a); no CPU fallback")
            _CUDA_SEEN = True
        if self.training and not self._warned_train_bn:
            self._warned_train_bn = True
            warnings.warn("Darknet is in training mode; this implementation always evaluates "
                          "BatchNorm with running statistics (== reference .eval())", stacklevel=2)
        if torch.is_grad_enabled() and not self._warned_grad and \
                (x.requires_grad or (self.training and any(p.requires_grad for p in self.parameters()))):
            self._warned_grad = True
            warnings.warn("Darknet.forward is inference-only here: the prediction has no grad_fn, "
                          "loss.backward() will not reach the parameters", stacklevel=2)
        if x.dim() != 4:
            raise ValueError("expected a [B, C, H, W] tensor, got %s" % (tuple(x.shape),))
        if x.is_cuda:
            device = x.device
        else:
            params = next(self.parameters(), None)
            device = params.device if params is not None and params.is_cuda else \
                torch.device("cuda", torch.cuda.current_device())
        x = x.detach()
        # uint8 [B, 3, H, W] planes (util.prep_frames(..., as_uint8=True)) mean value / 255 (src/util.py:396): the stem
        # takes them as they are, the scale is folded into its weights
        if x.dtype == torch.uint8:
            x = x.to(device=device, non_blocking=True)
        elif not x.is_cuda or x.dtype != torch.float32:
            x = x.to(device=device, dtype=torch.float32, non_blocking=True)
        x = x.contiguous()
        inp_dim = int(self.net_info["height"])                       # :258 -- read at every call
        key = (x.size(0), x.size(1), x.size(2), x.size(3), inp_dim)
        if x.size(0) == 0:
            raise ValueError("empty batch")

        with torch.cuda.device(device):
            plan = self._get_plan(key, device)
            plan.raise_if_failed()                           # a time-out reported by an earlier forward
            stream = torch.cuda.current_stream(device)
            has_heads = plan.n_rows > 0
            # Weights are re-ingested when a parameter / buffer changed since the plan's last sync.  The exact test reads
            # the version counter of all ~370 tensors (~45 us of Python): on the graph-replay path -- the batch-1
            # latency path -- the replay is launched FIRST with the weights of the last sync and the counters are read
            # while the GPU works; in the rare case that something did change, the weights are re-ingested and the
            # forward runs again (same result as checking first, the check is off the critical path).  load_weights /
            # load_state_dict / .to() bump a cheap epoch, which is always checked before the launch.
            deferred = (has_heads and self.use_cuda_graph and plan.calls >= 1 and plan.weight_version is not None and
                        plan.weight_epoch == self._weights_epoch and not _EAGER_WEIGHT_CHECK and
                        not torch.cuda.is_current_stream_capturing())
            if not deferred:
                self._sync_weights(plan, stream.cuda_stream)
            if x.dtype == torch.uint8 and not plan.takes_u8:
                x = x.to(torch.float32) / 255.0                  # (bf16 storage / unusual stems: scale here)
            pred = self._run(plan, x, int(bool(self.TRAIN)), stream) if has_heads else None
            if deferred:
                sig = self._weights_signature()
                if sig != plan.weight_version:                   # a tensor was modified in place: redo with the new weights
                    self._sync_weights(plan, stream.cuda_stream, sig)
                    pred = self._run(plan, x, int(bool(self.TRAIN)), stream)
            if not has_heads:
                _lib.check(plan.forward_fn(x)(plan.handle, x.data_ptr(), None, 0, stream.cuda_stream))

        # side effects of the reference's yolo branch (:239-243, :260); blocks and modules are fixed after construction
        side = self._yolo_side
        if side is None:
            anchors, classes = [], None
            for i, block in enumerate(self.blocks[1:]):
                if block["type"] == "yolo":
                    anchors.extend(self.module_list[i][0].anchors)
                    classes = int(block["classes"])
            side = self._yolo_side = (tuple(anchors), classes)
        if side[1] is not None:
            self.anchors = list(side[0])
            self.num_classes = side[1]
        plan.calls += 1
        return pred if has_heads else []

    def _run(self, plan: _Plan, x: torch.Tensor, train: int, stream) -> torch.Tensor:
        lib = plan.lib
        shape = (plan.key[0], plan.n_rows, plan.n_attrs)
        capturing = torch.cuda.is_current_stream_capturing()
        if not self.use_cuda_graph or capturing or plan.calls < 1:
            pred = torch.empty(shape, dtype=torch.float32, device=plan.device)
            _lib.check(plan.forward_fn(x)(plan.handle, x.data_ptr(), pred.data_ptr(), train, stream.cuda_stream))
            return pred
        # The launch sequence is replayed as a CUDA graph.  A graph embeds its input pointer: inputs that keep
        # arriving in the same buffers (the pipeline's ring slots, a benchmark's resident batches) get a graph
        # of their own -- no staging copy; any other input is copied into the staging graph's buffer.
        ptr = x.data_ptr()
        entry = plan.graphs.get((ptr, train))
        if entry is not None and entry["dtype"] != x.dtype:
            entry = None
        # (reference call semantics, borrow_output False: a caller that keeps passing the SAME buffer -- seen on two
        # consecutive calls -- gets a graph on that pointer too, without the graph holding on to the caller's tensor:
        # no staging copy in front of the replay on the batch-1 latency path)
        if entry is None and len(plan.graphs) < 6 and (self.borrow_output or ptr == plan.last_ptr):
            entry = self._capture(plan, x, train, stream, key=(ptr, train))
            if entry is not None and not self.borrow_output:
                entry["x"] = None
        plan.last_ptr = ptr
        if entry is None:
            entry = plan.graphs.get((str(x.dtype), train))
            if entry is None:
                entry = self._capture(plan, torch.empty_like(x), train, stream, key=(str(x.dtype), train))
            if entry is not None:
                entry["x"].copy_(x, non_blocking=True)
        if entry is None:                                      # capture unsupported: stream launches
            return self._run(plan, x, train, stream)
        entry["graph"].replay()
        return entry["pred"] if self.borrow_output else entry["pred"].clone()

    def _capture(self, plan: _Plan, x: torch.Tensor, train: int, stream, key):
        shape = (plan.key[0], plan.n_rows, plan.n_attrs)
        entry = {"x": x, "dtype": x.dtype, "pred": torch.empty(shape, dtype=torch.float32, device=plan.device),
                 "graph": torch.cuda.CUDAGraph()}
        try:
            stream.synchronize()
            with torch.cuda.graph(entry["graph"]):
                _lib.check(plan.forward_fn(x)(plan.handle, x.data_ptr(), entry["pred"].data_ptr(), train,
                                              torch.cuda.current_stream(plan.device).cuda_stream))
        except Exception as exc:                              # capture unsupported: stay eager
            warnings.warn("CUDA graph capture failed (%s); using stream launches" % (exc,))
            self.use_cuda_graph = False
            return None
        plan.graphs[key] = entry
        return entry

    # ------------------------------------------------------------------ validation helpers
    def read_layer(self, index: int, key=None) -> torch.Tensor:
        """fp32 NCHW copy of one layer's output of the LAST forward (meaningful for every layer only
        when ``plan_flags`` contains ``PLAN_KEEP_ALL``).  Validation/debug aid, not reference API."""
        plan = next(reversed(self._plans.values())) if key is None else self._plans[key]
        lib = plan.lib
        c, h, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(lib.rtod_plan_layer_shape(plan.handle, index, ctypes.byref(c), ctypes.byref(h),
                                             ctypes.byref(w)))
        out = torch.empty(plan.key[0], c.value, h.value, w.value, dtype=torch.float32, device=plan.device)
        with torch.cuda.device(plan.device):
            _lib.check(lib.rtod_plan_read_layer(plan.handle, index, out.data_ptr(),
                                                torch.cuda.current_stream(plan.device).cuda_stream))
        return out

    def check_device(self):
        """Raise if a kernel of the last forward reported a device-side failure."""
        for plan in self._plans.values():
            with torch.cuda.device(plan.device):
                _lib.check(plan.lib.rtod_plan_check(plan.handle,
                                                    torch.cuda.current_stream(plan.device).cuda_stream))
