#!/usr/bin/env python
"""Single-convolution timing probe (bring-up tool, GPU box only):

    python tools/layer_probe.py L1 L3 L5 ...            # named YOLOv3-416 layers at batch 64
    RTOD_WSPLIT_AI=0 python tools/layer_probe.py L1     # env knobs of the library apply (see csrc/*.cu)

Each layer is planned alone through the C ABI (same kernels, same bind-time autotuner as the network), run
`reps` times back to back and timed with CUDA events; prints time, TFLOP/s, algorithmic GB/s and the autotuner's
choice (RTOD_TC_TUNE_DBG=1)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from realtimeobjectdetection_b200 import _lib   # noqa: E402

# name: (cin, cout, k, stride, H_in, residual)
LAYERS = {
    "L1": (32, 64, 3, 2, 416, 0), "L2": (64, 32, 1, 1, 208, 0), "L3": (32, 64, 3, 1, 208, 1),
    "L5": (64, 128, 3, 2, 208, 0), "L6": (128, 64, 1, 1, 104, 0), "L7": (64, 128, 3, 1, 104, 1),
    "L12": (128, 256, 3, 2, 104, 0), "L13": (256, 128, 1, 1, 52, 0), "L14": (128, 256, 3, 1, 52, 1),
    "L37": (256, 512, 3, 2, 52, 0), "L38": (512, 256, 1, 1, 26, 0), "L39": (256, 512, 3, 1, 26, 1),
    "L62": (512, 1024, 3, 2, 26, 0), "L63": (1024, 512, 1, 1, 13, 0), "L64": (512, 1024, 3, 1, 13, 1),
    "L99": (384, 128, 1, 1, 52, 0), "L105": (256, 255, 1, 1, 52, 0), "L81": (1024, 255, 1, 1, 13, 0),
    "L87": (768, 256, 1, 1, 26, 0), "L93": (512, 255, 1, 1, 26, 0), "L76": (512, 1024, 3, 1, 13, 0),
}


def aligned(t):
    return (t.data_ptr() + 255) // 256 * 256


def probe(name, batch, reps, flags):
    cin, cout, k, stride, H, res = LAYERS[name]
    lib = _lib.load()
    d1 = _lib.RtodLayerDesc()
    d1.type, d1.filters, d1.size, d1.stride, d1.pad, d1.batch_normalize, d1.leaky = _lib.LAYER_CONV, cout, k, stride, (k - 1) // 2, 0, 1
    d1.src0 = d1.src1 = -1
    descs = [d1]
    arr = (_lib.RtodLayerDesc * len(descs))(*descs)
    plan = ctypes.c_void_p()
    # the plan's input has `cin` channels: a layout-conversion launch (not timed) feeds the convolution
    _lib.check(lib.rtod_plan_create(arr, len(descs), batch, cin, H, H, 416, flags, ctypes.byref(plan)))
    ws = torch.empty(lib.rtod_plan_workspace_bytes(plan) + 256, dtype=torch.uint8, device="cuda")
    wa = torch.empty(lib.rtod_plan_weight_bytes(plan) + 256, dtype=torch.uint8, device="cuda")
    _lib.check(lib.rtod_plan_bind(plan, aligned(ws), lib.rtod_plan_workspace_bytes(plan), aligned(wa), lib.rtod_plan_weight_bytes(plan)))
    rng = np.random.RandomState(0)
    w = torch.from_numpy((rng.randn(cout, cin, k, k) / np.sqrt(cin * k * k)).astype(np.float32)).cuda()
    b = torch.zeros(cout, device="cuda")
    _lib.check(lib.rtod_plan_set_conv_weights(plan, 0, w.data_ptr(), b.data_ptr(), None, None, None, None, 0.0, None))
    x = torch.rand(batch, cin, H, H, device="cuda")
    n = len(descs)
    ms = (ctypes.c_float * (n + 1))()
    kind = (ctypes.c_int * n)()
    best, tot = 1e9, 0.0
    for r in range(reps + 2):
        _lib.check(lib.rtod_plan_forward_profile(plan, x.data_ptr(), None, 0, None, ms, kind))
        if r >= 2:
            best = min(best, ms[0])
            tot += ms[0]
    _lib.check(lib.rtod_plan_check(plan, None))
    Ho = (H + 2 * ((k - 1) // 2) - k) // stride + 1
    flops = 2.0 * batch * Ho * Ho * cout * cin * k * k
    bytes_ = 2.0 * batch * (H * H * cin + Ho * Ho * cout)
    avg = tot / reps
    cfg12 = (ctypes.c_int * 12)()
    lib.rtod_plan_conv_config(plan, 0, cfg12)
    print("      config: backend %d BN %d ctas %d resident %d sbufs %d splitK %d epi %d aprod %d pipelines %d stages %d wsplit %d grid %d" % tuple(cfg12))
    print("%-5s B=%d %dx%d %d->%d k%d s%d  split=%d backend=%d  avg %.1f us  best %.1f us  %.0f TF/s  %.0f GB/s (algorithmic, no shortcut)"
          % (name, batch, H, H, cin, cout, k, stride, lib.rtod_plan_conv_w_split(plan, 0), lib.rtod_plan_conv_backend(plan, 0),
             avg * 1e3, best * 1e3, flops / avg / 1e9, bytes_ / avg / 1e6), flush=True)
    lib.rtod_plan_destroy(plan)


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if a in LAYERS] or ["L1", "L3", "L5", "L7", "L13"]
    batch = int(os.environ.get("PROBE_BATCH", "64"))
    flags = int(os.environ.get("PROBE_FLAGS", "0"))
    for nm in names:
        probe(nm, batch, int(os.environ.get("PROBE_REPS", "10")), flags)
