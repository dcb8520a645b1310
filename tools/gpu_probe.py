#!/usr/bin/env python
"""Staged on-GPU bring-up checks (run under gpurun; each stage in its own process so that a
faulting kernel does not hide the results of the others).  Writes gpurun_out/probe_<stage>.log.

    python tools/gpu_probe.py <stage>      stage in: decode nms simt tc_tiny tc_v3 timing
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle                                                   # noqa: E402
from helpers import make_network, oracle_forward, rows_equal, synth_pred   # noqa: E402
from conftest import golden_names, load_golden                  # noqa: E402
from realtimeobjectdetection_b200 import (Darknet, _lib, bbox_iou, predict_transform,   # noqa: E402
                                          write_results)


def stage_decode():
    for name in golden_names("decode_"):
        g = load_golden(name)
        anchors = [tuple(int(v) for v in a) for a in g["anchors"]]
        out = predict_transform(torch.from_numpy(g["x"]).cuda(), int(g["inp_dim"]), anchors,
                                int(g["num_class"]), True, TRAIN=bool(g["train"])).cpu().numpy()
        err = np.abs(out - g["out"]) / (np.abs(g["out"]) + 1e-6)
        print("decode %-14s max rel %.3e  max abs %.3e  exact %.4f" %
              (name, err.max(), np.abs(out - g["out"]).max(), (out == g["out"]).mean()))


def stage_nms():
    for name in golden_names("nms_"):
        g = load_golden(name)
        out = write_results(torch.from_numpy(g["pred"]).cuda(), int(g["num_class"]), float(g["conf"]),
                            float(g["nms"]))
        want = 0 if int(g["is_zero"]) else torch.from_numpy(g["out"])
        ok = rows_equal(out, want)
        print("nms %-22s exact=%s got=%s want=%s" % (name, ok, None if isinstance(out, int) else tuple(out.shape),
                                                      None if isinstance(want, int) else tuple(want.shape)))
        if not ok and not isinstance(out, int) and not isinstance(want, int):
            o = out.cpu()
            n = min(len(o), len(want))
            bad = (o[:n] != want[:n]).any(1).nonzero().flatten()[:5]
            for b in bad:
                print("   row", int(b), o[b].tolist(), want[b].tolist())
    for dens, clustered in ((0.01, False), (0.10, False), (0.50, False), (0.10, True), (0.50, True)):
        pred = torch.from_numpy(synth_pred(5, 2, 10647, 80, dens, clustered))
        want = oracle.write_results(pred.clone(), 80, 0.5, 0.4)
        t = time.time()
        out = write_results(pred.cuda(), 80, 0.5, 0.4)
        torch.cuda.synchronize()
        print("nms full-size density %.2f clustered %d exact=%s kept=%d (%.1f ms incl. H2D)" %
              (dens, clustered, rows_equal(out, want), 0 if isinstance(out, int) else len(out),
               (time.time() - t) * 1e3))
    g = load_golden("iou_broadcast")
    out = bbox_iou(torch.from_numpy(g["box1"]).cuda(), torch.from_numpy(g["box2"]).cuda()).cpu().numpy()
    print("iou exact", bool((out == g["out"]).all()))


def layerwise(cfg_name, reso, batch, flags, mode="calibrated", seed=3, tag=""):
    cfg, blocks, stream, state = make_network(cfg_name, seed, mode)
    x = torch.from_numpy(np.random.RandomState(21).rand(batch, 3, reso, reso).astype(np.float32))
    port = oracle.DarknetPort(cfg, state)
    port.net_info["height"] = reso
    with torch.no_grad():
        want = port(x)
    model = Darknet(cfg, True)
    model.load_state_dict({**model.state_dict(), **state})
    model.net_info["height"] = reso
    model.plan_flags = flags | _lib.PLAN_KEEP_ALL
    model.use_cuda_graph = False
    model.eval()
    pred = model(x.cuda())
    torch.cuda.synchronize()
    model.check_device()
    worst = 0.0
    for i, blk in enumerate(blocks[1:]):
        if blk["type"] == "yolo":
            continue
        if blk["type"] == "convolutional" and i + 2 < len(blocks) and blocks[i + 2]["type"] == "shortcut":
            continue                                 # holds the fused shortcut result by design
        got = model.read_layer(i).cpu()
        ref = port.layer_outputs[i]
        scale = float(ref.abs().max()) + 1e-12
        err = float((got - ref).abs().max()) / scale
        worst = max(worst, err)
        flag = "" if err < 3e-2 else "   <<<<<<"
        print("%s layer %3d %-13s %-18s max|d|/max|ref| %.3e%s" %
              (tag, i, blk["type"], tuple(ref.shape[1:]), err, flag))
    p, w = pred.cpu(), want
    tol = (p - w).abs() <= (1e-3 + 1e-2 * w.abs())
    print("%s pred %s within rtol1e-2/atol1e-3: %.4f  max abs %.3e  worst layer err %.3e" %
          (tag, tuple(p.shape), float(tol.float().mean()), float((p - w).abs().max()), worst))
    d_ref = oracle.write_results(w.clone(), 80, 0.5, 0.4)
    d_got = write_results(pred, 80, 0.5, 0.4)
    print("%s detections ref %s got %s" % (tag, 0 if isinstance(d_ref, int) else len(d_ref),
                                          0 if isinstance(d_got, int) else len(d_got)))


def stage_simt():
    layerwise("yolov3-tiny", 160, 2, _lib.PLAN_CONV_SIMT, tag="simt-tiny")
    layerwise("yolov3", 128, 1, _lib.PLAN_CONV_SIMT, tag="simt-v3")


def stage_tc_tiny():
    layerwise("yolov3-tiny", 160, 2, 0, tag="tc-tiny")


def stage_tc_v3():
    layerwise("yolov3", 128, 2, 0, tag="tc-v3-128")
    layerwise("yolov3", 416, 1, 0, tag="tc-v3-416")
    layerwise("yolov3", 128, 1, 0, mode="default", seed=4, tag="tc-v3-default")


def stage_timing():
    cfg, blocks, stream, state = make_network("yolov3", 3, "calibrated")
    for batch, use_graph in ((1, False), (1, True), (16, True), (64, True)):
        model = Darknet(cfg, True)
        model.load_state_dict({**model.state_dict(), **state})
        model.eval()
        model.use_cuda_graph = use_graph
        x = torch.rand(batch, 3, 416, 416, device="cuda")
        for _ in range(3):
            pred = model(x)
        torch.cuda.synchronize()
        model.check_device()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20 if batch <= 16 else 5
        e0.record()
        for _ in range(iters):
            pred = model(x)
        e1.record()
        torch.cuda.synchronize()
        fwd = e0.elapsed_time(e1) / iters
        e0.record()
        for _ in range(iters):
            det = write_results(pred, 80, 0.5, 0.4)
        e1.record()
        torch.cuda.synchronize()
        nms = e0.elapsed_time(e1) / iters
        plan = next(iter(model._plans.values()))
        print("yolov3-416 B=%d graph=%d forward %.3f ms (%.1f frames/s, %.1f TFLOP/s) write_results %.3f ms dets %s"
              % (batch, use_graph, fwd, batch / fwd * 1e3, plan.conv_flops / fwd / 1e9, nms,
                 0 if isinstance(det, int) else len(det)))
    pred = torch.from_numpy(synth_pred(0, 256, 10647, 80, 0.01, False)).cuda()
    for _ in range(3):
        det = write_results(pred, 80, 0.5, 0.4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        det = write_results(pred, 80, 0.5, 0.4)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("write_results [256,10647,85] 1%%: %.3f ms -> %.1f GB/s, dets %d" % (ms, pred.numel() * 4 / ms / 1e6, len(det)))


def stage_layers():
    """per-layer device time (CUDA events around every launch) for B in argv[2:] (default 64 1)"""
    import ctypes
    lib = _lib.load()
    cfg, blocks, stream, state = make_network("yolov3", 3, "calibrated")
    batches = [int(v) for v in sys.argv[2:]] or [64, 1]
    for batch in batches:
        model = Darknet(cfg, True)
        model.load_state_dict({**model.state_dict(), **state})
        model.eval()
        model.use_cuda_graph = False
        x = torch.rand(batch, 3, 416, 416, device="cuda")
        model(x)
        plan = next(iter(model._plans.values()))
        n = len(blocks) - 1
        ms = (ctypes.c_float * (n + 1))()
        kind = (ctypes.c_int * n)()
        pred = torch.empty(batch, plan.n_rows, plan.n_attrs, device="cuda")
        acc = [0.0] * (n + 1)
        reps = 5
        for r in range(reps + 1):
            _lib.check(lib.rtod_plan_forward_profile(plan.handle, x.data_ptr(), pred.data_ptr(), 0,
                                                     torch.cuda.current_stream().cuda_stream, ms, kind))
            if r:
                for i in range(n + 1):
                    acc[i] += ms[i] / reps
        model.check_device()
        tot = sum(acc)
        print("B=%d total %.3f ms" % (batch, tot))
        for i in range(n):
            if not kind[i]:
                continue
            blk = blocks[i + 1]
            fl = lib.rtod_plan_layer_flops(plan.handle, i)
            c, h, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            lib.rtod_plan_layer_shape(plan.handle, i, ctypes.byref(c), ctypes.byref(h), ctypes.byref(w))
            desc = "%s k%s s%s" % (blk["type"][:4], blk.get("size", "-"), blk.get("stride", "-"))
            print("  L%3d kind %d %-14s -> %4dx%3dx%3d  %8.1f us  %7.1f TF/s  %5.1f%%" %
                  (i, kind[i], desc, c.value, h.value, w.value, acc[i] * 1e3, fl / max(acc[i], 1e-9) / 1e9,
                   100 * acc[i] / tot))
        print("  decode %8.1f us" % (acc[n] * 1e3))


def stage_clocks():
    """SM clock measured in-kernel (rtod_sm_clock_probe on a side stream) while the forward runs"""
    lib = _lib.load()
    cfg, blocks, stream, state = make_network("yolov3", 3, "calibrated")
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    model = Darknet(cfg, True)
    model.load_state_dict({**model.state_dict(), **state})
    model.eval()
    x = torch.rand(batch, 3, 416, 416, device="cuda")
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    out = torch.zeros(400, device="cuda")
    if not os.environ.get("RTOD_NO_SIDE"):
        _lib.check(lib.rtod_sm_clock_probe(out.data_ptr(), 400, 500, side.cuda_stream))   # 200 ms
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(25):
        model(x)
    e1.record()
    torch.cuda.synchronize()
    mhz = out.cpu().numpy()
    print("forward %.3f ms/step; SM MHz per 0.5 ms window: first %s ... median of busy part %.0f, min %.0f, max %.0f" %
          (e0.elapsed_time(e1) / 25, np.round(mhz[:6]).tolist(), float(np.median(mhz[20:300])), mhz[20:300].min(), mhz.max()))
    print("  every 20th:", np.round(mhz[::20]).tolist())


def stage_ncufwd():
    """one stream-launched forward + write_results between cudaProfilerStart/Stop (ncu --profile-from-start off):
    the plan is bound (and autotuned) and warmed up outside the profiled range"""
    lib = _lib.load()
    cfg, blocks, stream, state = make_network("yolov3", 0, "calibrated")      # bench.py's weights
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    model = Darknet(cfg, True)
    model.load_state_dict({**model.state_dict(), **state})
    model.eval()
    model.use_cuda_graph = False
    x = torch.rand(batch, 3, 416, 416, device="cuda")
    pred = model(x)
    write_results(pred, 80, 0.5, 0.4)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    pred = model(x)
    det = write_results(pred, 80, 0.5, 0.4)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled one forward + write_results, B=%d, %d detections" % (batch, 0 if isinstance(det, int) else len(det)))


def stage_stream():
    """BASELINE configs[4]: YOLOv3-tiny 320x320 streaming, batch 1, pinned frames, async H2D on a side stream:
    per-frame latency (frame handed over -> detections on the host) and sustained frames/s"""
    from realtimeobjectdetection_b200.pipeline import DetectionPipeline
    cfg, blocks, stream, state = make_network("yolov3-tiny", 3, "calibrated")
    model = Darknet(cfg, True)
    model.load_state_dict({**model.state_dict(), **state})
    model.net_info["height"] = 320
    model.eval()
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    frames = [(torch.randint(0, 256, (1, 3, 320, 320), dtype=torch.uint8).float() / 255.0).pin_memory() for _ in range(8)]
    for lag in (0, 1):
        pipe = DetectionPipeline(model, 80, 0.5, 0.4, collect_lag=lag)
        for _ in pipe.run(frames[i & 7] for i in range(50)):
            pass
        torch.cuda.synchronize()
        stamps = []

        def source():
            for i in range(n):
                stamps.append(time.perf_counter())
                yield frames[i & 7]
        lat = []
        t0 = time.perf_counter()
        for k, det in enumerate(pipe.run(source())):
            lat.append(time.perf_counter() - stamps[k])
        total = time.perf_counter() - t0
        lat = np.sort(np.array(lat)) * 1e3
        print("tiny-320 streaming B=1 collect_lag=%d: %.0f frames/s, per-frame latency p50 %.3f ms p99 %.3f ms" %
              (lag, n / total, lat[len(lat) // 2], lat[int(len(lat) * 0.99)]))


def stage_nmsbench():
    """BASELINE configs[3]: write_results on [256, 10647, 85] at 1/10/50 % density, C-ABI call only
    (CUDA events), plus the decode microbench on [256, 255, G, G] heads."""
    import ctypes
    lib = _lib.load()
    B, N, C = int(os.environ.get("NMSBENCH_B", "256")), 10647, 80
    nbytes = lib.rtod_write_results_workspace_bytes(B, N, C)
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    rows = torch.empty(B * N, 8, device="cuda")
    count = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for dens, clustered in ((0.01, False), (0.01, True), (0.10, True), (0.50, True), (0.50, False)):
        pred = torch.from_numpy(synth_pred(0, B, N, C, dens, clustered)).cuda()
        for _ in range(3):
            _lib.check(lib.rtod_write_results(pred.data_ptr(), B, N, C, 0.5, 0.4, rows.data_ptr(), B * N,
                                              count.data_ptr(), ws_ptr, nbytes, st))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            _lib.check(lib.rtod_write_results(pred.data_ptr(), B, N, C, 0.5, 0.4, rows.data_ptr(), B * N,
                                              count.data_ptr(), ws_ptr, nbytes, st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("write_results [%d,10647,85] density %.2f clustered %d: %.1f us -> %.0f GB/s (%.2f of 6552.6), %d detections"
              % (B, dens, clustered, ms * 1e3, pred.numel() * 4 / ms / 1e6, pred.numel() * 4 / ms / 1e6 / 6552.6, int(count.item())))
        del pred
    anchors = (ctypes.c_float * 6)(116, 90, 156, 198, 373, 326)
    for G in (13, 26, 52):
        head = torch.randn(B, 255, G, G, device="cuda")
        out = torch.empty(B, G * G * 3, 85, device="cuda")
        for _ in range(3):
            _lib.check(lib.rtod_yolo_decode(head.data_ptr(), B, G, 3, 80, 416, anchors, 0, out.data_ptr(), st))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            _lib.check(lib.rtod_yolo_decode(head.data_ptr(), B, G, 3, 80, 416, anchors, 0, out.data_ptr(), st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("yolo_decode [256,255,%d,%d]: %.1f us -> %.0f GB/s (%.2f of 6552.6)" %
              (G, G, ms * 1e3, head.numel() * 8 / ms / 1e6, head.numel() * 8 / ms / 1e6 / 6552.6))


if __name__ == "__main__":
    stage = sys.argv[1]
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    print("== stage", stage, torch.cuda.get_device_name(0))
    globals()["stage_" + stage]()
    torch.cuda.synchronize()
    print("== stage", stage, "done")
