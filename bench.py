#!/usr/bin/env python
"""bench.py -- YOLOv3-416 frames/s (forward + decode + NMS), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the hot path over one batch of synthetic frames per rank:
``Darknet.forward`` (tcgen05 convolutions + fused decode) followed by ``write_results``
(threshold + per-class NMS), 80 classes, conf 0.5 / nms 0.4, YOLOv3 (cfg/yolov3.cfg) at
416x416.  Weights are synthetic (``synth.synth_stream(seed 0, "calibrated")``: random
BN-calibrated network whose objectness passes 0.5 for about 1 % of the rows), written as a
Darknet ``.weights`` file and ingested through ``load_weights`` like real ones.

value   whole-job frames/s with the frames already resident in HBM (CUDA events, barrier + sync
        on both sides, max over ranks).
e2e     the same metric through the public API with HOST buffers: pinned uint8 BGR frames (what a camera or
        cv2.imread delivers) are copied host->device inside the timed region (DetectionPipeline: side-stream
        copies overlap the previous batch), letterboxed / scaled on the device, and every step's detections
        are read back to the host.  e2e_fp32_frames: the same with fp32 [B,3,H,W] host tensors.
roofline  the dominant kernel (conv_tc_kernel, tensor bound): sum of 2*M*N*K over its launches
        divided by the sum of their device times, measured per layer with CUDA events by
        rtod_plan_forward_profile inside this run; peak = MEASURED_PEAKS.json.
cpu_baseline  the oracle port (PyTorch CPU ops, eval-mode BN) on the host cores, bounded sample.
--impl reference  times that CPU port as the reference arm (the reference is Python and cannot
        travel to the GPU box; oracle/ is its validated restatement), batch 1 per call like
        detect.py:27.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

METRIC = "YOLOv3-416 frames/s (fwd+decode+NMS)"
RESO, CLASSES, CONF, NMS = 416, 80, 0.5, 0.4
FRAME_GFLOP = 65.864                      # conv FLOPs per 416x416 frame


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cfg", default="yolov3")
    ap.add_argument("--reso", type=int, default=416, help="input size (BASELINE configs[2]: 608); default: the metric's 416")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


def synthetic_weights_file(cfg_name, seed=0):
    from realtimeobjectdetection_b200 import synth
    from realtimeobjectdetection_b200.cfg import builtin_cfg, parse_cfg
    cfg = builtin_cfg(cfg_name)
    blocks = parse_cfg(cfg)
    stream = synth.synth_stream(blocks, seed, "calibrated")
    path = os.path.join(tempfile.gettempdir(), "rtod_bench_%s_%d_%d.weights" % (cfg_name, seed, os.getpid()))
    synth.write_weights_file(path, stream)
    return cfg, blocks, stream, path


class ClockSampler:
    """SM clock, power and throttle reasons of one GPU DURING the timed region (B200_PROFILING.md).

    In-process NVML (pynvml) polled every 20 ms by a thread: `nvidia-smi -lms` -- the fallback when pynvml is missing --
    takes several hundred ms to start on an 8-GPU box and holds driver locks while it does, which showed up as two
    75-120 ms stalls of BOTH ranks inside a 0.1 s timed region in about one 2-GPU run out of five."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.handle, self.stop_flag, self.thread = None, None, False, None
        try:                                                  # NVML is initialised here, well before the timed region
            import pynvml
            pynvml.nvmlInit()
            handle = None
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.nvml, self.handle = pynvml, handle
        except Exception:
            self.nvml = None

    def _poll(self):
        nv, h = self.nvml, self.handle
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.rows.append((time.time(), mhz, self.max_mhz, [n for n, bit in names if mask & bit]))
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            time.sleep(1.0)                                   # let nvidia-smi finish its start-up outside the timed region
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            f = [v.strip() for v in line.split(",")]
            if len(f) < 8:
                continue
            try:
                reasons = [name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8])
                           if v.lower().startswith("active")]
                self.rows.append((time.time(), float(f[1]), float(f[2]), reasons))
            except ValueError:
                continue

    def stop(self, t0, t1):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05 if self.nvml is not None else 0.15)
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, mhz, max_mhz, why in list(self.rows):
            if not (t0 - 0.02 <= ts <= t1 + 0.02):
                continue
            sm.append(mhz)
            mx.append(max_mhz)
            reasons.update(why)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sample in timed region"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "how": "NVML polled in-process every 20 ms" if self.nvml is not None else "nvidia-smi -lms 100"}


# ------------------------------------------------------------------ reference / CPU arm
def cpu_port_frames_per_s(cfg, stream, blocks, frames, reps, batch_per_call, warmup):
    """oracle port (eval-mode BN) + oracle write_results on the host cores"""
    import oracle
    from realtimeobjectdetection_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    state = {k: torch.from_numpy(v) for k, v in synth.stream_to_state(blocks, stream).items()}
    port = oracle.DarknetPort(cfg, state)
    port.net_info["height"] = RESO
    x = torch.from_numpy(np.random.RandomState(1).rand(frames, 3, RESO, RESO).astype(np.float32))

    def one_pass():
        dets = 0
        with torch.no_grad():
            for lo in range(0, frames, batch_per_call):
                pred = port(x[lo:lo + batch_per_call])
                out = oracle.write_results(pred, CLASSES, CONF, NMS)
                dets += 0 if isinstance(out, int) else out.size(0)
        return dets

    for _ in range(warmup):
        one_pass()
    times = []
    for _ in range(reps):
        t = time.perf_counter()
        one_pass()
        times.append(time.perf_counter() - t)
    return frames / statistics.median(times), times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, blocks, stream, wpath = synthetic_weights_file(args.cfg)
    os.remove(wpath)
    frames_per_step = 2
    t0 = time.perf_counter()
    fps, times = cpu_port_frames_per_s(cfg, stream, blocks, frames_per_step, args.steps, 1, args.warmup)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": statistics.median(times) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "yolov3.cfg %dx%d, 80 classes, conf 0.5 / nms 0.4, CPU (CUDA=False), " % (RESO, RESO) +
                               "batch 1 per call as detect.py:27, %d frames per step" % frames_per_step,
                   "weights": "synthetic calibrated seed 0", "bn": "eval (running statistics)"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %d frames, oracle port of src/darknet.py + src/util.py (PyTorch %s CPU)"
                                   % (args.steps, frames_per_step, torch.__version__)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ B200 arm
def traffic_record(key):
    """ncu dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed summary of this tree
    (profiles/r2_conv_traffic.json, produced by tools/ncu_traffic.sh on YOLOv3-416 batch 64): `key` is a workload key
    ("yolov3-416-B64-fp16": the tcgen05 convolution launches of one forward) or a substring of a kernel-class name
    ("yolo_decode", "nms_": summed over the matching classes).  None if there is no record."""
    path = os.path.join(ROOT, "profiles", "r2_conv_traffic.json")
    if not os.path.exists(path):
        return None
    try:
        with open(path) as fh:
            rec = json.load(fh)
    except (OSError, ValueError):
        return None
    if key in rec:
        return rec[key]
    hit = [v for k, v in rec.get("classes", {}).items() if key in k]
    if not hit:
        return None
    return {"dram_bytes_per_launch": sum(v["dram_bytes_per_launch"] * v["launches"] for v in hit),
            "source": "ncu dram bytes summed over the %d launches of the matching kernels (tools/ncu_traffic.sh)"
                      % sum(v["launches"] for v in hit)}


def synth_microbench_tensor(dev, density, clustered, seed=7, B=256, N=10647, C=80):
    """BASELINE configs[3] tensor generated on the device: `density` of the rows of every image
    above objectness 0.5; clustered = the hot rows are jittered copies of 12 objects per image with a dominant class,
    so that NMS actually suppresses."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    mb = torch.rand(B, N, 5 + C, device=dev, generator=g)
    mb[..., 0:2] *= RESO
    mb[..., 2:4] = torch.exp(mb[..., 2:4] * 3 + 2)
    hot = torch.rand(B, N, device=dev, generator=g) < density
    mb[..., 4] = torch.where(hot, 0.5 + 0.4995 * mb[..., 4], 0.4995 * mb[..., 4])
    if clustered:
        n_obj = 12
        centres = torch.rand(B, n_obj, 2, device=dev, generator=g) * (RESO - 80) + 40
        sizes = torch.exp(torch.rand(B, n_obj, 2, device=dev, generator=g) * 2 + 3)
        classes = torch.randint(0, C, (B, n_obj), device=dev, generator=g)
        which = torch.randint(0, n_obj, (B, N), device=dev, generator=g)
        bi = torch.arange(B, device=dev).unsqueeze(1).expand(B, N)
        ctr = centres[bi, which] + torch.randn(B, N, 2, device=dev, generator=g) * 0.15 * sizes[bi, which]
        wh = sizes[bi, which] * torch.exp(torch.randn(B, N, 2, device=dev, generator=g) * 0.15)
        h3 = hot.unsqueeze(2)
        mb[..., 0:2] = torch.where(h3, ctr, mb[..., 0:2])
        mb[..., 2:4] = torch.where(h3, wh, mb[..., 2:4])
        mb[..., 5:] = torch.where(h3, mb[..., 5:] * 0.5, mb[..., 5:])
        dom = torch.rand(B, N, device=dev, generator=g) * 0.4 + 0.6
        cls_idx = classes[bi, which]
        cur = mb[..., 5:].gather(2, cls_idx.unsqueeze(2)).squeeze(2)
        mb[..., 5:].scatter_(2, cls_idx.unsqueeze(2), torch.where(hot, dom, cur).unsqueeze(2))
    return mb


def run_b200(args):
    import torch.distributed as dist
    from realtimeobjectdetection_b200 import Darknet, _lib, write_results, write_results_async
    from realtimeobjectdetection_b200.pipeline import DetectionPipeline
    from realtimeobjectdetection_b200.sharding import gather_detections_async

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    cfg, blocks, stream, wpath = synthetic_weights_file(args.cfg)
    model = Darknet(cfg, True)
    model.load_weights(wpath)                     # the reference's ingest path (src/darknet.py:316)
    os.remove(wpath)
    model.net_info["height"] = RESO
    model.eval()
    model.borrow_output = True                    # predictions are consumed in-stream: no copies around the CUDA graph
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    cap_rows = 256 * B                            # fixed capacity of the per-step detection gather (rows per rank)

    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    frames = [torch.rand(B, 3, RESO, RESO, device=dev, generator=gen) for _ in range(2)]   # 2 x 133 MB @ B=64

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def max_over_ranks(ms_local):
        ms = torch.tensor([ms_local], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def streaming_loop(net, batches, n_warm, n_steps, first_frame, capacity):
        """forward + write_results over resident frame batches, detections collected one step late (the GPU never
        idles on the host); at N > 1 every step's rows go to rank 0 through ONE fixed-capacity NCCL gather on a side
        stream.  Returns (ms of the n_steps timed steps on this rank, detections seen on rank 0, wall t0, t1)."""
        pending = [None]

        def collect():
            if pending[0] is None:
                return None
            det = pending[0].result()
            pending[0] = None
            return det

        def step(i):
            pred = net(batches[i & 1])
            handle = write_results_async(pred, CLASSES, CONF, NMS, device_count=world > 1)
            if world > 1:
                handle = gather_detections_async(handle.rows_device, handle.count_device, first_frame, capacity)
            det = collect()
            pending[0] = handle
            return det

        for i in range(n_warm):
            step(i)
        collect()
        barrier()
        n_seen = 0
        t0 = time.time()
        e0.record()
        trace = [] if os.environ.get("RTOD_BENCH_STEP_TIMES") else None
        for i in range(n_steps + 1):
            det = step(i) if i < n_steps else collect()      # n_steps enqueued, n_steps results collected
            if det is not None and not isinstance(det, int):
                n_seen += det.size(0)
            if trace is not None:
                trace.append(time.time())
        e1.record()
        if trace:
            sys.stderr.write("rank %d host ms per step: %s\n" % (rank, " ".join("%.1f" % ((b - a) * 1e3) for a, b in zip([t0] + trace[:-1], trace))))
        barrier()
        return e0.elapsed_time(e1), n_seen, t0, time.time()

    # ---- value: frames resident in HBM --------------------------------------------------------------------
    streaming_loop(model, frames, W, 0, rank * B, cap_rows)          # binds + autotunes + captures the graphs
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    ms_local, n_det, t0, t1 = streaming_loop(model, frames, 1, K, rank * B, cap_rows)
    ms_total = max_over_ranks(ms_local)
    clocks = sampler.stop(t0, t1)
    value = world * B * K / (ms_total / 1e3)
    model.check_device()
    plan = next(p for p in model._plans.values() if p.key[0] == B and p.key[2] == RESO)
    # nvidia-smi gets only a handful of samples inside a ~0.1 s timed region: a second, UNTIMED pass of the same
    # steps with a one-thread kernel counting SM cycles per 0.5 ms of wall time on a side stream
    # (rtod_sm_clock_probe; it slows the step down, which is why it is not part of the timed region)
    try:
        side = torch.cuda.Stream(dev)
        n_win = max(8, int(ms_total / 0.5))
        mhz = torch.zeros(n_win, device=dev)
        torch.cuda.synchronize()
        _lib.check(lib.rtod_sm_clock_probe(mhz.data_ptr(), n_win, 500, side.cuda_stream))
        streaming_loop(model, frames, 0, K + 2, rank * B, cap_rows)
        torch.cuda.synchronize()
        vals = sorted(float(v) for v in mhz.cpu() if v > 0)
        if vals:
            clocks["sm_mhz_in_kernel"] = {"median": vals[len(vals) // 2], "min": vals[0], "max": vals[-1],
                                          "windows": len(vals), "how": "clock64 / globaltimer per 0.5 ms, untimed pass"}
    except Exception as exc:                      # measurement aid only
        clocks["sm_mhz_in_kernel"] = {"error": str(exc)}

    # ---- e2e: HOST buffers through the public API ----------------------------------------------------------
    # uint8 BGR frames at network resolution in pinned host memory (what a camera / video decoder delivers, the
    # reference feeds cv2.imread output to prep_image): H2D of the raw bytes on a side stream, letterbox + BGR->RGB +
    # /255 on the device (rtod_prep_image), forward, write_results, detections read back to the host every step
    def e2e_run(host_batches, tag):
        gather = {"first_frame": rank * B, "capacity": cap_rows} if world > 1 else None
        pipe = DetectionPipeline(model, CLASSES, CONF, NMS, device=dev, gather=gather)
        for _ in pipe.run(host_batches[i & 1] for i in range(W)):
            pass
        pipe.h2d_bytes = pipe.d2h_bytes = 0
        barrier()
        e0.record()
        for _ in pipe.run(host_batches[i & 1] for i in range(K)):
            pass
        e1.record()
        barrier()
        ms_e = max_over_ranks(e0.elapsed_time(e1))
        return {"value": world * B * K / (ms_e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": pipe.h2d_bytes // K,
                "d2h_bytes_per_step": pipe.d2h_bytes // K, "ms_per_step": ms_e / K, "api": tag}

    rng = np.random.RandomState(99 + rank)
    host_u8 = [torch.from_numpy(rng.randint(0, 256, (B, RESO, RESO, 3), dtype=np.uint8)).pin_memory() for _ in range(2)]
    e2e = e2e_run(host_u8, "DetectionPipeline(model).run(pinned uint8 [B,%d,%d,3] BGR host frames) -> rtod_prep_image -> "
                           "Darknet.forward -> write_results -> host rows" % (RESO, RESO))
    host_f32 = [torch.rand(B, 3, RESO, RESO).pin_memory() for _ in range(2)]
    e2e_f32 = e2e_run(host_f32, "same pipeline fed pinned fp32 [B,3,%d,%d] host tensors (what prep_image returns)" % (RESO, RESO))
    # device pre-processing alone (rtod_prep_image): frames at network resolution (identity resize) and 640x480 camera
    # frames (cubic letterbox), uint8 planes out
    from realtimeobjectdetection_b200.util import prep_frames
    prep_times = {}
    for tag, (fh, fw) in (("%dx%d" % (RESO, RESO), (RESO, RESO)), ("640x480", (480, 640))):
        src = torch.randint(0, 256, (B, fh, fw, 3), dtype=torch.uint8, device=dev)
        dst = torch.empty(B, 3, RESO, RESO, dtype=torch.uint8, device=dev)
        for _ in range(3):
            prep_frames(src, RESO, out=dst, as_uint8=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            prep_frames(src, RESO, out=dst, as_uint8=True)
        e1.record()
        torch.cuda.synchronize()
        prep_times[tag] = e0.elapsed_time(e1) / 10
        del src, dst
    e2e["prep_ms_per_batch"] = prep_times
    del host_u8, host_f32

    # ---- roofline of the dominant kernel, measured live per layer --------------------------------
    peaks = measured_peaks()
    n_layers = len(blocks) - 1
    ms_arr = (ctypes.c_float * (n_layers + 1))()
    kind_arr = (ctypes.c_int * n_layers)()
    pred_buf = torch.empty(B, plan.n_rows, plan.n_attrs, device=dev)
    tc_ms, tc_flops, per_layer, decode_ms = 0.0, 0.0, [], 0.0
    reps = 3
    for r in range(reps + 1):
        _lib.check(lib.rtod_plan_forward_profile(plan.handle, frames[r & 1].data_ptr(), pred_buf.data_ptr(), 0,
                                                 torch.cuda.current_stream(dev).cuda_stream, ms_arr, kind_arr))
        if r == 0:
            continue                                   # warm-up pass of the profiled (non-graph) path
        decode_ms += ms_arr[n_layers] / reps
        for i in range(n_layers):
            if kind_arr[i] == 1:
                tc_ms += ms_arr[i]
                tc_flops += lib.rtod_plan_layer_flops(plan.handle, i)
    cfg12 = (ctypes.c_int * 12)()
    from realtimeobjectdetection_b200 import synth as _synth
    table = _synth.layer_table(blocks)
    c_, h_, w_ = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    for i in range(n_layers):
        if kind_arr[i]:
            row = [i, int(kind_arr[i]), round(float(ms_arr[i]), 4),
                   round(lib.rtod_plan_layer_flops(plan.handle, i) / max(ms_arr[i], 1e-6) / 1e9, 1)]
            if kind_arr[i] == 1 and lib.rtod_plan_conv_config(plan.handle, i, cfg12) == 0:
                info = {"pair": int(cfg12[0] == 3), "bn": cfg12[1], "ctas": cfg12[2], "resident": cfg12[3],
                        "pipelines": cfg12[8], "stages": cfg12[9], "two_term_weights": cfg12[10]}
                # the layer's own roofline: max(FLOP / sustained tensor peak, algorithmic bytes / HBM peak) against its
                # measured time (which carries the 2-5 us of its profiling events)
                t = table[i]
                lib.rtod_plan_layer_shape(plan.handle, i, ctypes.byref(c_), ctypes.byref(h_), ctypes.byref(w_))
                out_px = B * h_.value * w_.value
                in_px = out_px * t["stride"] * t["stride"]
                head = i + 1 < len(table) and table[i + 1]["type"] == "yolo"
                res = i + 1 < len(table) and table[i + 1]["type"] == "shortcut"
                nbytes = in_px * t["cin"] * 2 + out_px * t["cout"] * (4 if head else 2) * (2 if res else 1) + t["cout"] * t["cin"] * t["size"] ** 2 * 2
                bound_ms = max(lib.rtod_plan_layer_flops(plan.handle, i) / (peaks["tflops_sustained"] * 1e9), nbytes / (peaks["hbm_gbs"] * 1e6))
                info["bound_us"] = round(bound_ms * 1e3, 1)
                info["frac_of_bound"] = round(bound_ms / max(float(ms_arr[i]), 1e-9), 3)
                row.append(info)
            per_layer.append(row)
    n_tc = sum(1 for i in range(n_layers) if kind_arr[i] == 1)
    # the figure that is reported: CUDA events only where the stream switches between the tcgen05 convolution
    # launches and the other kernels (an event after each of the 74 launches adds a 2-5 us gap to every one)
    conv_ms_c, other_ms_c = ctypes.c_float(), ctypes.c_float()
    seg_conv, seg_other, seg_reps = 0.0, 0.0, 5
    for r in range(seg_reps + 1):
        _lib.check(lib.rtod_plan_forward_segments(plan.handle, frames[r & 1].data_ptr(), pred_buf.data_ptr(), 0,
                                                  torch.cuda.current_stream(dev).cuda_stream,
                                                  ctypes.byref(conv_ms_c), ctypes.byref(other_ms_c)))
        if r:
            seg_conv += conv_ms_c.value / seg_reps
            seg_other += other_ms_c.value / seg_reps
    per_launch_achieved = tc_flops / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
    achieved = (tc_flops / reps) / (seg_conv / 1e3) / 1e12 if seg_conv > 0 else 0.0
    fwd_ms = seg_conv + seg_other
    workload_key = "%s-%d-B%d-%s" % (args.cfg, RESO, B, "fp16" if plan.is_f16 else "bf16")
    same_workload = args.cfg == "yolov3" and RESO == 416 and B == 64
    traffic = traffic_record(workload_key)
    dec_traffic = traffic_record("yolo_decode") if same_workload else None
    nms_traffic = traffic_record("nms_") if same_workload else None
    roofline = {"bound": "tensor", "kernel": "conv_tc_kernel + conv_pair_kernel (the %d tcgen05 convolution launches of a forward)" % n_tc,
                "achieved": achieved, "achieved_with_per_launch_events": per_launch_achieved,
                "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops_sustained"],
                "frac_of_burst_peak": achieved / peaks["tflops_burst"],
                # dram bytes per launch (mean over the launches of one forward) from the committed ncu pass, else null
                "traffic": (traffic["dram_bytes_per_launch"] if traffic else None),
                "traffic_source": (traffic.get("source") if traffic else None),
                "algorithmic_bytes_per_launch": (traffic.get("algorithmic_bytes_per_launch") if traffic else None),
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s): kernel timed inside a long step"
                               % peaks["source"],
                "launches_per_step": n_tc, "flops_per_step": tc_flops / reps,
                "avg_launch_ms": seg_conv / max(n_tc, 1), "share_of_forward": seg_conv / fwd_ms}
    dec_bytes = B * plan.n_rows * plan.n_attrs * 8            # fp32 logits in, fp32 prediction out
    roofline_decode = {"bound": "hbm", "kernel": "yolo_decode_heads_ring_kernel (one launch, all heads)", "ms": decode_ms,
                       "achieved": dec_bytes / max(decode_ms, 1e-9) / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                       "frac": dec_bytes / max(decode_ms, 1e-9) / 1e6 / peaks["hbm_gbs"],
                       "traffic": (dec_traffic["dram_bytes_per_launch"] if dec_traffic else None),
                       "bytes_per_launch": dec_bytes}

    # NMS stage (HBM bound): device time of the rtod_write_results C-ABI call on this step's prediction tensor, and on
    # the BASELINE configs[3] microbench tensors
    def time_write_results(pred_t, iters=10):
        Bq, Nq, Lq = pred_t.shape
        nbytes = lib.rtod_write_results_workspace_bytes(Bq, Nq, Lq - 5)
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (ws.data_ptr() + 255) // 256 * 256
        rows_t = torch.empty(Bq * Nq, 8, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        for i in range(iters + 3):
            if i == 3:
                torch.cuda.synchronize()
                e0.record()
            _lib.check(lib.rtod_write_results(pred_t.data_ptr(), Bq, Nq, Lq - 5, CONF, NMS, rows_t.data_ptr(), Bq * Nq,
                                              cnt.data_ptr(), ws_ptr, nbytes, st))
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters, int(cnt.item())

    pred = model(frames[0]).clone()
    nms_ms, _ = time_write_results(pred)
    nms_bytes = pred.numel() * 4
    # algorithmic bytes = the whole tensor once; the sparse scan only pulls the objectness sectors and
    # the surviving rows from DRAM, so `achieved` may exceed the peak -- `traffic` holds the measured DRAM bytes
    roofline_nms = {"bound": "hbm",
                    "kernel": "rtod_write_results: nms_scan_sparse + nms_image (resident, boxes in smem, emit fused)"
                              if pred.size(0) <= 148 else "rtod_write_results: nms_scan_sparse + nms_image x2 + nms_emit",
                    "achieved": nms_bytes / nms_ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": nms_bytes / nms_ms / 1e6 / peaks["hbm_gbs"],
                    "traffic": (nms_traffic["dram_bytes_per_launch"] if nms_traffic else None), "ms": nms_ms,
                    "algorithmic_bytes": nms_bytes, "tensor": list(pred.shape)}
    microbench = None
    if rank == 0 and not args.no_latency:
        microbench = []
        for density, clustered in ((0.01, False), (0.01, True), (0.10, True), (0.50, True)):
            mb = synth_microbench_tensor(dev, density, clustered)
            mb_ms, mb_det = time_write_results(mb)
            microbench.append({"what": "write_results on [256,10647,85] fp32, %.0f %% of rows above conf 0.5, %s boxes"
                                       % (100 * density, "clustered" if clustered else "uniform"),
                               "ms": mb_ms, "GB/s": mb.numel() * 4 / mb_ms / 1e6,
                               "frac_of_hbm_peak": mb.numel() * 4 / mb_ms / 1e6 / peaks["hbm_gbs"], "detections": mb_det})
            del mb
        g3 = torch.Generator(device=dev)
        g3.manual_seed(11)
        heads = torch.randn(256, 255, 52, 52, device=dev, generator=g3)
        from realtimeobjectdetection_b200 import predict_transform
        for _ in range(3):
            predict_transform(heads, RESO, [(10, 13), (16, 30), (33, 23)], 80, True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            predict_transform(heads, RESO, [(10, 13), (16, 30), (33, 23)], 80, True)
        e1.record()
        torch.cuda.synchronize()
        d_ms = e0.elapsed_time(e1) / 10
        microbench.append({"what": "predict_transform (rtod_yolo_decode) on randn [256,255,52,52] fp32 NCHW", "ms": d_ms,
                           "GB/s": heads.numel() * 8 / d_ms / 1e6, "frac_of_hbm_peak": heads.numel() * 8 / d_ms / 1e6 / peaks["hbm_gbs"]})
        del heads

    # ---- p50 batch-1 latency (BASELINE configs[1]) ---------------------------------------------------
    latency = None
    if not args.no_latency and rank == 0:
        model.borrow_output = False
        x1 = torch.rand(1, 3, RESO, RESO, device=dev)
        for _ in range(20):
            write_results(model(x1), CLASSES, CONF, NMS)
        lat = []
        for _ in range(200):
            e0.record()
            write_results(model(x1), CLASSES, CONF, NMS)
            e1.record()
            e1.synchronize()
            lat.append(e0.elapsed_time(e1))
        latency = {"p50_ms": statistics.median(lat), "p99_ms": sorted(lat)[197], "batch": 1,
                   "what": "Darknet.forward + write_results (reference call semantics: fresh tensors, synchronous result), "
                           "frame resident in HBM, CUDA events"}
        model.borrow_output = True

    # ---- BASELINE configs[2]: YOLOv3 608x608, GLOBAL batch 64 sharded over the ranks (strong scaling) -------
    config2 = None
    if not args.no_latency and args.cfg == "yolov3" and RESO == 416 and 64 % world == 0:
        b2 = 64 // world
        model.net_info["height"] = 608
        g2 = torch.Generator(device=dev)
        g2.manual_seed(4321 + rank)
        big = [torch.rand(b2, 3, 608, 608, device=dev, generator=g2) for _ in range(2)]
        k2 = max(4, K // 2)
        streaming_loop(model, big, 3, 0, rank * b2, 256 * b2)
        ms2, _, _, _ = streaming_loop(model, big, 1, k2, rank * b2, 256 * b2)
        ms2 = max_over_ranks(ms2)
        config2 = {"workload": "yolov3.cfg 608x608 forward+decode+NMS, global batch 64 = %d frames per GPU x %d GPUs, "
                               "detections gathered to rank 0" % (b2, world),
                   "value": 64 * k2 / (ms2 / 1e3), "unit": "frames/s", "ms_per_step": ms2 / k2, "steps": k2, "scaling": "strong",
                   "forward_tflops": 140.692 * 64 / (ms2 / k2)}
        model.net_info["height"] = RESO
        del big

    # ---- BASELINE configs[4]: YOLOv3-tiny 320x320 streaming, batch 1, uint8 frames, pinned async H2D --------------
    config4 = None
    if not args.no_latency and rank == 0:
        from realtimeobjectdetection_b200.cfg import builtin_cfg
        tcfg, tblocks, tstream, tpath = synthetic_weights_file("yolov3-tiny")
        tiny = Darknet(builtin_cfg("yolov3-tiny"), True)
        tiny.load_weights(tpath)
        os.remove(tpath)
        tiny.net_info["height"] = 320
        tiny.eval()
        rng4 = np.random.RandomState(5)
        vid = [torch.from_numpy(rng4.randint(0, 256, (1, 320, 320, 3), dtype=np.uint8)).pin_memory() for _ in range(8)]
        pipe4 = DetectionPipeline(tiny, CLASSES, CONF, NMS, device=dev, collect_lag=0)
        for _ in pipe4.run(vid[i & 7] for i in range(50)):
            pass
        lat4 = []
        it4 = pipe4.run(vid[i & 7] for i in range(1000))
        torch.cuda.synchronize()
        t_all = time.perf_counter()
        while True:
            t = time.perf_counter()
            try:
                next(it4)
            except StopIteration:
                break
            lat4.append((time.perf_counter() - t) * 1e3)
        t_all = time.perf_counter() - t_all
        lat4.sort()
        pipe5 = DetectionPipeline(tiny, CLASSES, CONF, NMS, device=dev, collect_lag=1)
        torch.cuda.synchronize()
        t5 = time.perf_counter()
        for _ in pipe5.run(vid[i & 7] for i in range(1000)):
            pass
        t5 = time.perf_counter() - t5
        config4 = {"workload": "yolov3-tiny.cfg 320x320 streaming, batch 1, 1000 uint8 [1,320,320,3] frames in pinned host "
                               "memory, async H2D on a side stream, device letterbox, detections to the host",
                   "latency_ms_p50": lat4[len(lat4) // 2], "latency_ms_p99": lat4[int(len(lat4) * 0.99)],
                   "frames_per_s_latency_mode": len(lat4) / t_all, "frames_per_s_throughput_mode": 1000 / t5,
                   "how": "host wall clock per frame: H2D enqueue -> detections on the host (collect_lag=0); "
                          "throughput mode collects one frame late"}
        del tiny

    # ---- CPU baseline (rank 0, N=1 only; bounded sample) ------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, times = cpu_port_frames_per_s(cfg, stream, blocks, 32, 4, 4, 1)
        cpu = {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "32 frames (batches of 4) x 4 reps + 1 warm-up, oracle port, eval-mode BN, %.1f s"
                         % sum(times)}

    if rank == 0:
        # forward (convs, upsample, decode) + NMS: sparse scan + resident image pass (B <= 148), else scan / image x2 / emit
        launches_per_step = plan.launches + (2 if B <= 148 else 4)
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16" if plan.is_f16 else "bf16", "data": "synthetic",
            "config": {"workload": "yolov3.cfg %dx%d forward+decode+NMS, 80 classes, conf 0.5 / nms 0.4, " % (RESO, RESO) +
                                   "batch %d per GPU" % B,
                       "global_batch": world * B, "parallelism": "frames sharded, dp%d, no collective on the hot "
                                                                 "path; detections gathered to rank 0 by one "
                                                                 "fixed-capacity NCCL gather per step on a side stream" % world,
                       "weights": "synthetic calibrated seed 0 via load_weights", "bn": "folded (eval)",
                       "storage": "fp16 activations/weights, fp32 accumulate, two-term fp16 weights in the memory-bound layers"
                                  if plan.is_f16 else "bf16 activations/weights, fp32 accumulate",
                       "l2": "inputs alternate between two %.0f MB frame batches and every step streams >1 GB of "
                             "activations (> 126 MB L2)" % (B * 3 * RESO * RESO * 4 / 1e6),
                       "detections_per_step": n_det / max(K, 1), "cuda_graph": bool(model.use_cuda_graph)},
            "clocks": clocks, "e2e": e2e, "e2e_fp32_frames": e2e_f32, "gpu_launches": launches_per_step * K,
            "gpu_launches_per_step": launches_per_step,
            "roofline": roofline, "roofline_decode": roofline_decode, "roofline_nms": roofline_nms,
            "nms_microbench": microbench,
            "forward_tflops": FRAME_GFLOP * world * B / (ms_total / K) if args.cfg == "yolov3" and RESO == 416 else None,
            "latency_batch1": latency, "config2_608_batch64": config2, "config4_tiny320_streaming": config4,
            "cpu_baseline": cpu, "layers": per_layer,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    RESO = a.reso
    if RESO != 416:
        METRIC = "YOLOv3-%d frames/s (fwd+decode+NMS)" % RESO
    # libraries (NCCL's version banner) write to stdout: keep fd 1 for the one JSON line
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _print = print

    def print(*args, **kwargs):                                 # noqa: A001
        sys.stdout.flush()
        os.write(_real_stdout, (" ".join(str(x) for x in args) + "\n").encode())

    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
