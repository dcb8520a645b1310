"""Batch-1 latency split (YOLOv3-416, bench.py's weights): forward alone, write_results alone, both (CUDA events, p50 of 200)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import make_network
from realtimeobjectdetection_b200 import Darknet, write_results
cfg, blocks, stream, state = make_network("yolov3", 0, "calibrated")
model = Darknet(cfg, True); model.load_state_dict({**model.state_dict(), **state}); model.eval()
torch.manual_seed(0)
x = torch.rand(1, 3, 416, 416, device="cuda")
for _ in range(5):
    pred = model(x); det = write_results(pred, 80, 0.5, 0.4)
torch.cuda.synchronize()
def p50(fn, n=200):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.percentile(ts, 99))
import time
def wall(fn, n=200):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))
print("forward only          p50 %.3f ms p99 %.3f | wall %.3f" % (*p50(lambda: model(x)), wall(lambda: model(x))))
pred = model(x)
print("write_results only    p50 %.3f ms p99 %.3f | wall %.3f" % (*p50(lambda: write_results(pred, 80, 0.5, 0.4)), wall(lambda: write_results(pred, 80, 0.5, 0.4))))
print("forward+write_results p50 %.3f ms p99 %.3f | wall %.3f" % (*p50(lambda: write_results(model(x), 80, 0.5, 0.4)), wall(lambda: write_results(model(x), 80, 0.5, 0.4))))
