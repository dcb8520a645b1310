"""Forward time of YOLOv3-416 (bench.py's weights, autotuned plan, CUDA-graph replay) at batch PROBE_BATCH (64), for A/B runs
of library env knobs:  RTOD_TC_HINT_OUT=2 python tools/fwd_time.py [tag]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import make_network
from realtimeobjectdetection_b200 import Darknet, write_results
B = int(os.environ.get("PROBE_BATCH", "64"))
cfg, blocks, stream, state = make_network("yolov3", 0, "calibrated")
model = Darknet(cfg, True); model.load_state_dict({**model.state_dict(), **state}); model.eval()
torch.manual_seed(0)
x = torch.rand(B, 3, 416, 416, device="cuda")
for _ in range(4):
    pred = model(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(3):
    e0.record()
    for _ in range(20):
        pred = model(x)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
model.check_device()
print("%-28s B=%d forward %.3f ms  (checksum %.6f)" % (sys.argv[1] if len(sys.argv) > 1 else "", B, best, float(pred.double().sum())), flush=True)
