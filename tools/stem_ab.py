"""A/B timing of the stem kernel (layer 0 of YOLOv3-416, batch 64) and the whole forward, fp32 and uint8 frames:
RTOD_LIB=<other build> python tools/stem_ab.py"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import make_network
from realtimeobjectdetection_b200 import Darknet, _lib
lib = _lib.load()
cfg, blocks, stream, state = make_network("yolov3", 0, "calibrated")
model = Darknet(cfg, True); model.load_state_dict({**model.state_dict(), **state}); model.eval(); model.use_cuda_graph = False
model.plan_flags |= _lib.PLAN_NO_AUTOTUNE
B = 64
xf = torch.rand(B, 3, 416, 416, device="cuda")
xu = torch.randint(0, 256, (B, 3, 416, 416), dtype=torch.uint8, device="cuda")
model(xf)
plan = next(iter(model._plans.values()))
n = len(blocks) - 1
ms = (ctypes.c_float * (n + 1))(); kind = (ctypes.c_int * n)()
pred = torch.empty(B, plan.n_rows, plan.n_attrs, device="cuda")
acc = []
for r in range(11):
    _lib.check(lib.rtod_plan_forward_profile(plan.handle, xf.data_ptr(), pred.data_ptr(), 0, torch.cuda.current_stream().cuda_stream, ms, kind))
    if r: acc.append(ms[0])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {}
for name, x in (("fp32", xf), ("u8", xu)):
    for _ in range(3): model(x)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): model(x)
    e1.record(); torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1) / 10
print("stem (fp32 frames) median %.1f us min %.1f us | forward fp32 %.3f ms, u8 %.3f ms" % (sorted(acc)[5] * 1e3, min(acc) * 1e3, res["fp32"], res["u8"]))
