"""Do two independent frame batches in flight (two plans, two streams) fill the bubbles of the launch sequence --
tail waves, pipeline fill/drain, small kernels -- or does the power cap eat it?   python tools/two_stream_probe.py [B]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_network
from realtimeobjectdetection_b200 import Darknet, write_results_async

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg, blocks, stream, state = make_network("yolov3", 0, "calibrated")
models = []
for _ in range(2):
    m = Darknet(cfg, True); m.load_state_dict({**m.state_dict(), **state}); m.eval(); m.borrow_output = True
    models.append(m)
frames = [torch.rand(B, 3, 416, 416, device="cuda") for _ in range(4)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]

def run(n_steps, dual):
    pend = [None, None]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n_steps):
        k = (i & 1) if dual else 0
        with torch.cuda.stream(streams[k]):
            pred = models[k](frames[(i & 1) * 2 + k] if dual else frames[i & 1])
            h = write_results_async(pred, 80, 0.5, 0.4)
        if pend[k] is not None:
            pend[k].result()
        pend[k] = h
    for p in pend:
        if p is not None:
            p.result()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n_steps * 1e3

for dual in (False, True):
    run(6, dual)
for rep in range(2):
    print("single stream: %.3f ms/step   two streams: %.3f ms/step" % (run(40, False), run(40, True)))
