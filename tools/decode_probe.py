"""predict_transform (rtod_yolo_decode, NCHW drop-in) timing on the three YOLOv3-416 head shapes: GB/s of algorithmic
bytes (read + write) against the HBM peak"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realtimeobjectdetection_b200 import predict_transform
for B, G in ((256, 52), (256, 26), (256, 13), (64, 52), (64, 19)):
    x = torch.randn(B, 255, G, G, device="cuda")
    for _ in range(3):
        y = predict_transform(x, 416, [(10, 13), (16, 30), (33, 23)], 80, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y = predict_transform(x, 416, [(10, 13), (16, 30), (33, 23)], 80, True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gb = 2 * x.numel() * 4 / ms / 1e6
    print("predict_transform [%d,255,%d,%d]: %.3f ms  %.0f GB/s = %.2f of 6552.6" % (B, G, G, ms, gb, gb / 6552.6))
