import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch, torch.nn.functional as F
from helpers import make_network
from realtimeobjectdetection_b200 import Darknet, _lib
from test_gpu_network import _fold, build_model
cfg, blocks, stream, state = make_network("yolov3-tiny", 8, "calibrated")
rng = np.random.RandomState(3)
u8 = torch.from_numpy(rng.randint(0, 256, (2, 3, 320, 320)).astype(np.uint8))
xf = u8.float().div(255.0)
model = build_model(cfg, state, 320, _lib.PLAN_KEEP_ALL, graph=False)
model(u8.cuda()); a = model.read_layer(0).cpu()
model(xf.cuda()); b = model.read_layer(0).cpu()
w, bb = _fold(state, 0, blocks[1])
want = F.leaky_relu(F.conv2d(xf.double(), w.double(), bb.double(), 1, 1), 0.1).float()
wq = want.half().float()
print("u8 != f32 path:", (a != b).float().mean().item(), " u8 != fp16(exact):", (a != wq).float().mean().item(), " f32 != fp16(exact):", (b != wq).float().mean().item())
d = (a != b)
print("mismatch by column band:", [d[..., :, lo:lo+64].float().mean().item() for lo in range(0, 320, 64)])
print("max abs diff", (a-b).abs().max().item(), "mean |want|", want.abs().mean().item())
