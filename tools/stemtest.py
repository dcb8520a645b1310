import sys, torch, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import make_network
from realtimeobjectdetection_b200 import Darknet, _lib
import oracle
cfg, blocks, stream, state = make_network("yolov3-tiny", 6, "default")
for reso in (128, 160, 416):
    x = torch.from_numpy(np.random.RandomState(24).rand(1,3,reso,reso).astype(np.float32))
    m = Darknet(cfg, True); m.load_state_dict({**m.state_dict(), **state}); m.net_info["height"]=reso; m.eval(); m.use_cuda_graph=False
    m.plan_flags = _lib.PLAN_KEEP_ALL
    try:
        p = m(x.cuda()); torch.cuda.synchronize(); m.check_device()
        port = oracle.DarknetPort(cfg, state); port.net_info["height"]=reso
        with torch.no_grad(): port(x)
        got = m.read_layer(0).cpu(); ref = port.layer_outputs[0]
        print(reso, "layer0 err", float((got-ref).abs().max()/ref.abs().max()))
    except Exception as e:
        print(reso, "ERR", str(e)[:300]); break
