"""GPU parity tests of the whole network (run with -m gpu on the B200 box): ``Darknet.forward`` +
``write_results`` through the Python mirror of the reference API, against the CPU oracle and the committed
reference vectors -- including the exact workload ``bench.py`` times."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from conftest import golden_names, load_golden
from helpers import make_network, oracle_forward, rows_equal
from realtimeobjectdetection_b200 import Darknet, _lib, synth, write_results
from test_gpu_parity import BLOCK_TOL, DTYPES, frac_within, q16, q16_split

pytestmark = pytest.mark.gpu


def build_model(cfg, state, reso, flags=0, graph=True):
    model = Darknet(cfg, True)
    model.load_state_dict({**model.state_dict(), **state})
    model.net_info["height"] = reso
    model.plan_flags = flags
    model.use_cuda_graph = graph
    return model.eval()


def match_detections(got, want, iou_thr=0.9):
    """Detection-set agreement: how many of the oracle's detections have a counterpart of the same image and
    class with IoU >= iou_thr among ours (each of ours used once), and how many of ours are unmatched."""
    got = np.zeros((0, 8), np.float32) if isinstance(got, int) else got.cpu().numpy()
    want = np.zeros((0, 8), np.float32) if isinstance(want, int) else want.cpu().numpy()
    used = np.zeros(len(got), bool)
    matched = 0
    for w in want:
        cand = np.flatnonzero((got[:, 0] == w[0]) & (got[:, 7] == w[7]) & ~used)
        if cand.size == 0:
            continue
        ious = oracle.bbox_iou(torch.from_numpy(w[None, 1:5]), torch.from_numpy(got[cand, 1:5])).numpy()
        k = int(ious.argmax())
        if ious[k] >= iou_thr:
            used[cand[k]] = True
            matched += 1
    return matched, len(want), int((~used).sum())


def boundary_rows(pred, conf, margin):
    """rows whose objectness lies within `margin` of the threshold: the only ones whose presence may differ"""
    return int(((pred[..., 4] - conf).abs() <= margin).sum())


# ------------------------------------------------------------------ reference vectors, contract
@pytest.mark.parametrize("name", golden_names("forward_"))
@pytest.mark.parametrize("dtype,dflag", DTYPES)
def test_forward_against_reference_vectors(name, dtype, dflag):
    g = load_golden(name)
    cfg, blocks, stream, state = make_network(str(g["cfg"]), int(g["weight_seed"]), str(g["weight_mode"]))
    reso, batch = int(g["reso"]), int(g["batch"])
    x = torch.from_numpy(np.random.RandomState(int(g["input_seed"])).rand(batch, 3, reso, reso).astype(np.float32))
    model = build_model(cfg, state, reso, dflag)
    pred = model(x.cuda())
    model.check_device()
    want = torch.from_numpy(g["pred"])
    assert pred.shape == want.shape and pred.dtype == torch.float32 and pred.is_cuda
    assert model.anchors is not None and model.num_classes == 80
    frac = frac_within(pred.cpu(), want)
    if str(g["weight_mode"]) == "default":
        # the north-star contract: default-initialised network, rtol 1e-2 / atol 1e-3, every element
        assert frac == 1.0
        # (degenerate network: every objectness lies within 0.012 of the threshold, so the detection count is only
        # defined up to the rows that sit ON it)
        det = write_results(pred, 80, 0.5, 0.4)
        n_det = 0 if isinstance(det, int) else det.size(0)
        assert abs(n_det - g["det"].shape[0]) <= boundary_rows(want, 0.5, 1e-3)
    elif dtype == "fp16":
        # BN-calibrated random network in the shipped mode: the element-wise band holds almost everywhere and the
        # detections are the reference's
        assert frac >= 0.99, frac
        matched, n_want, extra = match_detections(write_results(pred, 80, 0.5, 0.4), torch.from_numpy(g["det"]))
        slack = boundary_rows(want, 0.5, 0.02)
        assert n_want - matched <= slack and extra <= slack, (matched, n_want, extra, slack)
    else:
        # bf16 storage (the north-star wording): 8 significant bits per stored value, ~75 layers deep -- the bulk
        # stays in tolerance, probabilities stay close on average ("hard parts" 1)
        assert frac > 0.30
        assert float((pred.cpu()[..., 4:] - want[..., 4:]).abs().mean()) < 0.02


def objectness_ties(pred, num_class, conf):
    """True if some (image, class) holds two candidates with bit-equal objectness."""
    for b in range(pred.size(0)):
        rows = pred[b][pred[b, :, 4] > conf]
        cls = rows[:, 5:5 + num_class].argmax(1)
        for c in cls.unique():
            obj = rows[cls == c][:, 4]
            if obj.unique().numel() != obj.numel():
                return True
    return False


@pytest.mark.parametrize("cfg_name,reso,batch", [("yolov3-tiny", 320, 2), ("yolov3", 416, 1), ("yolov3", 608, 1),
                                                 ("yolov3-tiny", 416, 3)])
@pytest.mark.parametrize("dtype,dflag", DTYPES)
def test_forward_default_init_contract(cfg_name, reso, batch, dtype, dflag):
    """BASELINE configs: default-initialised weights, eval-mode oracle, rtol 1e-2 / atol 1e-3."""
    cfg, blocks, stream, state = make_network(cfg_name, 31, "default")
    x = torch.from_numpy(np.random.RandomState(reso).rand(batch, 3, reso, reso).astype(np.float32))
    want = oracle_forward(cfg, state, x, reso)
    model = build_model(cfg, state, reso, dflag)
    pred = model(x.cuda())
    model.check_device()
    assert frac_within(pred.cpu(), want) == 1.0
    # detections: NMS is bit-exact on the same tensor -- unless two candidates of one (image, class)
    # have bit-equal objectness (common in this degenerate network): torch.sort(descending=True) is
    # not stable, so the reference's order among such rows is unspecified (ours: lower row first)
    got = write_results(pred, 80, 0.5, 0.4)
    host = pred.cpu()
    if not objectness_ties(host, 80, 0.5):
        assert rows_equal(got, oracle.write_results(host.clone(), 80, 0.5, 0.4))
    elif not isinstance(got, int):
        img, obj, cls = got[:, 0], got[:, 5], got[:, 7]
        ordered = (img[1:] > img[:-1]) | ((img[1:] == img[:-1]) & ((cls[1:] > cls[:-1]) |
                                                                  ((cls[1:] == cls[:-1]) & (obj[1:] <= obj[:-1]))))
        assert bool(ordered.all()) and bool((obj > 0.5).all())


# ------------------------------------------------------------------ the benchmarked workload
def test_bench_workload_parity_b64():
    """Exactly what bench.py times: YOLOv3-416, calibrated seed-0 weights ingested through load_weights, batch 64,
    autotuned plan (CTA-pair kernels on the 3x3 body), CUDA-graph replay.  Frames are independent, so four of
    the 64 are compared with the CPU oracle: prediction tensor inside rtol 1e-2 / atol 1e-3 on >= 99 % of
    the elements (fp16 storage; the remainder sits in the coarse 13x13 / 26x26 heads, whose logits carry the
    rounding of ~80 layers), and the detection sets agree up to rows whose objectness is at the threshold."""
    cfg, blocks, stream, state = make_network("yolov3", 0, "calibrated")
    path = os.path.join(os.environ.get("TMPDIR", "/tmp"), "rtod_test_bench_%d.weights" % os.getpid())
    synth.write_weights_file(path, stream)
    model = Darknet(cfg, True)
    model.load_weights(path)
    os.remove(path)
    model.net_info["height"] = 416
    model.eval()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234)                                     # bench.py: rank 0's first frame batch
    x = torch.rand(64, 3, 416, 416, device="cuda", generator=gen)
    model(x)                                                  # stream launches (binds + autotunes)
    pred = model(x)                                           # graph replay
    model.check_device()
    plan = next(iter(model._plans.values()))
    backends = [plan.lib.rtod_plan_conv_backend(plan.handle, i) for i in range(len(blocks) - 1)]
    assert backends.count(3) >= 10 and plan.graphs            # conv_pair_kernel on the 3x3 body, graph captured
    assert plan.is_f16
    frames = [0, 1, 31, 63]
    want = oracle_forward(cfg, state, x[frames].cpu(), 416)
    got = pred[frames].cpu()
    frac = frac_within(got, want)
    heads = [0, 507, 2535, 10647]
    per_head = [frac_within(got[:, heads[h]:heads[h + 1]], want[:, heads[h]:heads[h + 1]]) for h in range(3)]
    print("bench workload parity: %.4f of elements in tolerance (13x13 %.4f, 26x26 %.4f, 52x52 %.4f)" % (frac, *per_head))
    assert frac >= 0.99, (frac, per_head)
    assert per_head[2] >= 0.998 and per_head[0] >= 0.90
    # the whole batch agrees with itself: the same frames in a batch of 4 (other kernels, same arithmetic)
    small = model(x[frames].contiguous()).cpu()
    assert frac_within(got, small, 1e-3, 1e-4) >= 0.9999
    # detection sets
    det = write_results(pred, 80, 0.5, 0.4)
    sub = det[torch.isin(det[:, 0], torch.tensor(frames, device=det.device, dtype=det.dtype))].cpu()
    remap = {float(f): float(k) for k, f in enumerate(frames)}
    sub[:, 0] = torch.tensor([remap[float(v)] for v in sub[:, 0]])
    want_det = oracle.write_results(want.clone(), 80, 0.5, 0.4)
    matched, n_want, extra = match_detections(sub, want_det)
    slack = boundary_rows(want, 0.5, 0.01)
    print("bench workload detections: %d of %d oracle detections matched (class, IoU >= 0.9), %d extra, "
          "%d rows within 0.01 of the threshold" % (matched, n_want, extra, slack))
    assert n_want > 100 and n_want - matched <= slack and extra <= slack
    # ... and NMS itself is bit-exact on the device's own tensor
    assert rows_equal(write_results(got.cuda(), 80, 0.5, 0.4), oracle.write_results(got.clone(), 80, 0.5, 0.4))


# ------------------------------------------------------------------ every layer, tight
def _fold(state, i, blk):
    """BN-folded fp32 weight and bias of convolution i (what rtod_plan_set_conv_weights computes)"""
    w = state["module_list.%d.conv_%d.weight" % (i, i)]
    try:
        has_bn = bool(int(blk["batch_normalize"]))
    except (KeyError, ValueError):
        has_bn = False
    if not has_bn:
        return w, state["module_list.%d.conv_%d.bias" % (i, i)]
    pre = "module_list.%d.batch_norm_%d." % (i, i)
    scale = state[pre + "weight"] * (1.0 / torch.sqrt(state[pre + "running_var"] + 1e-5))
    return w * scale.view(-1, 1, 1, 1), (0 - state[pre + "running_mean"]) * scale + state[pre + "bias"]


@pytest.mark.parametrize("cfg_name,reso,batch", [("yolov3", 416, 1), ("yolov3-tiny", 416, 2), ("yolov3", 160, 3)])
@pytest.mark.parametrize("dtype,dflag", DTYPES)
def test_forward_every_layer_against_its_own_inputs(cfg_name, reso, batch, dtype, dflag):
    """Every launch of a non-degenerate network, checked in isolation: the oracle's fp32 op of layer i is
    evaluated on the DEVICE's stored outputs of the layers it reads (and the rounded weights the kernel
    multiplies with), so nothing accumulates and the comparison is at the storage type's rounding -- a wrong
    shortcut operand, concat offset, tap, stride or tile edge anywhere in the graph fails this."""
    cfg, blocks, stream, state = make_network(cfg_name, 3, "calibrated")
    x = torch.from_numpy(np.random.RandomState(21).rand(batch, 3, reso, reso).astype(np.float32))
    model = build_model(cfg, state, reso, dflag | _lib.PLAN_KEEP_ALL, graph=False)
    pred = model(x.cuda())
    model.check_device()
    plan = next(iter(model._plans.values()))
    n = len(blocks) - 1
    outs = {i: model.read_layer(i).cpu() for i in range(n) if blocks[i + 1]["type"] != "yolo"}
    rtol, atol = BLOCK_TOL[dtype]
    checked = {"convolutional": 0, "shortcut": 0, "route": 0, "upsample": 0, "maxpool": 0}
    for i, blk in enumerate(blocks[1:]):
        kind = blk["type"]
        if kind == "yolo":
            continue
        src = x if i == 0 else outs.get(i - 1)                   # (None behind a yolo layer: only routes follow)
        if kind == "convolutional":
            w, b = _fold(state, i, blk)
            k, stride = int(blk["size"]), int(blk["stride"])
            pad = (k - 1) // 2 if int(blk["pad"]) else 0
            if i > 0 or dtype == "bf16":                          # the fp16 stem multiplies in ~fp32 (two-term operands)
                w = q16_split(w) if plan.lib.rtod_plan_conv_w_split(plan.handle, i) else q16(w, dtype)
            y = F.conv2d(src, w, b, stride, pad)
            if blk["activation"] == "leaky":
                y = F.leaky_relu(y, 0.1)
            nxt = blocks[i + 2] if i + 2 <= n else None
            if nxt is not None and nxt["type"] == "shortcut":     # fused: the buffer holds conv + shortcut operand
                y = y + outs[i + 1 + int(nxt["from"])]
                assert torch.equal(outs[i], outs[i + 1])
            if nxt is not None and nxt["type"] == "yolo":         # fp32 logits
                assert frac_within(outs[i], y, 1e-4, 1e-4) == 1.0, i
            else:
                assert frac_within(outs[i], y, rtol, atol) == 1.0, (i, float((outs[i] - y).abs().max()))
        elif kind == "shortcut":
            if blocks[i]["type"] != "convolutional":              # (never in the reference cfgs: unfused add)
                assert frac_within(outs[i], outs[i - 1] + outs[i + int(blk["from"])], rtol, atol) == 1.0, i
        elif kind == "route":
            refs = oracle.DarknetPort._route_sources(i, blk)
            want = outs[refs[0]] if len(refs) == 1 else torch.cat((outs[refs[0]], outs[refs[1]]), 1)
            assert torch.equal(outs[i], want), i                  # zero-copy concat / alias: exact
        elif kind == "upsample":
            want = F.interpolate(src, scale_factor=2, mode="bilinear", align_corners=False)
            assert frac_within(outs[i], want, rtol, atol) == 1.0, i
        elif kind == "maxpool":
            assert torch.equal(outs[i], oracle.DarknetPort._maxpool(blk, src)), i
        checked[kind] += 1
    assert checked["convolutional"] == (75 if cfg_name == "yolov3" else 13)
    # the decode launch on the device's own logits
    heads = []
    for i, blk in enumerate(blocks[1:]):
        if blk["type"] == "yolo":
            mask = [int(v) for v in blk["mask"].split(",")]
            flat = [int(v) for v in blk["anchors"].split(",")]
            pairs = [(flat[k], flat[k + 1]) for k in range(0, len(flat), 2)]
            heads.append(oracle.predict_transform(outs[i - 1], reso, [pairs[m] for m in mask], 80, False))
    np.testing.assert_allclose(pred.cpu().numpy(), torch.cat(heads, 1).numpy(), rtol=2e-6, atol=1e-6)


def test_forward_tensor_core_and_cuda_core_paths_agree():
    cfg, blocks, stream, state = make_network("yolov3-tiny", 3, "calibrated")
    x = torch.from_numpy(np.random.RandomState(21).rand(2, 3, 160, 160).astype(np.float32)).cuda()
    a = build_model(cfg, state, 160, 0, graph=False)(x)
    b = build_model(cfg, state, 160, _lib.PLAN_CONV_SIMT, graph=False)(x)
    assert frac_within(a.cpu(), b.cpu(), 2e-3, 2e-3) >= 0.999


@pytest.mark.parametrize("cfg_name,reso,batch", [("yolov3", 416, 2), ("yolov3", 608, 1), ("yolov3-tiny", 416, 5)])
def test_forward_decode_ring_and_streaming_kernels_agree(monkeypatch, cfg_name, reso, batch):
    """heads decode inside the forward: the bulk-copy ring kernel (default for the 3 x 85 layout) and the streaming kernel
    produce the same prediction tensor bit for bit; ragged last tiles (13^2 = 169, 19^2 = 361 cells), 2 to 4 ring stages"""
    cfg, blocks, stream, state = make_network(cfg_name, 4, "calibrated")
    x = torch.from_numpy(np.random.RandomState(reso).rand(batch, 3, reso, reso).astype(np.float32)).cuda()
    monkeypatch.setenv("RTOD_DECODE_NO_RING", "1")
    want = build_model(cfg, state, reso, _lib.PLAN_NO_AUTOTUNE, graph=False)(x)
    monkeypatch.delenv("RTOD_DECODE_NO_RING")
    for stages in ("2", "3", "4"):
        monkeypatch.setenv("RTOD_DECODE_STAGES", stages)
        got = build_model(cfg, state, reso, _lib.PLAN_NO_AUTOTUNE, graph=False)(x)
        assert torch.equal(got, want), stages
    monkeypatch.delenv("RTOD_DECODE_STAGES")


@pytest.mark.parametrize("cfg_name,reso,batch", [("yolov3", 416, 2), ("yolov3-tiny", 160, 3), ("yolov3", 608, 1)])
def test_forward_upsample_block_and_per_output_kernels_agree(monkeypatch, cfg_name, reso, batch):
    """bilinear x2 (src/darknet.py:591-592): the kernel that builds a 2 x 2 output block per input pixel and the one-thread-
    per-output kernel evaluate the same expression: identical prediction tensors (5x5 .. 38x38 inputs, clamped borders)"""
    cfg, blocks, stream, state = make_network(cfg_name, 6, "calibrated")
    x = torch.from_numpy(np.random.RandomState(reso + batch).rand(batch, 3, reso, reso).astype(np.float32)).cuda()
    got = build_model(cfg, state, reso, _lib.PLAN_NO_AUTOTUNE, graph=False)(x)
    monkeypatch.setenv("RTOD_UPSAMPLE_PER_OUTPUT", "1")
    want = build_model(cfg, state, reso, _lib.PLAN_NO_AUTOTUNE, graph=False)(x)
    monkeypatch.delenv("RTOD_UPSAMPLE_PER_OUTPUT")
    assert torch.equal(got, want)


# ------------------------------------------------------------------ runtime behaviour
def test_forward_graph_replay_host_input_and_state_changes():
    cfg, blocks, stream, state = make_network("yolov3-tiny", 8, "calibrated")
    x = torch.from_numpy(np.random.RandomState(2).rand(2, 3, 224, 224).astype(np.float32))
    model = build_model(cfg, state, 224)
    p1 = model(x.cuda())                          # stream launches
    p2 = model(x.cuda())                          # CUDA graph capture + replay
    p3 = model(x)                                 # host tensor: staged H2D, result on the device
    assert torch.equal(p1, p2) and torch.equal(p1, p3) and p3.is_cuda
    assert p2.data_ptr() != p3.data_ptr()         # reference semantics: every call returns a fresh tensor
    # net_info["height"] is read at every call (src/darknet.py:258): centres scale with the stride,
    # widths do not ((exp * a/stride) * stride)
    model.net_info["height"] = 448
    p4 = model(x.cuda())
    assert torch.equal(p4[..., :2], p1[..., :2] * 2) and torch.equal(p4[..., 2:], p1[..., 2:])
    model.net_info["height"] = 224
    with model.train_mode():                      # TRAIN decode (src/util.py:211): sigmoids only
        pt = model(x.cuda())
    assert float(pt[..., :2].max()) <= 1.0 and torch.equal(pt[..., 4:], p1[..., 4:])
    with torch.no_grad():                         # in-place parameter updates are picked up
        model.module_list[0][0].weight.mul_(0.5)
    p5 = model(x.cuda())
    assert not torch.equal(p5, p1)
    model.check_device()


def test_forward_borrowed_output_replays_without_copies():
    """borrow_output (DetectionPipeline, bench): a graph per input buffer, the graph's own output tensor is returned"""
    cfg, blocks, stream, state = make_network("yolov3-tiny", 8, "calibrated")
    model = build_model(cfg, state, 224)
    xs = [torch.rand(2, 3, 224, 224, device="cuda") for _ in range(2)]
    want = [model(x).clone() for x in xs]
    model.borrow_output = True
    outs = [model(xs[k & 1]) for k in range(6)]
    plan = next(iter(model._plans.values()))
    assert {(xs[0].data_ptr(), 0), (xs[1].data_ptr(), 0)} <= set(plan.graphs)
    assert outs[2].data_ptr() == outs[0].data_ptr() and outs[1].data_ptr() != outs[0].data_ptr()
    assert torch.equal(outs[4], want[0]) and torch.equal(outs[5], want[1])
    xs[0].copy_(xs[1])                            # new frame in the same buffer: replay reads it
    assert torch.equal(model(xs[0]), want[1])


def test_device_failure_is_reported_without_a_sync():
    """a pipeline time-out must not be silent.  The kernels store their failure code in a pinned int the
    host polls at the next call; here the plan's sink is written directly (a real time-out takes 2 s)."""
    cfg, blocks, stream, state = make_network("yolov3-tiny", 8, "calibrated")
    model = build_model(cfg, state, 160)
    x = torch.rand(1, 3, 160, 160, device="cuda")
    p1 = model(x)
    plan = next(iter(model._plans.values()))
    plan.err_host[0] = 2
    with pytest.raises(_lib.RtodError, match="device-side failure"):
        model(x)
    assert torch.equal(model(x), p1)              # re-armed: flag, split-K counters and graphs reset


def test_inference_only_warning():
    cfg, blocks, stream, state = make_network("yolov3-tiny", 8, "calibrated")
    model = build_model(cfg, state, 160)
    model.train()
    x = torch.rand(1, 3, 160, 160, device="cuda")
    with pytest.warns(UserWarning) as rec:
        pred = model(x)
    assert any("inference-only" in str(w.message) for w in rec) and not pred.requires_grad


def test_load_weights_file_equals_state_dict(tmp_path):
    cfg, blocks, stream, state = make_network("yolov3-tiny", 12, "calibrated")
    path = str(tmp_path / "w.weights")
    synth.write_weights_file(path, stream)
    x = torch.rand(1, 3, 160, 160, device="cuda")
    a = build_model(cfg, state, 160)
    b = Darknet(cfg, True)
    b.load_weights(path)
    b.net_info["height"] = 160
    b.eval()
    assert torch.equal(a(x), b(x))


def test_streaming_pipeline_matches_direct_calls():
    from realtimeobjectdetection_b200.pipeline import DetectionPipeline
    cfg, blocks, stream, state = make_network("yolov3-tiny", 8, "calibrated")
    model = build_model(cfg, state, 320)
    batches = [torch.from_numpy(np.random.RandomState(k).rand(2, 3, 320, 320).astype(np.float32)) for k in range(5)]
    pipe = DetectionPipeline(model, 80, 0.5, 0.4)
    got = list(pipe.run(batches))
    assert len(got) == 5 and pipe.h2d_bytes == 5 * batches[0].numel() * 4
    assert model.borrow_output is False                          # restored after the run
    for b, det in zip(batches, got):
        want = write_results(model(b.cuda()), 80, 0.5, 0.4)
        assert rows_equal(det, want) and (isinstance(det, int) or not det.is_cuda)


@pytest.mark.parametrize("cfg_name,reso", [("yolov3-tiny", 320), ("yolov3", 416)])
def test_forward_uint8_planes(cfg_name, reso):
    """uint8 [B, 3, H, W] input = value / 255 (prep_image): the stem folds the scale into its two-term weights and feeds
    the pixels to the tensor cores exactly; against the oracle's stem on x / 255 and against the fp32-input forward"""
    cfg, blocks, stream, state = make_network(cfg_name, 8, "calibrated")
    rng = np.random.RandomState(3)
    u8 = torch.from_numpy(rng.randint(0, 256, (2, 3, reso, reso)).astype(np.uint8))
    xf = u8.float().div(255.0)                                   # host: true division like the reference
    model = build_model(cfg, state, reso, _lib.PLAN_KEEP_ALL, graph=False)
    pred_u8 = model(u8.cuda())
    model.check_device()
    plan = next(iter(model._plans.values()))
    assert plan.takes_u8
    stem_u8 = model.read_layer(0).cpu()
    w, b = _fold(state, 0, blocks[1])
    want = F.leaky_relu(F.conv2d(xf, w, b, 1, 1), 0.1)
    assert frac_within(stem_u8, want, *BLOCK_TOL["fp16"]) == 1.0
    pred_f = model(xf.cuda())
    stem_f = model.read_layer(0).cpu()
    assert frac_within(stem_f, want, *BLOCK_TOL["fp16"]) == 1.0
    assert float((stem_u8 != stem_f).float().mean()) < 1e-2       # both ~22-bit products, then one fp16 rounding
    # the two stems differ in ~0.2 % of their outputs by one fp16 ulp, which a 75-layer random network spreads out
    assert frac_within(pred_u8.cpu(), pred_f.cpu()) >= (0.999 if cfg_name == "yolov3-tiny" else 0.97)
    # graph replay and host tensors
    fast = build_model(cfg, state, reso)
    assert torch.equal(fast(u8.cuda()), fast(u8.cuda())) and torch.equal(fast(u8), fast(u8.cuda()))
    # bf16 storage has no uint8 stem: scaled on the way in
    bf = build_model(cfg, state, reso, _lib.PLAN_BF16)
    # (x / 255 is a reciprocal multiply on the device: one fp32 ulp, which 8-bit storage amplifies layer by layer)
    assert frac_within(bf(u8.cuda()).cpu(), bf(xf.cuda()).cpu()) >= 0.5


def test_streaming_pipeline_uint8_frames():
    """BASELINE configs[4]: uint8 frames in pinned host memory; letterbox + /255 run on the device (prep_frames),
    a quarter of the fp32 H2D bytes"""
    from realtimeobjectdetection_b200.pipeline import DetectionPipeline
    from realtimeobjectdetection_b200.util import prep_frames
    cfg, blocks, stream, state = make_network("yolov3-tiny", 8, "calibrated")
    model = build_model(cfg, state, 320)
    rng = np.random.RandomState(4)
    batches = [torch.from_numpy(rng.randint(0, 256, (2, 240, 320, 3)).astype(np.uint8)) for _ in range(4)]
    pipe = DetectionPipeline(model, 80, 0.5, 0.4)
    got = list(pipe.run(batches))
    assert len(got) == 4 and pipe.h2d_bytes == 4 * batches[0].numel()
    for b, det in zip(batches, got):
        want = write_results(model(prep_frames(b, 320, as_uint8=True)), 80, 0.5, 0.4)
        assert rows_equal(det, want)


def test_sharded_detection_over_nccl_single_rank():
    """The multi-GPU entry points on the NCCL backend (the driver's test tier has one GPU: world size 1; the world-2
    logic runs on gloo in test_host_cpu.py): detect_sharded, the synchronous gather and the fixed-capacity asynchronous
    gather return exactly write_results' rows; a pipeline with gather= yields the same."""
    import torch.distributed as dist
    from realtimeobjectdetection_b200.pipeline import DetectionPipeline
    from realtimeobjectdetection_b200.sharding import detect_sharded, gather_detections, gather_detections_async
    from realtimeobjectdetection_b200.util import write_results_async
    cfg, blocks, stream, state = make_network("yolov3-tiny", 8, "calibrated")
    model = build_model(cfg, state, 224)
    x = torch.from_numpy(np.random.RandomState(4).rand(3, 3, 224, 224).astype(np.float32))
    want = write_results(model(x.cuda()), 80, 0.3, 0.4)
    assert not isinstance(want, int)
    port = 29600 + os.getpid() % 300
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, world_size=1, rank=0,
                            device_id=torch.device("cuda", torch.cuda.current_device()))
    try:
        assert torch.equal(detect_sharded(model, x.cuda(), 80, 0.3, 0.4).cpu(), want.cpu())
        assert torch.equal(gather_detections(want, 0).cpu(), want.cpu())
        h = write_results_async(model(x.cuda()), 80, 0.3, 0.4)
        got = gather_detections_async(h.rows_device, h.count_device, 0, capacity=want.size(0) + 5).result()
        assert torch.equal(got, want.cpu())
        h = write_results_async(model(x.cuda()), 80, 0.3, 0.4)              # image column shifted by the rank's first frame
        got = gather_detections_async(h.rows_device, h.count_device, 6, capacity=want.size(0)).result()
        shifted = want.cpu().clone()
        shifted[:, 0] += 6.0
        assert torch.equal(got, shifted)
        h = write_results_async(model(x.cuda()), 80, 0.3, 0.4)              # more detections than the fixed capacity: loud
        with pytest.raises(RuntimeError):
            gather_detections_async(h.rows_device, h.count_device, 0, capacity=want.size(0) - 1).result()
        pipe = DetectionPipeline(model, 80, 0.3, 0.4, gather={"first_frame": 0, "capacity": want.size(0) + 5})
        outs = list(pipe.run([x.pin_memory(), x.pin_memory()]))
        assert len(outs) == 2 and all(torch.equal(o, want.cpu()) for o in outs)
    finally:
        dist.destroy_process_group()
