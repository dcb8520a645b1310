"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI
(ctypes) and its Python mirror, against the CPU oracle on the same seeded inputs and against the
committed reference vectors.  Integer/index work (NMS) is bit-exact; floating point is compared
at the tolerance written next to each assert."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from conftest import golden_names, load_golden
from helpers import make_network, oracle_forward, rows_equal, synth_pred
from realtimeobjectdetection_b200 import (Darknet, _lib, bbox_iou, confidence_mask, predict_transform,
                                          write_results)

pytestmark = pytest.mark.gpu

# north-star tolerance for the prediction tensor (bf16 convolutions)
RTOL, ATOL = 1e-2, 1e-3
# 16-bit storage modes of a plan: fp16 is the default, bf16 (the north-star wording) a plan flag.  Per block, against
# an fp32 reference fed the SAME rounded operands, what remains is the fp32 accumulation order and one rounding of
# the stored result: 2^-11 relative for fp16, 2^-8 for bf16
DTYPES = [("fp16", 0), ("bf16", _lib.PLAN_BF16)]
BLOCK_TOL = {"fp16": (1.5e-3, 1e-3), "bf16": (1e-2, 1e-3)}


def q16(t, dtype):
    """round to the plan's storage type"""
    return t.half().float() if dtype == "fp16" else t.bfloat16().float()


def q16_split(t):
    """two-term fp16 weight (hi + lo), what rtod_plan_conv_w_split layers multiply with"""
    hi = t.half().float()
    return hi + (t - hi).half().float()
# fp32 decode: expf/division are within a few ulp of the CPU's
DEC_RTOL, DEC_ATOL = 2e-6, 1e-6


def frac_within(got, want, rtol=RTOL, atol=ATOL):
    return float(((got - want).abs() <= atol + rtol * want.abs()).float().mean())


# ------------------------------------------------------------------ decode -----------------
@pytest.mark.parametrize("name", golden_names("decode_"))
def test_decode_against_reference_vectors(name):
    g = load_golden(name)
    x = torch.from_numpy(g["x"]).cuda()
    keep = x.clone()
    anchors = [tuple(int(v) for v in a) for a in g["anchors"]]
    out = predict_transform(x, int(g["inp_dim"]), anchors, int(g["num_class"]), True, TRAIN=bool(g["train"]))
    assert torch.equal(x, keep)                                  # input not modified
    assert out.is_cuda and out.shape == g["out"].shape
    np.testing.assert_allclose(out.cpu().numpy(), g["out"], rtol=DEC_RTOL, atol=DEC_ATOL)


@pytest.mark.parametrize("grid,inp_dim", [(13, 416), (26, 416), (52, 416), (19, 608), (10, 320)])
def test_decode_against_oracle_full_heads(grid, inp_dim):
    torch.manual_seed(grid)
    x = torch.randn(4, 255, grid, grid) * 1.5
    anchors = [(116, 90), (156, 198), (373, 326)]
    want = oracle.predict_transform(x.clone(), inp_dim, anchors, 80, False)
    got = predict_transform(x.cuda(), inp_dim, anchors, 80, True)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=DEC_RTOL, atol=DEC_ATOL)
    host = predict_transform(x, inp_dim, anchors, 80, False)     # host tensor in -> host tensor out
    assert not host.is_cuda and torch.equal(host, got.cpu())


@pytest.mark.parametrize("B,grid,inp_dim,classes,train", [(70, 26, 416, 80, False), (3, 52, 416, 80, True), (5, 10, 320, 80, False),
                                                          (9, 20, 640, 20, False), (2, 76, 608, 80, False)])
def test_decode_tma_ring_and_one_tile_kernels_agree(monkeypatch, B, grid, inp_dim, classes, train):
    """G*G % 4 == 0: the TMA-pipelined persistent kernel (more tiles than resident CTAs x stages, ragged last tile of an
    image, 2 and 3 ring stages) against the oracle and, bit for bit, against the one-tile-per-CTA kernel"""
    torch.manual_seed(B + grid)
    x = torch.randn(B, 3 * (5 + classes), grid, grid) * 1.5
    anchors = [(10, 13), (16, 30), (33, 23)]
    want = oracle.predict_transform(x.clone(), inp_dim, anchors, classes, False, TRAIN=train)
    xd = x.cuda()
    outs = []
    for stages in ("2", "3"):
        monkeypatch.setenv("RTOD_DECODE_STAGES", stages)
        outs.append(predict_transform(xd, inp_dim, anchors, classes, True, TRAIN=train))
    monkeypatch.delenv("RTOD_DECODE_STAGES")
    monkeypatch.setenv("RTOD_DECODE_NO_TMA", "1")
    one_tile = predict_transform(xd, inp_dim, anchors, classes, True, TRAIN=train)
    monkeypatch.delenv("RTOD_DECODE_NO_TMA")
    np.testing.assert_allclose(outs[0].cpu().numpy(), want.numpy(), rtol=DEC_RTOL, atol=DEC_ATOL)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], one_tile)


def test_decode_rejects_bad_shapes():
    with pytest.raises(RuntimeError):
        predict_transform(torch.zeros(1, 254, 13, 13).cuda(), 416, [(1, 1)] * 3, 80, True)
    with pytest.raises(_lib.RtodError):                          # 100 // 13 = 7, 100 // 7 = 14 != 13
        predict_transform(torch.zeros(1, 255, 13, 13).cuda(), 100, [(1, 1)] * 3, 80, True)


# ------------------------------------------------------------------ NMS --------------------
@pytest.mark.parametrize("name", golden_names("nms_"))
def test_write_results_against_reference_vectors(name):
    g = load_golden(name)
    pred = torch.from_numpy(g["pred"]).cuda()
    keep = pred.clone()
    out = write_results(pred, int(g["num_class"]), float(g["conf"]), float(g["nms"]))
    assert torch.equal(pred, keep)
    want = 0 if int(g["is_zero"]) else torch.from_numpy(g["out"])
    assert rows_equal(out, want)                                 # bit-exact, including the int 0
    if not isinstance(out, int):
        assert out.is_cuda and out.dtype == torch.float32


@pytest.mark.parametrize("density,clustered", [(0.01, False), (0.10, False), (0.50, False),
                                               (0.01, True), (0.10, True), (0.50, True)])
def test_write_results_against_oracle_yolov3_rows(density, clustered):
    pred = torch.from_numpy(synth_pred(17, 2, 10647, 80, density, clustered))
    want = oracle.write_results(pred.clone(), 80, 0.5, 0.4)
    got = write_results(pred.cuda(), 80, 0.5, 0.4)
    assert rows_equal(got, want)


@pytest.mark.parametrize("B,N,C,conf,nms", [(1, 22743, 80, 0.5, 0.4), (5, 63, 1, 0.3, 0.5),
                                            (3, 2535, 20, 0.6, 0.45), (7, 33, 3, 0.5, 0.4),
                                            (1, 1, 2, 0.5, 0.4), (2, 1500, 80, 0.9, 0.1)])
def test_write_results_ragged_shapes(B, N, C, conf, nms):
    pred = torch.from_numpy(synth_pred(B * 1000 + N, B, N, C, 0.2, True, span=608.0))
    want = oracle.write_results(pred.clone(), C, conf, nms)
    got = write_results(pred.cuda(), C, conf, nms)
    assert rows_equal(got, want)


@pytest.mark.parametrize("density,clustered", [(0.01, False), (0.10, True), (0.50, True)])
def test_write_results_every_kernel_path_agrees(monkeypatch, density, clustered):
    """sparse scan + resident pass with fused emit (default, B <= 148)  ==  tensor-streaming scan + light / heavy pass +
    emit kernel (what B > 148 runs, forced here)  ==  the oracle"""
    pred = torch.from_numpy(synth_pred(21, 3, 10647, 80, density, clustered))
    want = oracle.write_results(pred.clone(), 80, 0.5, 0.4)
    outs = [write_results(pred.cuda(), 80, 0.5, 0.4)]
    monkeypatch.setenv("RTOD_NMS_NO_FUSED_EMIT", "1")
    outs.append(write_results(pred.cuda(), 80, 0.5, 0.4))
    monkeypatch.setenv("RTOD_NMS_STREAM", "1")
    outs.append(write_results(pred.cuda(), 80, 0.5, 0.4))
    monkeypatch.delenv("RTOD_NMS_NO_FUSED_EMIT")
    outs.append(write_results(pred.cuda(), 80, 0.5, 0.4))
    for o in outs:
        assert rows_equal(o, want)


def test_write_results_dense_single_class_global_sort_path():
    """> 16384 candidates in one image: the kernel sorts in global memory instead of shared."""
    pred = torch.from_numpy(synth_pred(3, 1, 22743, 2, 0.9, True, span=608.0))
    want = oracle.write_results(pred.clone(), 2, 0.5, 0.4)
    got = write_results(pred.cuda(), 2, 0.5, 0.4)
    assert rows_equal(got, want)


def test_write_results_properties_at_microbench_size():
    """[256, 10647, 85] (BASELINE configs[3]): too slow for the CPU oracle, so check what the domain
    guarantees -- ordering, threshold, per-class non-overlap -- plus equality with the oracle on the
    first images."""
    pred = torch.from_numpy(synth_pred(0, 256, 10647, 80, 0.01, True)).cuda()
    out = write_results(pred, 80, 0.5, 0.4)
    assert out.shape[1] == 8
    img, obj, cls = out[:, 0], out[:, 5], out[:, 7]
    key_ok = (img[1:] > img[:-1]) | ((img[1:] == img[:-1]) & ((cls[1:] > cls[:-1]) |
                                                              ((cls[1:] == cls[:-1]) & (obj[1:] <= obj[:-1]))))
    assert bool(key_ok.all())                                    # image asc, class asc, objectness desc
    assert bool((obj > 0.5).all())
    sub = pred[:4].cpu()
    want = oracle.write_results(sub.clone(), 80, 0.5, 0.4)
    assert rows_equal(out[out[:, 0] < 4], want)
    assert rows_equal(write_results(sub.cuda(), 80, 0.5, 0.4), want)
    rows = out[out[:, 0] < 32].cpu()                             # survivors of one class never overlap >= thr
    for b in rows[:, 0].unique():
        r = rows[rows[:, 0] == b]
        for c in r[:, 7].unique():
            boxes = r[r[:, 7] == c][:, 1:5]
            for i in range(len(boxes) - 1):
                assert bool((oracle.bbox_iou(boxes[i:i + 1], boxes[i + 1:]) < 0.4).all())


def test_bbox_iou_and_confidence_mask():
    g = load_golden("iou_broadcast")
    out = bbox_iou(torch.from_numpy(g["box1"]).cuda(), torch.from_numpy(g["box2"]).cuda())
    assert torch.equal(out.cpu(), torch.from_numpy(g["out"]))    # bit-exact
    a = torch.from_numpy(g["box2"][:100]).cuda()
    b = torch.from_numpy(g["box2"][100:200]).cuda()
    assert torch.equal(bbox_iou(a, b).cpu(), oracle.bbox_iou(a.cpu(), b.cpu()))
    pred = torch.from_numpy(synth_pred(1, 2, 300, 7, 0.3, False))
    assert torch.equal(confidence_mask(pred.cuda(), 0.5).cpu(), oracle.confidence_mask(pred, 0.5))


# ------------------------------------------------------------------ single convolution blocks
def _aligned(t):
    return (t.data_ptr() + 255) // 256 * 256


def run_block(lib, descs, x, weights, flags=0, backends=None, splits=None, configs=None):
    """Drive the C ABI directly: plan over `descs`, input x [B,C,H,W]; returns the fp32 NCHW output
    of every layer (and appends each layer's rtod_plan_conv_backend to `backends` / rtod_plan_conv_w_split to
    `splits` if given)."""
    B, C, H, W = x.shape
    arr = (_lib.RtodLayerDesc * len(descs))(*descs)
    plan = ctypes.c_void_p()
    _lib.check(lib.rtod_plan_create(arr, len(descs), B, C, H, W, 416, flags | _lib.PLAN_KEEP_ALL,
                                    ctypes.byref(plan)))
    ws = torch.empty(lib.rtod_plan_workspace_bytes(plan) + 256, dtype=torch.uint8, device="cuda")
    wa = torch.empty(lib.rtod_plan_weight_bytes(plan) + 256, dtype=torch.uint8, device="cuda")
    _lib.check(lib.rtod_plan_bind(plan, _aligned(ws), lib.rtod_plan_workspace_bytes(plan), _aligned(wa),
                                  lib.rtod_plan_weight_bytes(plan)))
    keep = []

    def ptr(t):
        if t is None:
            return None
        keep.append(t.cuda().contiguous())
        return keep[-1].data_ptr()

    for i, w in weights.items():
        _lib.check(lib.rtod_plan_set_conv_weights(plan, i, ptr(w["w"]), ptr(w.get("b")), ptr(w.get("gamma")),
                                                  ptr(w.get("beta")), ptr(w.get("mean")), ptr(w.get("var")),
                                                  1e-5, None))
    xd = x.cuda().contiguous()
    _lib.check(lib.rtod_plan_forward(plan, xd.data_ptr(), None, 0, None))
    _lib.check(lib.rtod_plan_check(plan, None))
    if backends is not None:
        backends.extend(lib.rtod_plan_conv_backend(plan, i) for i in range(len(descs)))
    if splits is not None:
        splits.extend(lib.rtod_plan_conv_w_split(plan, i) for i in range(len(descs)))
    if configs is not None:
        for i in range(len(descs)):
            cfg12 = (ctypes.c_int * 12)()
            if lib.rtod_plan_conv_config(plan, i, cfg12) == 0:
                configs.append(dict(zip(("backend", "bn", "ctas", "resident", "sbufs", "split_k", "epi_warps", "a_producers",
                                         "pipelines", "stages", "w_split", "grid"), list(cfg12))))
                configs[-1]["row"] = lib.rtod_plan_conv_row_mode(plan, i)
    outs = []
    for i in range(len(descs)):
        c, h, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(lib.rtod_plan_layer_shape(plan, i, ctypes.byref(c), ctypes.byref(h), ctypes.byref(w)))
        o = torch.empty(B, c.value, h.value, w.value, device="cuda")
        _lib.check(lib.rtod_plan_read_layer(plan, i, o.data_ptr(), None))
        outs.append(o.cpu())
    torch.cuda.synchronize()
    lib.rtod_plan_destroy(plan)
    return outs


def conv_desc(cout, k, stride, bn=True, leaky=True):
    d = _lib.RtodLayerDesc()
    d.type, d.filters, d.size, d.stride = _lib.LAYER_CONV, cout, k, stride
    d.pad, d.batch_normalize, d.leaky = (k - 1) // 2, int(bn), int(leaky)
    d.src0 = d.src1 = -1
    return d


def rand_conv(rng, cin, cout, k, bn=True):
    w = {"w": torch.from_numpy((rng.randn(cout, cin, k, k) / np.sqrt(cin * k * k)).astype(np.float32))}
    if bn:
        w.update(gamma=torch.from_numpy(rng.uniform(.5, 1.5, cout).astype(np.float32)),
                 beta=torch.from_numpy((rng.randn(cout) * .2).astype(np.float32)),
                 mean=torch.from_numpy((rng.randn(cout) * .3).astype(np.float32)),
                 var=torch.from_numpy(rng.uniform(.5, 2., cout).astype(np.float32)))
    else:
        w["b"] = torch.from_numpy((rng.randn(cout) * .2).astype(np.float32))
    return w


def ref_block(x, w, k, stride, leaky, emulate, split=False):
    """fp32 PyTorch reference of one block (src/darknet.py:488-501).  With emulate = "fp16" / "bf16" the input
    and the BN-folded weight are first rounded to that type -- what the tensor cores are fed (split: the
    weight as the two-term hi + lo sum) -- so that the remaining difference is accumulation order and the
    rounding of the output."""
    wt, bias = w["w"], w.get("b")
    if "gamma" in w:
        scale = w["gamma"] * (1.0 / torch.sqrt(w["var"] + 1e-5))
        wt = wt * scale.view(-1, 1, 1, 1)
        bias = (0 - w["mean"]) * scale + w["beta"]
    if emulate:
        x = q16(x, emulate)
        wt = q16_split(wt) if split else q16(wt, emulate)
    y = F.conv2d(x, wt, bias, stride, (k - 1) // 2)
    return F.leaky_relu(y, 0.1) if leaky else y


CONV_CASES = [  # cin, cout, k, stride, H, batch  (the shapes of SURVEY.md 2.2 at small spatial sizes)
    (32, 64, 3, 2, 32, 2), (64, 32, 1, 1, 32, 2), (64, 128, 3, 2, 24, 1), (128, 64, 1, 1, 24, 3),
    (128, 256, 3, 1, 26, 2), (256, 128, 1, 1, 26, 2), (256, 512, 3, 2, 26, 2), (512, 1024, 3, 1, 13, 2),
    (1024, 512, 1, 1, 13, 3), (768, 256, 1, 1, 19, 1), (384, 128, 1, 1, 10, 2), (16, 32, 3, 1, 40, 2),
    (32, 64, 3, 1, 21, 5), (512, 256, 1, 1, 5, 1)]


@pytest.mark.parametrize("cin,cout,k,stride,H,batch", CONV_CASES)
@pytest.mark.parametrize("flags", [0, _lib.PLAN_CONV_SIMT])
@pytest.mark.parametrize("dtype,dflag", DTYPES)
def test_conv_block_against_fp32_reference(lib, cin, cout, k, stride, H, batch, flags, dtype, dflag):
    rng = np.random.RandomState(cin * 7 + cout + k + stride)
    x = torch.from_numpy(rng.randn(batch, cin, H, H).astype(np.float32))
    w = rand_conv(rng, cin, cout, k)
    got = run_block(lib, [conv_desc(cout, k, stride)], x, {0: w}, flags | dflag)[0]
    ref_q = ref_block(x, w, k, stride, True, dtype)
    ref = ref_block(x, w, k, stride, True, None)
    assert got.shape == ref.shape
    # same rounded operands: only fp32 accumulation order + one rounding of the stored result remain
    assert frac_within(got, ref_q, *BLOCK_TOL[dtype]) == 1.0
    # against the exact fp32 block: operand rounding (2^-9 relative per operand for bf16, 2^-12 for fp16)
    assert float((got - ref).abs().max()) <= (2e-2 if dtype == "bf16" else 3e-3) * float(ref.abs().max())


# two-term fp16 weights (hi + lo, two MMAs per K step): chosen by shape for the large HBM-bound layers; forced on
# these small test shapes through the size knob.  Ring and resident weight tiles, 1x1 and 3x3, stride 2, shortcut.
SPLIT_CASES = [(32, 64, 3, 2, 32, 2), (64, 32, 1, 1, 32, 2), (64, 128, 3, 2, 24, 1), (128, 64, 1, 1, 24, 3),
               (256, 128, 1, 1, 26, 2), (384, 128, 1, 1, 10, 2), (32, 64, 3, 1, 21, 5), (16, 32, 3, 1, 40, 2)]


@pytest.mark.parametrize("cin,cout,k,stride,H,batch", SPLIT_CASES)
@pytest.mark.parametrize("flags", [0, _lib.PLAN_NO_AUTOTUNE, _lib.PLAN_CONV_SIMT])
def test_conv_block_two_term_weights(lib, monkeypatch, cin, cout, k, stride, H, batch, flags):
    monkeypatch.setenv("RTOD_WSPLIT_ELEMS", "0")
    rng = np.random.RandomState(cin * 5 + cout + k + stride)
    x = torch.from_numpy(rng.randn(batch, cin, H, H).astype(np.float32))
    w = rand_conv(rng, cin, cout, k)
    splits = []
    got = run_block(lib, [conv_desc(cout, k, stride)], x, {0: w}, flags, splits=splits)[0]
    assert splits == [1]
    ref_q = ref_block(x, w, k, stride, True, "fp16", split=True)
    assert frac_within(got, ref_q, *BLOCK_TOL["fp16"]) == 1.0
    # the weight rounding is gone: against fp32 weights on the same rounded input only the output rounding remains
    ref_w = ref_block(q16(x, "fp16"), w, k, stride, True, None)
    assert frac_within(got, ref_w, *BLOCK_TOL["fp16"]) == 1.0
    # ... and the plan flag switches it off
    splits = []
    run_block(lib, [conv_desc(cout, k, stride)], x, {0: w}, flags | _lib.PLAN_NO_WSPLIT, splits=splits)
    assert splits == [0]


DUAL_CASES = [(32, 64, 3, 1, 64, 4, 64), (32, 64, 3, 2, 104, 2, 64), (64, 32, 1, 1, 48, 3, 32), (16, 32, 3, 1, 40, 2, 32),
              (128, 64, 1, 1, 40, 3, 64), (32, 32, 3, 1, 30, 5, 32)]


@pytest.mark.parametrize("cin,cout,k,stride,H,batch,bn", DUAL_CASES)
@pytest.mark.parametrize("split", [0, 1])
@pytest.mark.parametrize("dtype,dflag", DTYPES)
def test_conv_block_dual_pipeline(lib, monkeypatch, cin, cout, k, stride, H, batch, bn, split, dtype, dflag):
    """two warp sets (producers, MMA issuer, epilogue, ring, accumulators) in one CTA sharing the resident weights,
    walking alternate tiles: same result, bit for bit, as the single-pipeline configuration; odd tile counts, shortcut"""
    if split and dtype == "bf16":
        pytest.skip("two-term weights are an fp16 feature")
    monkeypatch.setenv("RTOD_WSPLIT_ELEMS", "0")
    monkeypatch.setenv("RTOD_WSPLIT_AI", "10000" if split else "0")
    rng = np.random.RandomState(cin * 11 + cout + k + stride)
    x = torch.from_numpy(rng.randn(batch, cin, H, H).astype(np.float32))
    w = rand_conv(rng, cin, cout, k)
    outs = {}
    for sbufs in (2, 1):
        monkeypatch.setenv("RTOD_TC_FORCE", "0,%d,1,1,%d,0,2" % (bn, sbufs))
        cfgs = []
        outs[sbufs] = run_block(lib, [conv_desc(cout, k, stride)], x, {0: w}, dflag, configs=cfgs)[0]
        assert cfgs[0]["pipelines"] == 2 and cfgs[0]["resident"] == 1 and cfgs[0]["w_split"] == split, cfgs
    monkeypatch.setenv("RTOD_TC_FORCE", "0,%d,1,0,2,0,1" % bn)
    cfgs = []
    single = run_block(lib, [conv_desc(cout, k, stride)], x, {0: w}, dflag, configs=cfgs)[0]
    assert cfgs[0]["pipelines"] == 1
    monkeypatch.delenv("RTOD_TC_FORCE")
    ref_q = ref_block(x, w, k, stride, True, dtype, split=bool(split))
    assert frac_within(single, ref_q, *BLOCK_TOL[dtype]) == 1.0
    assert torch.equal(outs[2], single) and torch.equal(outs[1], single)


def test_residual_block_dual_pipeline(lib, monkeypatch):
    monkeypatch.setenv("RTOD_WSPLIT_ELEMS", "0")
    monkeypatch.setenv("RTOD_WSPLIT_AI", "10000")
    rng = np.random.RandomState(31)
    x = torch.from_numpy(rng.randn(3, 32, 96, 96).astype(np.float32))     # >= 120 M tiles: the heuristic keeps one N tile
    w0, w1, w2 = rand_conv(rng, 32, 64, 3), rand_conv(rng, 64, 32, 1), rand_conv(rng, 32, 64, 3)
    sc = _lib.RtodLayerDesc()
    sc.type, sc.src0, sc.src1 = _lib.LAYER_SHORTCUT, 2, 0
    monkeypatch.setenv("RTOD_TC_FORCE", "0,0,1,1,2,0,2")      # bn 0 = by shape
    cfgs = []
    outs = run_block(lib, [conv_desc(64, 3, 1), conv_desc(32, 1, 1), conv_desc(64, 3, 1), sc], x, {0: w0, 1: w1, 2: w2}, configs=cfgs)
    monkeypatch.delenv("RTOD_TC_FORCE")
    assert [c["pipelines"] for c in cfgs] == [2, 2, 2] and [c["w_split"] for c in cfgs] == [1, 1, 1]
    tol = BLOCK_TOL["fp16"]
    assert frac_within(outs[0], ref_block(x, w0, 3, 1, True, "fp16", split=True), *tol) == 1.0
    assert frac_within(outs[1], ref_block(outs[0], w1, 1, 1, True, "fp16", split=True), *tol) == 1.0
    assert frac_within(outs[3], ref_block(outs[1], w2, 3, 1, True, "fp16", split=True) + outs[0], *tol) == 1.0


# row mode (3x3 / stride 1, few input channels): a tile is a 128-pixel segment of one image row, one tiled TMA load per
# (filter row, 32-channel slice), the three kx taps as row-shifted views of the staged slab.  Images wider / narrower than
# a segment, ragged last segments, one and two channel slices, two-term weights, both pipelines counts.
ROW_CASES = [(32, 64, 136, 2, 64), (32, 64, 40, 3, 64), (64, 128, 104, 1, 128), (64, 64, 129, 1, 64), (32, 32, 260, 1, 32),
             (128, 32, 50, 2, 32)]


@pytest.mark.parametrize("cin,cout,H,batch,bn", ROW_CASES)
@pytest.mark.parametrize("split", [0, 1])
@pytest.mark.parametrize("dtype,dflag", DTYPES)
def test_conv_block_row_mode(lib, monkeypatch, cin, cout, H, batch, bn, split, dtype, dflag):
    if split and dtype == "bf16":
        pytest.skip("two-term weights are an fp16 feature")
    if (cout << split) * cin * 9 * 2 > 150 * 1024:
        pytest.skip("row mode keeps the whole weight matrix in shared memory")
    monkeypatch.setenv("RTOD_WSPLIT_ELEMS", "0")
    monkeypatch.setenv("RTOD_WSPLIT_AI", "10000" if split else "0")
    rng = np.random.RandomState(cin * 13 + cout + H)
    x = torch.from_numpy(rng.randn(batch, cin, H, H).astype(np.float32))
    w = rand_conv(rng, cin, cout, 3)
    outs = []
    for force in ("0,%d,1,1,2,0,1,1" % bn, "0,%d,1,1,1,0,1,1" % bn, "0,%d,1,1,2,0,2,1" % bn):
        monkeypatch.setenv("RTOD_TC_FORCE", force)
        cfgs = []
        got = run_block(lib, [conv_desc(cout, 3, 1)], x, {0: w}, dflag, configs=cfgs)[0]
        assert cfgs[0]["w_split"] == split
        if cfgs[0]["row"] == 1:                          # (a configuration that does not fit falls back to the default)
            assert cfgs[0]["resident"] == 1
            outs.append(got)
    assert len(outs) >= 1
    monkeypatch.setenv("RTOD_TC_FORCE", "0,%d,1,0,2,0,1,0" % bn)
    cfgs = []
    gather = run_block(lib, [conv_desc(cout, 3, 1)], x, {0: w}, dflag, configs=cfgs)[0]
    assert cfgs[0]["row"] == 0
    monkeypatch.delenv("RTOD_TC_FORCE")
    ref_q = ref_block(x, w, 3, 1, True, dtype, split=bool(split))
    assert frac_within(gather, ref_q, *BLOCK_TOL[dtype]) == 1.0
    for o in outs:                                       # same K order as the im2col gather: bit-identical
        assert torch.equal(o, gather)


def test_residual_block_row_mode(lib, monkeypatch):
    rng = np.random.RandomState(37)
    x = torch.from_numpy(rng.randn(2, 64, 104, 104).astype(np.float32))
    w0, w1, w2 = rand_conv(rng, 64, 128, 3), rand_conv(rng, 128, 64, 1), rand_conv(rng, 64, 128, 3)
    sc = _lib.RtodLayerDesc()
    sc.type, sc.src0, sc.src1 = _lib.LAYER_SHORTCUT, 2, 0
    descs = [conv_desc(128, 3, 1), conv_desc(64, 1, 1), conv_desc(128, 3, 1), sc]
    monkeypatch.setenv("RTOD_TC_ROW", "1")
    cfgs = []
    outs = run_block(lib, descs, x, {0: w0, 1: w1, 2: w2}, _lib.PLAN_NO_AUTOTUNE, configs=cfgs)
    assert [c["row"] for c in cfgs] == [1, 0, 1]
    monkeypatch.setenv("RTOD_TC_ROW", "0")
    cfgs = []
    plain = run_block(lib, descs, x, {0: w0, 1: w1, 2: w2}, _lib.PLAN_NO_AUTOTUNE, configs=cfgs)
    assert [c["row"] for c in cfgs] == [0, 0, 0]
    tol = BLOCK_TOL["fp16"]
    assert frac_within(outs[0], ref_block(x, w0, 3, 1, True, "fp16"), *tol) == 1.0
    assert frac_within(outs[3], ref_block(outs[1], w2, 3, 1, True, "fp16") + outs[0], *tol) == 1.0
    for a, b in zip(outs, plain):
        assert torch.equal(a, b)


def test_conv_block_two_term_weights_every_launch_configuration(lib, monkeypatch):
    monkeypatch.setenv("RTOD_WSPLIT_ELEMS", "0")
    monkeypatch.setenv("RTOD_WSPLIT_AI", "10000")
    rng = np.random.RandomState(77)
    for cin, cout, k, stride, H, batch in [(64, 128, 3, 1, 52, 3), (32, 64, 3, 2, 104, 2), (256, 128, 1, 1, 52, 4)]:
        x = torch.from_numpy(rng.randn(batch, cin, H, H).astype(np.float32))
        w = rand_conv(rng, cin, cout, k)
        ref_q = ref_block(x, w, k, stride, True, "fp16", split=True)
        first = None
        for force in FORCED:
            monkeypatch.setenv("RTOD_TC_FORCE", force)
            splits = []
            got = run_block(lib, [conv_desc(cout, k, stride)], x, {0: w}, splits=splits)[0]
            assert splits == [1]
            assert frac_within(got, ref_q, *BLOCK_TOL["fp16"]) >= 0.9999, force
            if first is None:
                first = got
            assert torch.equal(got, first), force
        monkeypatch.delenv("RTOD_TC_FORCE")


CONV_TC_PAIR, CONV_TC = 3, 2                # include/rtod.h RTOD_CONV_*

# shapes large enough for the CTA-pair kernel (Cout % 256 == 0 and >= 60 pair tiles): an odd number of
# 128-row tiles (the peer CTA of the last pair is entirely out of bounds), 1x1, stride 2, two N tiles
# ... and Cout = 128 / 384 (N tile 128: each CTA loads 64 weight rows)
PAIR_CASES = [(128, 256, 3, 1, 52, 6), (256, 512, 1, 1, 26, 12), (128, 256, 3, 2, 104, 6), (64, 512, 3, 1, 26, 11),
              (64, 128, 3, 1, 104, 2), (256, 128, 1, 1, 52, 6), (64, 128, 3, 2, 208, 2), (32, 384, 3, 1, 52, 3)]


@pytest.mark.parametrize("cin,cout,k,stride,H,batch", PAIR_CASES)
@pytest.mark.parametrize("dtype,dflag", DTYPES)
def test_conv_block_cta_pair_kernel(lib, cin, cout, k, stride, H, batch, dtype, dflag):
    rng = np.random.RandomState(cin + cout + k + stride + H)
    x = torch.from_numpy(rng.randn(batch, cin, H, H).astype(np.float32))
    w = rand_conv(rng, cin, cout, k)
    backends = []
    # heuristic plan: the bind-time autotuner may prefer the one-CTA kernel at this (small) size
    got = run_block(lib, [conv_desc(cout, k, stride)], x, {0: w}, _lib.PLAN_NO_AUTOTUNE | _lib.PLAN_NO_WSPLIT | dflag,
                    backends)[0]
    assert backends == [CONV_TC_PAIR]           # the test is about conv_pair.cu, make sure it ran
    ref_q = ref_block(x, w, k, stride, True, dtype)
    assert got.shape == ref_q.shape
    assert frac_within(got, ref_q, *BLOCK_TOL[dtype]) == 1.0

@pytest.mark.parametrize("cin,cout,k,stride,H,batch", [(256, 128, 1, 1, 52, 6), (64, 128, 3, 2, 208, 2), (64, 256, 3, 1, 52, 6)])
def test_conv_block_cta_pair_kernel_two_term_weights(lib, monkeypatch, cin, cout, k, stride, H, batch):
    monkeypatch.setenv("RTOD_WSPLIT_ELEMS", "0")
    monkeypatch.setenv("RTOD_WSPLIT_AI", "10000")
    rng = np.random.RandomState(cin + cout + k + stride + H + 1)
    x = torch.from_numpy(rng.randn(batch, cin, H, H).astype(np.float32))
    w = rand_conv(rng, cin, cout, k)
    backends, splits = [], []
    got = run_block(lib, [conv_desc(cout, k, stride)], x, {0: w}, _lib.PLAN_NO_AUTOTUNE, backends, splits)[0]
    assert backends == [CONV_TC_PAIR] and splits == [1]
    ref_q = ref_block(x, w, k, stride, True, "fp16", split=True)
    assert frac_within(got, ref_q, *BLOCK_TOL["fp16"]) == 1.0


@pytest.mark.parametrize("dtype,dflag", DTYPES)
def test_residual_block_cta_pair_kernel(lib, dtype, dflag):
    """shortcut operand TMA-loaded in place by the per-warp epilogue, on the CTA-pair kernel"""
    rng = np.random.RandomState(23)
    x = torch.from_numpy(rng.randn(6, 64, 52, 52).astype(np.float32))
    w0, w1, w2 = rand_conv(rng, 64, 256, 3), rand_conv(rng, 256, 128, 1), rand_conv(rng, 128, 256, 3)
    sc = _lib.RtodLayerDesc()
    sc.type, sc.src0, sc.src1 = _lib.LAYER_SHORTCUT, 2, 0
    backends = []
    outs = run_block(lib, [conv_desc(256, 3, 1), conv_desc(128, 1, 1), conv_desc(256, 3, 1), sc], x,
                     {0: w0, 1: w1, 2: w2}, _lib.PLAN_NO_AUTOTUNE | _lib.PLAN_NO_WSPLIT | dflag, backends)
    assert backends[0] == CONV_TC_PAIR and backends[2] == CONV_TC_PAIR and backends[1] in (CONV_TC, CONV_TC_PAIR)
    # every stage is checked against the reference evaluated on the DEVICE's own stored input of that stage
    tol = BLOCK_TOL[dtype]
    assert frac_within(outs[0], ref_block(x, w0, 3, 1, True, dtype), *tol) == 1.0
    assert frac_within(outs[1], ref_block(outs[0], w1, 1, 1, True, dtype), *tol) == 1.0
    assert frac_within(outs[3], ref_block(outs[1], w2, 3, 1, True, dtype) + outs[0], *tol) == 1.0
    assert torch.equal(outs[2], outs[3])


# every launch configuration the bind-time autotuner may pick must give the same (correct) result:
# "pair,bn,ctas,resident,sbufs" is forced through RTOD_TC_FORCE; combinations that do not fit a shape fall
# back to the default, which is covered anyway
FORCED = ["0,%d,%d,%d,%d" % (bn, c, r, s) for bn in (256, 128, 64) for c in (1, 2, 3) for r in (0, 1) for s in (1, 2)]
FORCED.append("1,256,1,0,2")
# dual pipeline (two warp sets sharing resident weights): "pair,bn,ctas,resident,sbufs,split,pipelines"
FORCED += ["0,%d,1,1,%d,0,2" % (bn, s) for bn in (128, 64, 32) for s in (1, 2)]


@pytest.mark.parametrize("cin,cout,k,stride,H,batch", [(64, 128, 3, 1, 52, 3), (32, 64, 3, 2, 104, 2),
                                                       (256, 256, 1, 1, 26, 12), (128, 256, 3, 1, 52, 6)])
def test_conv_block_every_launch_configuration(lib, monkeypatch, cin, cout, k, stride, H, batch):
    rng = np.random.RandomState(cin * 3 + cout + k + stride)
    x = torch.from_numpy(rng.randn(batch, cin, H, H).astype(np.float32))
    w = rand_conv(rng, cin, cout, k)
    ref_q = ref_block(x, w, k, stride, True, "fp16")
    first = None
    for force in FORCED:
        monkeypatch.setenv("RTOD_TC_FORCE", force)
        got = run_block(lib, [conv_desc(cout, k, stride)], x, {0: w}, _lib.PLAN_NO_WSPLIT)[0]
        assert frac_within(got, ref_q, *BLOCK_TOL["fp16"]) >= 0.9999, force   # (a rounding flip in a million elements is possible)
        if first is None:
            first = got
        assert torch.equal(got, first), force           # same accumulation order in every configuration
    monkeypatch.delenv("RTOD_TC_FORCE")


@pytest.mark.parametrize("cin,cout,k,stride,H,batch", [(512, 1024, 3, 1, 13, 1), (256, 512, 3, 1, 26, 1),
                                                       (1024, 512, 1, 1, 13, 2), (128, 256, 3, 2, 52, 1)])
@pytest.mark.parametrize("split", [2, 4, 8])
def test_conv_block_split_k(lib, monkeypatch, cin, cout, k, stride, H, batch, split):
    """small batches: K split over several CTAs per tile, partials summed in slice order by the last arriver;
    run twice (the arrival counters must be back at zero) and with a shortcut operand"""
    rng = np.random.RandomState(cin + cout + k + split)
    x = torch.from_numpy(rng.randn(batch, cin, H, H).astype(np.float32))
    w = rand_conv(rng, cin, cout, k)
    ref_q = ref_block(x, w, k, stride, True, "fp16")
    monkeypatch.setenv("RTOD_TC_FORCE", "0,32,2,0,2,%d" % split)
    got = run_block(lib, [conv_desc(cout, k, stride)], x, {0: w})[0]
    assert frac_within(got, ref_q, *BLOCK_TOL["fp16"]) >= 0.9999
    monkeypatch.setenv("RTOD_TC_FORCE", "0,64,1,0,2,%d" % split)
    got2 = run_block(lib, [conv_desc(cout, k, stride)], x, {0: w})[0]
    assert frac_within(got2, ref_q, *BLOCK_TOL["fp16"]) >= 0.9999
    monkeypatch.delenv("RTOD_TC_FORCE")


def test_residual_block_split_k(lib, monkeypatch):
    rng = np.random.RandomState(29)
    x = torch.from_numpy(rng.randn(1, 512, 13, 13).astype(np.float32))
    w0, w1, w2 = rand_conv(rng, 512, 1024, 3), rand_conv(rng, 1024, 512, 1), rand_conv(rng, 512, 1024, 3)
    sc = _lib.RtodLayerDesc()
    sc.type, sc.src0, sc.src1 = _lib.LAYER_SHORTCUT, 2, 0
    monkeypatch.setenv("RTOD_TC_FORCE", "0,32,2,0,2,4")
    outs = run_block(lib, [conv_desc(1024, 3, 1), conv_desc(512, 1, 1), conv_desc(1024, 3, 1), sc], x,
                     {0: w0, 1: w1, 2: w2})
    monkeypatch.delenv("RTOD_TC_FORCE")
    tol = BLOCK_TOL["fp16"]
    assert frac_within(outs[0], ref_block(x, w0, 3, 1, True, "fp16"), *tol) >= 0.9999
    assert frac_within(outs[1], ref_block(outs[0], w1, 1, 1, True, "fp16"), *tol) >= 0.9999
    assert frac_within(outs[3], ref_block(outs[1], w2, 3, 1, True, "fp16") + outs[0], *tol) >= 0.9999
    assert torch.equal(outs[2], outs[3])


def test_conv_head_keeps_fp32_logits(lib):
    """A convolution that only feeds a yolo layer stores fp32 (255 channels, no activation) and the
    single decode launch turns it into prediction rows."""
    rng = np.random.RandomState(5)
    x = torch.from_numpy(rng.randn(2, 256, 13, 13).astype(np.float32))
    w = rand_conv(rng, 256, 255, 1, bn=False)
    yolo = _lib.RtodLayerDesc()
    yolo.type, yolo.num_anchors, yolo.classes, yolo.src0, yolo.src1 = _lib.LAYER_YOLO, 3, 80, -1, -1
    for k, v in enumerate([116, 90, 156, 198, 373, 326]):
        yolo.anchors[k] = v
    arr = (_lib.RtodLayerDesc * 2)(conv_desc(255, 1, 1, bn=False, leaky=False), yolo)
    plan = ctypes.c_void_p()
    _lib.check(lib.rtod_plan_create(arr, 2, 2, 256, 13, 13, 416, 0, ctypes.byref(plan)))
    assert lib.rtod_plan_num_rows(plan) == 507 and lib.rtod_plan_num_attrs(plan) == 85
    ws = torch.empty(lib.rtod_plan_workspace_bytes(plan) + 256, dtype=torch.uint8, device="cuda")
    wa = torch.empty(lib.rtod_plan_weight_bytes(plan) + 256, dtype=torch.uint8, device="cuda")
    _lib.check(lib.rtod_plan_bind(plan, _aligned(ws), lib.rtod_plan_workspace_bytes(plan), _aligned(wa),
                                  lib.rtod_plan_weight_bytes(plan)))
    wd, bd = w["w"].cuda(), w["b"].cuda()
    _lib.check(lib.rtod_plan_set_conv_weights(plan, 0, wd.data_ptr(), bd.data_ptr(), None, None, None, None,
                                              0.0, None))
    xd = x.cuda()
    pred = torch.empty(2, 507, 85, device="cuda")
    _lib.check(lib.rtod_plan_forward(plan, xd.data_ptr(), pred.data_ptr(), 0, None))
    _lib.check(lib.rtod_plan_check(plan, None))
    logits = F.conv2d(x.half().float(), w["w"].half().float(), w["b"])
    want = oracle.predict_transform(logits, 416, [(116, 90), (156, 198), (373, 326)], 80, False)
    np.testing.assert_allclose(pred.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-4)
    lib.rtod_plan_destroy(plan)


def test_residual_block_fuses_shortcut(lib):
    """conv3x3 -> conv1x1 -> conv3x3 -> shortcut(-3): the add happens in the last convolution's
    epilogue and the shortcut layer is an alias of it."""
    rng = np.random.RandomState(11)
    x = torch.from_numpy(rng.randn(2, 64, 20, 20).astype(np.float32))
    w0, w1, w2 = rand_conv(rng, 64, 128, 3), rand_conv(rng, 128, 64, 1), rand_conv(rng, 64, 128, 3)
    sc = _lib.RtodLayerDesc()
    sc.type, sc.src0, sc.src1 = _lib.LAYER_SHORTCUT, 2, 0
    descs = [conv_desc(128, 3, 1), conv_desc(64, 1, 1), conv_desc(128, 3, 1), sc]
    for dtype, dflag in DTYPES:
        for flags in (0, _lib.PLAN_CONV_SIMT):
            outs = run_block(lib, descs, x, {0: w0, 1: w1, 2: w2}, flags | dflag)
            tol = BLOCK_TOL[dtype]
            assert frac_within(outs[0], ref_block(x, w0, 3, 1, True, dtype), *tol) == 1.0
            assert frac_within(outs[1], ref_block(outs[0], w1, 1, 1, True, dtype), *tol) == 1.0
            assert frac_within(outs[3], ref_block(outs[1], w2, 3, 1, True, dtype) + outs[0], *tol) == 1.0
            assert torch.equal(outs[2], outs[3])


