#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

For every case the script (1) runs the reference, (2) runs the oracle port on the same
input and asserts ``torch.equal`` -- this is what pins the oracle -- and (3) stores inputs
(or the seeds they are regenerated from) and the reference's outputs as small .npz
fixtures.  Weights for the forward cases are NOT stored: they are regenerated from
``realtimeobjectdetection_b200.synth.synth_stream(seed, mode)`` (numpy RandomState, stable
across machines).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from src.darknet import Darknet as RefDarknet            # noqa: E402  (the reference)
from src import util as ref_util                         # noqa: E402

import oracle                                            # noqa: E402
from realtimeobjectdetection_b200 import synth           # noqa: E402
from realtimeobjectdetection_b200.cfg import builtin_cfg, parse_cfg  # noqa: E402

torch.set_num_threads(max(1, os.cpu_count() or 1))
SUMMARY = {"torch": torch.__version__, "numpy": np.__version__, "cases": {}}


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    SUMMARY["cases"][name] = {k: list(np.shape(v)) for k, v in arrays.items()}
    print("wrote %-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


# ---------------------------------------------------------------- cfg ---------------
def golden_cfg():
    blocks = {}
    for name in ("yolov3", "yolov3-tiny"):
        ref_blocks = RefDarknet.parse_cfg("/root/reference/cfg/%s.cfg" % name)
        mine = parse_cfg(builtin_cfg(name))
        assert mine == ref_blocks, name
        assert oracle.parse_cfg(builtin_cfg(name)) == ref_blocks, name
        assert parse_cfg("/root/reference/cfg/%s.cfg" % name) == ref_blocks
        blocks[name] = ref_blocks
    with open(os.path.join(HERE, "cfg_blocks.json"), "w") as fh:
        json.dump(blocks, fh, indent=0, sort_keys=True)
    print("wrote cfg_blocks.json")


# ---------------------------------------------------------------- decode ------------
def golden_decode():
    rng = np.random.RandomState(11)
    cases = {
        # SURVEY.md section 4 decode KAT
        "kat": (np.zeros((1, 21, 2, 2), np.float32), 64, [(10, 13), (16, 30), (33, 23)], 2, False),
        "g7c80": ((rng.randn(2, 255, 7, 7) * 1.5).astype(np.float32), 224,
                  [(116, 90), (156, 198), (373, 326)], 80, False),
        "g13c20": ((rng.randn(1, 75, 13, 13) * 2.0).astype(np.float32), 416,
                   [(10, 13), (16, 30), (33, 23)], 20, False),
        "g10c3a2": ((rng.randn(3, 16, 10, 10)).astype(np.float32), 320,
                    [(81, 82), (135, 169)], 3, False),
        "train_g7c80": ((rng.randn(2, 255, 7, 7) * 1.5).astype(np.float32), 224,
                        [(116, 90), (156, 198), (373, 326)], 80, True),
    }
    cases["kat"][0][0, 0, 0, 1] = 1.0
    cases["kat"][0][0, 9, 1, 0] = 0.5
    for name, (x, inp_dim, anchors, ncls, train) in cases.items():
        xt = torch.from_numpy(x)
        keep = xt.clone()
        want = ref_util.predict_transform(xt, inp_dim, anchors, ncls, False, TRAIN=train)
        assert torch.equal(xt, keep), "reference mutated its input"
        got = oracle.predict_transform(torch.from_numpy(x), inp_dim, anchors, ncls, False, TRAIN=train)
        assert torch.equal(want, got), name
        save("decode_" + name, x=x, inp_dim=np.int64(inp_dim), anchors=np.array(anchors, np.int64),
             num_class=np.int64(ncls), train=np.int64(train), out=want.numpy())


# ---------------------------------------------------------------- NMS ---------------
def synth_pred(rng, B, N, C, density, clustered, span=416.0):
    """[B,N,5+C] prediction tensor with exactly round(density*N) rows above 0.5 per image,
    all objectness values distinct."""
    pred = np.zeros((B, N, 5 + C), np.float32)
    K = int(round(density * N))
    for b in range(B):
        pred[b, :, 0:2] = rng.uniform(0, span, size=(N, 2))
        pred[b, :, 2:4] = np.exp(rng.uniform(2, 5, size=(N, 2)))
        pred[b, :, 5:] = rng.uniform(0, 1, size=(N, C))
        hi = 0.5 + 0.4995 * (np.arange(K) + 0.5) / max(K, 1)
        lo = 0.4995 * (np.arange(N - K) + 0.5) / max(N - K, 1)
        obj = np.concatenate([hi, lo]).astype(np.float32)
        perm = rng.permutation(N)
        pred[b, perm, 4] = obj
        if clustered and K > 0:
            n_obj = 12
            centres = rng.uniform(40, span - 40, size=(n_obj, 2))
            sizes = np.exp(rng.uniform(3, 5, size=(n_obj, 2)))
            classes = rng.randint(0, C, size=n_obj)
            rows = perm[:K]
            which = rng.randint(0, n_obj, size=K)
            pred[b, rows, 0:2] = centres[which] + rng.randn(K, 2) * 0.15 * sizes[which]
            pred[b, rows, 2:4] = sizes[which] * np.exp(rng.randn(K, 2) * 0.15)
            pred[b, rows, 5:] *= 0.5
            pred[b, rows, 5 + classes[which]] = rng.uniform(0.6, 1.0, size=K)
    return pred


def golden_nms():
    rng = np.random.RandomState(7)
    kat = np.zeros((2, 6, 8), np.float32)
    kat[0] = [[100, 100, 50, 50, .90, .1, .8, .2], [104, 100, 50, 50, .80, .1, .7, .2],
              [300, 300, 40, 60, .70, .9, .1, .2], [100, 100, 50, 50, .50, .1, .8, .2],
              [160, 100, 50, 50, .95, .1, .6, .2], [10, 10, 5, 5, .99, .3, .3, .3]]
    kat[1] = kat[0]
    kat[1, :, 4] = [.45, .40, .35, .25, .30, .10]
    single = synth_pred(rng, 2, 64, 5, 0.0, False)
    single[1, 17, 4] = 0.93                                        # exactly one surviving row
    cases = {
        "kat": (kat, 3, 0.5, 0.4),
        "uniform10_c20": (synth_pred(rng, 2, 507, 20, 0.10, False), 20, 0.5, 0.4),
        "uniform50_c20": (synth_pred(rng, 2, 507, 20, 0.50, False), 20, 0.5, 0.4),
        "cluster50_c20": (synth_pred(rng, 2, 507, 20, 0.50, True), 20, 0.5, 0.4),
        "cluster10_c80": (synth_pred(rng, 3, 300, 80, 0.10, True), 80, 0.5, 0.45),
        "cluster100_c4": (synth_pred(rng, 1, 700, 4, 1.00, True), 4, 0.25, 0.3),
        "none": (synth_pred(rng, 2, 64, 5, 0.0, False), 5, 0.5, 0.4),
        "single": (single, 5, 0.6, 0.4),
    }
    for name, (pred, ncls, conf, nms) in cases.items():
        pt = torch.from_numpy(pred)
        keep = pt.clone()
        want = ref_util.write_results(pt, ncls, conf, nms)
        assert torch.equal(pt, keep), "reference mutated its input"
        got = oracle.write_results(torch.from_numpy(pred), ncls, conf, nms)
        if isinstance(want, int):
            assert isinstance(got, int) and got == want == 0, name
            out = np.zeros((0, 8), np.float32)
        else:
            assert torch.equal(want, got), name
            out = want.numpy()
        save("nms_" + name, pred=pred, num_class=np.int64(ncls), conf=np.float64(conf),
             nms=np.float64(nms), out=out, is_zero=np.int64(isinstance(want, int)))
        print("   %-16s kept %d" % (name, out.shape[0]))

    # bbox_iou broadcast KAT
    b1 = (rng.uniform(0, 400, size=(1, 7))).astype(np.float32)
    b1[:, 2:4] = b1[:, 0:2] + rng.uniform(5, 80, size=(1, 2))
    b2 = (rng.uniform(0, 400, size=(257, 7))).astype(np.float32)
    b2[:, 2:4] = b2[:, 0:2] + rng.uniform(5, 200, size=(257, 2))
    want = ref_util.bbox_iou(torch.from_numpy(b1), torch.from_numpy(b2))
    assert torch.equal(want, oracle.bbox_iou(torch.from_numpy(b1), torch.from_numpy(b2)))
    save("iou_broadcast", box1=b1, box2=b2, out=want.numpy())

    # The reference's only shipped golden artefact: det/metrics.json (real weights, conf 0.6,
    # nms 0.5).  Re-feeding its rows (one-hot class scores) must return them unchanged.
    with open("/root/reference/det/metrics.json") as fh:
        det = json.load(fh)
    rows = []
    names = sorted(det)
    for k, name in enumerate(names):
        if det[name] == 0:
            continue
        for r in det[name]:
            rows.append([k] + list(r[1:]))
    rows = np.array(rows, np.float32)                                # [D,8], img = ordinal
    B, C = len(names), 80
    per_img = max(int((rows[:, 0] == k).sum()) for k in range(B))
    pred = np.zeros((B, per_img, 5 + C), np.float32)
    fill = [0] * B
    for r in rows:
        b = int(r[0])
        j = fill[b]
        fill[b] += 1
        w, h = r[3] - r[1], r[4] - r[2]
        pred[b, j, 0:4] = [r[1] + w / 2, r[2] + h / 2, w, h]
        pred[b, j, 4] = r[5]
        pred[b, j, 5 + int(r[7])] = r[6]
    want = ref_util.write_results(torch.from_numpy(pred), C, 0.6, 0.5)
    got = oracle.write_results(torch.from_numpy(pred), C, 0.6, 0.5)
    assert torch.equal(want, got)
    save("nms_ref_metrics_json", pred=pred, num_class=np.int64(C), conf=np.float64(0.6),
         nms=np.float64(0.5), out=want.numpy(), is_zero=np.int64(0), shipped_rows=rows)
    print("   metrics.json rows %d -> kept %d" % (rows.shape[0], want.shape[0]))


# ---------------------------------------------------------------- forward -----------
def golden_forward():
    cases = [  # (name, cfg, reso, batch, weight seed, weight mode, input seed)
        ("yolov3_128_cal", "yolov3", 128, 2, 3, "calibrated", 21),
        ("yolov3_128_def", "yolov3", 128, 1, 4, "default", 22),
        ("tiny_160_cal", "yolov3-tiny", 160, 2, 5, "calibrated", 23),
        ("tiny_128_def", "yolov3-tiny", 128, 1, 6, "default", 24),
    ]
    for name, cfg_name, reso, batch, wseed, mode, xseed in cases:
        cfg = builtin_cfg(cfg_name)
        blocks = parse_cfg(cfg)
        stream = synth.synth_stream(blocks, wseed, mode)
        wpath = "/tmp/golden_%s.weights" % name
        synth.write_weights_file(wpath, stream)
        ref = RefDarknet("/root/reference/cfg/%s.cfg" % cfg_name, False)
        ref.load_weights(wpath)
        ref.eval()                                     # parity oracle = eval-mode BN
        ref.net_info["height"] = reso
        x = np.random.RandomState(xseed).rand(batch, 3, reso, reso).astype(np.float32)
        with torch.no_grad():
            want = ref(torch.from_numpy(x))
        state = {k: torch.from_numpy(v) for k, v in synth.stream_to_state(blocks, stream).items()}
        for k, v in ref.state_dict().items():
            if k.endswith("num_batches_tracked"):
                continue
            assert torch.equal(v, state[k]), k         # stream_to_state == load_weights
        port = oracle.DarknetPort(cfg, state)
        port.net_info["height"] = reso
        with torch.no_grad():
            got = port(torch.from_numpy(x))
        assert torch.equal(want, got), name
        assert port.anchors == ref.anchors and port.num_classes == ref.num_classes
        det = ref_util.write_results(want.clone(), 80, 0.5, 0.4)
        det_rows = np.zeros((0, 8), np.float32) if isinstance(det, int) else det.numpy()
        save("forward_" + name, cfg=np.array(cfg_name), reso=np.int64(reso), batch=np.int64(batch),
             weight_seed=np.int64(wseed), weight_mode=np.array(mode), input_seed=np.int64(xseed),
             weight_checksum=np.float64(np.abs(stream.astype(np.float64)).sum()),
             pred=want.numpy(), det=det_rows)
        print("   %-16s pred %s det %d obj>0.5 %.4f" % (name, tuple(want.shape), det_rows.shape[0],
                                                      float((want[..., 4] > 0.5).float().mean())))
        os.remove(wpath)


# ---------------------------------------------------------------- pre-processing ---------
def golden_prep():
    """util.prep_image / letterbox_image (src/util.py:349-397) on the reference's own images.  cv2 is a
    third-party dependency (this container: the stock wheel, which routes 8-bit cubic resizes to IPP):
    the oracle's "opencv" mode must equal the reference bit for bit with IPP switched off, its "float" mode
    must stay within one grey level of the IPP result on a vanishing fraction of the values."""
    import glob
    import cv2
    from oracle import prep_port
    worst_frac = 0.0
    for path in sorted(glob.glob("/root/reference/imgs/*.jpg")):
        img = cv2.imread(path)
        for dim in (416, 608, 320):
            cv2.ipp.setUseIPP(False)
            want = ref_util.prep_image(img, dim)
            assert torch.equal(want, prep_port.prep_image(img, dim, resize="opencv")), (path, dim)
            assert np.array_equal(ref_util.letterbox_image(img, (dim, dim)),
                                  prep_port.letterbox_image(img, (dim, dim), "opencv")), (path, dim)
            cv2.ipp.setUseIPP(True)
            ipp = ref_util.prep_image(img, dim)
            mine = prep_port.prep_image(img, dim, resize="float")
            diff = (ipp - mine).abs() * 255.0
            assert float(diff.max()) <= 1.0001, (path, dim)
            worst_frac = max(worst_frac, float((diff > 0.5).float().mean()))
    assert worst_frac < 5e-4, worst_frac
    print("   prep: oracle == reference (IPP off) on 11 images x 3 sizes; float mode vs IPP: one grey level on "
          "<= %.4f %% of the values" % (100 * worst_frac))
    # committed fixtures: two of the reference's images (one portrait, one landscape), their canvases from
    # cv2 with IPP off (the oracle's exact target) and where the IPP-backed result differs from it
    for name, dim in (("img1", 416), ("img4", 320)):
        img = cv2.imread("/root/reference/imgs/%s.jpg" % name)
        cv2.ipp.setUseIPP(False)
        canvas = ref_util.letterbox_image(img, (dim, dim)).astype(np.uint8)
        rgb = ref_util.prep_image(img, dim)
        assert torch.equal(rgb, torch.from_numpy(canvas[:, :, ::-1].transpose(2, 0, 1).copy()).float().div(255.0).unsqueeze(0))
        cv2.ipp.setUseIPP(True)
        canvas_ipp = ref_util.letterbox_image(img, (dim, dim)).astype(np.uint8)
        where = np.flatnonzero(canvas_ipp != canvas).astype(np.int32)
        save("prep_%s_%d" % (name, dim), img=img, inp_dim=np.int64(dim), canvas=canvas,
             ipp_index=where, ipp_value=canvas_ipp.reshape(-1)[where])


# ---------------------------------------------------------------- post-processing --------
def golden_post():
    """detect.py:120-136 (box rescale + clamp) by calling the reference's own methods, and test.py:139-151
    (validator IoU matrix; test.py itself needs matplotlib and cannot be imported: the loop is replayed here
    around the reference's bbox_iou)."""
    import detect as ref_detect
    from oracle import post_port
    rng = np.random.RandomState(5)

    class Stub:
        pass

    for name, inp_dim, n_img, D in (("416", 416, 5, 40), ("608", 608, 3, 25), ("320", 320, 1, 7)):
        dims = np.stack([rng.randint(200, 1400, n_img), rng.randint(150, 1100, n_img)], 1).astype(np.float32)
        im_dim_list = torch.from_numpy(dims).repeat(1, 2)                       # detect.py:247-248
        rows = np.zeros((D, 8), np.float32)
        rows[:, 0] = np.sort(rng.randint(0, n_img, D))
        c = rng.uniform(-20, inp_dim + 20, (D, 2)); wh = np.exp(rng.uniform(2, 5.5, (D, 2)))
        rows[:, 1:3], rows[:, 3:5] = c - wh / 2, c + wh / 2
        rows[:, 5:7] = rng.uniform(0.5, 1, (D, 2)); rows[:, 7] = rng.randint(0, 80, D)
        out = torch.from_numpy(rows.copy())
        stub = Stub()
        stub.inp_dim = inp_dim
        sel = ref_detect.Darknetv3Detector.convert_box_dims_to_original_image(stub, 0, im_dim_list.clone(), out)
        ref_detect.Darknetv3Detector.clamp_box_dims(stub, sel, out)
        got, got_dims = post_port.rescale_boxes(torch.from_numpy(rows), im_dim_list, inp_dim)
        assert torch.equal(got, out) and torch.equal(got_dims, sel), name
        save("rescale_" + name, rows=rows, im_dim_list=im_dim_list.numpy(), inp_dim=np.int64(inp_dim), out=out.numpy(),
             dims=sel.numpy())

    for name, P, T, thr in (("p12_t7", 12, 7, 0.5), ("p40_t1", 40, 1, 0.4), ("p3_t50", 3, 50, 0.0)):
        def boxes(n):
            c = rng.uniform(0, 416, (n, 2)); wh = np.exp(rng.uniform(3, 5, (n, 2)))
            return np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        tb = boxes(T)
        pb = boxes(P)
        pb[: min(P, T)] = tb[: min(P, T)] + rng.randn(min(P, T), 4).astype(np.float32) * 6     # overlapping pairs
        pred = np.concatenate([np.zeros((P, 1), np.float32), pb, rng.uniform(0, 1, (P, 3)).astype(np.float32)], 1)
        target = np.concatenate([tb, np.ones((T, 1), np.float32)], 1)
        ious = []
        for box in torch.from_numpy(pred):                                      # test.py:141-149
            row = []
            for t_box in torch.from_numpy(target):
                iou = ref_util.bbox_iou(box[1:5].cpu(), t_box[0:4].cpu())
                row.append(iou.item() if iou.item() > thr else 0.0)
            ious.append(row)
        want = torch.FloatTensor(ious)
        assert torch.equal(want, post_port.iou_matrix(torch.from_numpy(pred), torch.from_numpy(target), thr)), name
        save("ioumat_" + name, pred=pred, target=target, threshold=np.float64(thr), out=want.numpy())


if __name__ == "__main__":
    golden_cfg()
    golden_decode()
    golden_nms()
    golden_forward()
    golden_prep()
    golden_post()
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as fh:
        json.dump(SUMMARY, fh, indent=1, sort_keys=True)
    print("oracle == reference on every case; fixtures written to", HERE)
