"""GPU parity tests of the callers either side of the hot path: device
pre-processing, box rescale / clamp, validator IoU matrix -- bit-exact against the committed reference vectors
and the oracle ports."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from oracle import post_port, prep_port
from realtimeobjectdetection_b200 import (bbox_iou_matrix, letterbox_image, metrics_rows, prep_frames, prep_image,
                                          rescale_boxes)
from realtimeobjectdetection_b200.util import RESIZE_FLOAT, RESIZE_OPENCV, letterbox_geometry

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names("prep_"))
def test_prep_image_against_reference_vectors(name):
    g = load_golden(name)
    img, dim = g["img"], int(g["inp_dim"])
    want = torch.from_numpy(g["canvas"][:, :, ::-1].transpose(2, 0, 1).copy()).float().div(255.0).unsqueeze(0)
    got = prep_image(img, dim, resize=RESIZE_OPENCV)                   # numpy in -> host tensor out, like the reference
    assert not got.is_cuda and got.shape == want.shape
    assert torch.equal(got, want)                                      # bit-exact: cv2's own fixed-point path
    assert np.array_equal(letterbox_image(img, (dim, dim), RESIZE_OPENCV), g["canvas"].astype(np.int64))
    assert torch.equal(prep_image(img, dim, mode="RGB", resize=RESIZE_OPENCV), want.flip(1))
    # default mode = fp32 weights: the stock cv2 wheel's IPP result to within one grey level on a few pixels
    ipp = g["canvas"].copy().reshape(-1)
    ipp[g["ipp_index"]] = g["ipp_value"]
    ipp = torch.from_numpy(ipp.reshape(dim, dim, 3)[:, :, ::-1].transpose(2, 0, 1).copy()).float().div(255.0)
    diff = (prep_image(img, dim)[0] - ipp).abs() * 255.0
    assert float(diff.max()) <= 1.0001 and float((diff > 0.5).float().mean()) < 5e-4
    assert torch.equal(prep_image(img, dim), prep_port.prep_image(img, dim, resize="float"))   # ... and == the oracle


@pytest.mark.parametrize("h,w,dim", [(480, 640, 416), (1080, 1920, 608), (416, 416, 416), (97, 33, 320), (300, 301, 160)])
def test_prep_frames_against_oracle(h, w, dim):
    rng = np.random.RandomState(h + w)
    frames = rng.randint(0, 256, (3, h, w, 3)).astype(np.uint8)
    for resize, mode in ((RESIZE_OPENCV, "opencv"), (RESIZE_FLOAT, "float")):
        got = prep_frames(torch.from_numpy(frames).cuda(), dim, resize=resize)
        assert got.is_cuda and got.shape == (3, 3, dim, dim)
        for b in range(3):
            assert torch.equal(got[b:b + 1].cpu(), prep_port.prep_image(frames[b], dim, resize=mode)), (b, mode)
    u8 = prep_frames(torch.from_numpy(frames), dim, resize=RESIZE_OPENCV, as_uint8=True)
    # (on the host: CUDA's div-by-scalar multiplies with the reciprocal, prep_image's is a true division)
    assert u8.dtype == torch.uint8 and torch.equal(u8.cpu().float().div(255.0),
                                                   prep_frames(torch.from_numpy(frames), dim, resize=RESIZE_OPENCV).cpu())
    assert letterbox_geometry(w, h, dim) == prep_port.letterbox_geometry(w, h, dim, dim)
    if h == w == dim:                                                  # same size: the resize is the identity
        assert torch.equal(u8.cpu(), torch.from_numpy(frames).flip(3).permute(0, 3, 1, 2))


@pytest.mark.parametrize("name", golden_names("rescale_"))
def test_rescale_boxes_against_reference_vectors(name):
    g = load_golden(name)
    rows = torch.from_numpy(g["rows"]).cuda()
    keep = rows.clone()
    out, dims = rescale_boxes(rows, torch.from_numpy(g["im_dim_list"]), int(g["inp_dim"]))
    assert torch.equal(rows, keep) and out.is_cuda
    assert torch.equal(out.cpu(), torch.from_numpy(g["out"])) and torch.equal(dims.cpu(), torch.from_numpy(g["dims"]))
    assert metrics_rows(out) == g["out"].tolist() and metrics_rows(0) == 0
    # ref_dim = inp_dim (the evident intent of detect.py:130) against the oracle
    out2, _ = rescale_boxes(rows, torch.from_numpy(g["im_dim_list"]), int(g["inp_dim"]), ref_dim=int(g["inp_dim"]))
    want2, _ = post_port.rescale_boxes(rows.cpu(), torch.from_numpy(g["im_dim_list"]), int(g["inp_dim"]), int(g["inp_dim"]))
    assert torch.equal(out2.cpu(), want2)


@pytest.mark.parametrize("name", golden_names("ioumat_"))
def test_iou_matrix_against_reference_vectors(name):
    g = load_golden(name)
    out = bbox_iou_matrix(torch.from_numpy(g["pred"]).cuda(), torch.from_numpy(g["target"]).cuda(), float(g["threshold"]))
    assert torch.equal(out.cpu(), torch.from_numpy(g["out"]))


def test_iou_matrix_large_against_oracle():
    rng = np.random.RandomState(9)
    c = rng.uniform(0, 608, (700, 2)); wh = np.exp(rng.uniform(2, 5, (700, 2)))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    pred = torch.from_numpy(np.concatenate([np.zeros((300, 1), np.float32), boxes[:300], np.ones((300, 3), np.float32)], 1))
    target = torch.from_numpy(boxes[300:])
    from oracle import bbox_iou
    want = bbox_iou(pred[:, None, 1:5], target[None, :, :4])                 # broadcasting form of the same arithmetic
    assert torch.equal(bbox_iou_matrix(pred, target), want)
    got = bbox_iou_matrix(pred.cuda(), target.cuda(), 0.3).cpu()
    assert torch.equal(got, torch.where(want.double() > 0.3, want, torch.zeros_like(want)))
    assert bbox_iou_matrix(pred[:0], target).shape == (0, 400)


@pytest.mark.parametrize("n_rows,count,capacity,first", [(50, 17, 40, 128), (50, 0, 8, 0), (10, 10, 10, 64), (30, 45, 20, 3), (0, 0, 5, 7)])
def test_pack_detections_payload(n_rows, count, capacity, first):
    """rtod_pack_detections (the per-step gather payload of a rank): row 0 carries the count, rows 1.. the first
    min(count, n_rows, capacity) detections with the image column shifted, zeros beyond -- what the tensor expression
    of sharding.gather_detections_async builds (which multiplies stale rows by 0 instead of overwriting them)"""
    from realtimeobjectdetection_b200 import _lib
    rng = np.random.RandomState(n_rows + count)
    rows = torch.from_numpy(rng.rand(max(n_rows, 1), 8).astype(np.float32)).cuda()
    cnt = torch.tensor([count], dtype=torch.int32, device="cuda")
    payload = torch.full((capacity + 1, 8), float("nan"), device="cuda")
    n_src = min(capacity, n_rows)
    _lib.check(_lib.load().rtod_pack_detections(rows.data_ptr(), n_src, cnt.data_ptr(), float(first), capacity, payload.data_ptr(), None))
    want = torch.zeros(capacity + 1, 8)
    want[0, 0] = float(count)
    n = min(count, n_src)
    want[1:1 + n] = rows[:n].cpu()
    want[1:1 + n, 0] += float(first)
    assert torch.equal(payload.cpu(), want)
