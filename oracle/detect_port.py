"""Oracle (test infrastructure): CPU restatement of the reference's YOLO decode and NMS.

Every function names the reference lines it follows.  The arithmetic is done with the
same PyTorch fp32 ops in the same order as the reference so that results are
bit-identical on CPU (checked by tests/golden/make_golden.py against the reference
itself, and by tests/test_oracle_golden.py against the committed vectors).
"""
from __future__ import annotations

import torch


def confidence_mask(pred: torch.Tensor, confidence: float) -> torch.Tensor:
    """src/util.py:106-117 -- zero every row whose objectness is not > confidence.

    The comparison is strict and done in fp32 (the python scalar is cast down).
    Returns a new tensor; the argument is untouched.
    """
    keep = (pred[:, :, 4] > confidence).to(torch.float32)
    return pred * keep.unsqueeze(2)


def bbox_iou(box1: torch.Tensor, box2: torch.Tensor) -> torch.Tensor:
    """src/util.py:120-153 -- broadcasting IoU with the +1 pixel convention.

    Each intermediate is rounded to fp32 separately (no fused multiply-add):
    inter = clamp(min(x2)-max(x1)+1, 0) * clamp(min(y2)-max(y1)+1, 0);
    area  = (x2-x1+1)*(y2-y1+1);  iou = inter / (area1 + area2 - inter).
    """
    ax1, ay1, ax2, ay2 = box1[..., 0], box1[..., 1], box1[..., 2], box1[..., 3]
    bx1, by1, bx2, by2 = box2[..., 0], box2[..., 1], box2[..., 2], box2[..., 3]

    left = torch.max(ax1, bx1)
    top = torch.max(ay1, by1)
    right = torch.min(ax2, bx2)
    bottom = torch.min(ay2, by2)

    inter_w = torch.clamp(right - left + 1, min=0)
    inter_h = torch.clamp(bottom - top + 1, min=0)
    inter = inter_w * inter_h

    area_a = (ax2 - ax1 + 1) * (ay2 - ay1 + 1)
    area_b = (bx2 - bx1 + 1) * (by2 - by1 + 1)
    return inter / (area_a + area_b - inter)


def predict_transform(head: torch.Tensor, inp_dim: int, anchors, num_class: int,
                      CUDA: bool = False, TRAIN: bool = False) -> torch.Tensor:
    """src/util.py:175-239 -- one YOLO head, NCHW [B, A*(5+C), G, G] -> [B, G*G*A, 5+C].

    Row order is cell (y*G+x) major, anchor minor; columns cx, cy, w, h, obj, classes.
    Exact op order: cx = (sigmoid(tx) + x) * stride, w = (exp(tw) * f32(aw/stride)) * stride,
    where aw/stride is evaluated in python double precision and then rounded to fp32.
    TRAIN=True stops after the sigmoids (src/util.py:211).
    """
    batch = head.size(0)
    stride = inp_dim // head.size(2)              # src/util.py:194
    grid = inp_dim // stride                      # src/util.py:195
    attrs = 5 + num_class
    n_anchor = len(anchors)

    # src/util.py:199-203 -- NCHW -> [B, cells*anchors, attrs]; the copy made here is the
    # only tensor the function writes to, so the caller's tensor is never mutated.
    out = head.view(batch, attrs * n_anchor, grid * grid).transpose(1, 2).contiguous()
    out = out.view(batch, grid * grid * n_anchor, attrs)

    # src/util.py:206-208
    out[:, :, 0] = torch.sigmoid(out[:, :, 0])
    out[:, :, 1] = torch.sigmoid(out[:, :, 1])
    out[:, :, 4:] = torch.sigmoid(out[:, :, 4:])
    if TRAIN:
        return out

    # src/util.py:213-216 -- python-double division, then fp32
    scaled = torch.tensor([(a[0] / stride, a[1] / stride) for a in anchors],
                          dtype=torch.float32)

    # src/util.py:219-233 -- integer cell offsets; x is the column index
    ticks = torch.arange(grid)
    ys, xs = torch.meshgrid(ticks, ticks, indexing="ij")
    offs = torch.stack((xs.reshape(-1), ys.reshape(-1)), 1)          # [G*G, 2] int64
    offs = offs.repeat(1, n_anchor).view(-1, 2).unsqueeze(0)         # [1, G*G*A, 2]
    out[:, :, :2] += offs

    # src/util.py:235-237
    out[:, :, 2:4] = torch.exp(out[:, :, 2:4]) * scaled.repeat(grid * grid, 1).unsqueeze(0)
    out[:, :, :4] *= stride
    return out


def _greedy_suppress(rows: torch.Tensor, nms_conf: float) -> torch.Tensor:
    """src/util.py:314-329 -- rows [n,7] already sorted by objectness (descending).

    Box i, if still present, removes every later box whose IoU with it is NOT < nms_conf
    (so a NaN IoU suppresses).  Removed rows are multiplied by zero and then dropped by
    their (now zero) objectness, exactly like the reference.
    """
    total = rows.size(0)
    for i in range(total):
        if i >= rows.size(0):                       # reference: IndexError -> break
            break
        ious = bbox_iou(rows[i].unsqueeze(0), rows[i + 1:])
        keep = (ious < nms_conf).to(torch.float32).unsqueeze(1)
        rows[i + 1:] *= keep
        alive = torch.nonzero(rows[:, 4]).squeeze()
        rows = rows[alive].view(-1, 7)
    return rows


def write_results(pred: torch.Tensor, num_class: int, confidence: float = 0.6,
                  nms_conf: float = 0.4):
    """src/util.py:242-346 -- threshold + per-image per-class greedy NMS.

    Returns [D, 8] fp32 rows [img, x1, y1, x2, y2, obj, cls_conf, cls] ordered image
    ascending, class ascending, objectness descending -- or the int 0 when nothing
    survives.  The sort is torch.sort(descending=True), which is NOT stable: rows with
    bit-equal objectness inside one (image, class) have implementation-defined order.
    """
    pred = confidence_mask(pred, confidence)                          # :260

    # :263-268 -- centre/size -> corners (w/2 is exact, the subtraction rounds once)
    corners = torch.empty_like(pred[:, :, :4])
    corners[:, :, 0] = pred[:, :, 0] - pred[:, :, 2] / 2
    corners[:, :, 1] = pred[:, :, 1] - pred[:, :, 3] / 2
    corners[:, :, 2] = pred[:, :, 0] + pred[:, :, 2] / 2
    corners[:, :, 3] = pred[:, :, 1] + pred[:, :, 3] / 2
    pred[:, :, :4] = corners

    chunks = []
    for img in range(pred.size(0)):                                   # :275
        rows = pred[img]
        best, best_idx = torch.max(rows[:, 5:5 + num_class], 1)       # :279 first max wins
        rows = torch.cat((rows[:, :5], best.float().unsqueeze(1),
                          best_idx.float().unsqueeze(1)), 1)          # [N,7]
        hit = torch.nonzero(rows[:, 4]).squeeze(1)                    # :286 obj != 0
        if hit.numel() == 0:
            continue
        rows = rows[hit].view(-1, 7)

        for cls in torch.unique(rows[:, -1]):                         # :298 ascending
            same = rows * (rows[:, -1] == cls).to(torch.float32).unsqueeze(1)
            idx = torch.nonzero(same[:, -2]).squeeze(1)               # :305 cls_conf != 0
            seg = rows[idx].view(-1, 7)
            order = torch.sort(seg[:, 4], descending=True)[1]         # :309
            seg = _greedy_suppress(seg[order], nms_conf)
            tag = seg.new_full((seg.size(0), 1), float(img))          # :332
            chunks.append(torch.cat((tag, seg), 1))

    if not chunks:
        return 0                                                      # :343-346
    return torch.cat(chunks, 0)
