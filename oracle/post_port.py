"""Oracle (test infrastructure): CPU restatement of the reference's detection post-processing
(detect.py:104-136, 155) and of the validator's IoU matrix (test.py:139-151), with the same PyTorch
fp32 ops in the same order."""
from __future__ import annotations

import torch

from .detect_port import bbox_iou


def rescale_boxes(output: torch.Tensor, im_dim_list: torch.Tensor, inp_dim: int, ref_dim: int = 416):
    """detect.py:120-136 -- letterbox coordinates -> source-image coordinates, then clamp.

    ``output``: ``[D, 8]`` rows of ``write_results`` whose column 0 indexes ``im_dim_list``;
    ``im_dim_list``: ``[n_img, 4]`` fp32 rows ``(w, h, w, h)`` (detect.py:247-248).  The reference divides a
    hard-coded 416 by the image size (detect.py:130) while centring with ``inp_dim`` (:131-134):
    ``ref_dim`` reproduces that (pass ``ref_dim=inp_dim`` for the evident intent).  Returns
    ``(output, im_dim_rows)`` like the reference's in-place update + returned index_select.
    """
    output = output.clone()
    dims = torch.index_select(im_dim_list, 0, output[:, 0].long())           # :129
    scaling_factor = torch.min(ref_dim / dims, 1)[0].view(-1, 1)               # :130
    output[:, [1, 3]] -= (inp_dim - scaling_factor * dims[:, 0].view(-1, 1)) / 2   # :131-132
    output[:, [2, 4]] -= (inp_dim - scaling_factor * dims[:, 1].view(-1, 1)) / 2   # :133-134
    output[:, 1:5] /= scaling_factor                                            # :135
    for j in range(output.shape[0]):                                            # :121-126
        output[j, [1, 3]] = torch.clamp(output[j, [1, 3]], 0.0, dims[j, 0])
        output[j, [2, 4]] = torch.clamp(output[j, [2, 4]], 0.0, dims[j, 1])
    return output, dims


def metrics_rows(prediction):
    """detect.py:107, 164 -- what ``metrics.json`` stores per image name: the batch's ``write_results`` rows
    as nested lists (``prediction.tolist()``), or the int 0 when nothing was detected."""
    return prediction if isinstance(prediction, int) else prediction.tolist()


def iou_matrix(pred: torch.Tensor, target: torch.Tensor, threshold: float) -> torch.Tensor:
    """test.py:139-151 -- ``[P, T]`` matrix of ``bbox_iou(pred[i, 1:5], target[j, 0:4])``, entries not above
    ``threshold`` zeroed (strict ``>``, compared as Python floats of the fp32 IoU)."""
    box_ious, row = [], []
    for box in pred:
        for t_box in target:
            iou = bbox_iou(box[1:5].cpu(), t_box[0:4].cpu())
            row.append(iou.item() if iou.item() > threshold else 0.0)
        box_ious.append(row.copy())
        row.clear()
    return torch.FloatTensor(box_ious)
