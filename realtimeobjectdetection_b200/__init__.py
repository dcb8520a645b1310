"""B200-native (sm_100a) YOLOv3 detection hot path with the reference's Python API.

    from realtimeobjectdetection_b200 import Darknet, write_results
    model = Darknet(cfg_path, CUDA=True); model.load_weights(path)
    pred = model(x)                                   # [B, N, 5+C]  (src/darknet.py:199-253)
    det = write_results(pred, 80, 0.5, 0.4)           # [D, 8] or 0  (src/util.py:242-346)

Everything numerical runs in ``librtod.so`` (hand-written CUDA, C ABI in include/rtod.h); this
package is the host-side mirror of the reference interface.  Importing it does not need a GPU;
calling it does, and there is no CPU fallback.
"""
from .cfg import builtin_cfg, parse_cfg  # noqa: F401
from .darknet import Darknet, DetectionLayer, EmptyLayer, MaxPoolStride1  # noqa: F401
from .util import (bbox_iou, bbox_iou_matrix, confidence_mask, letterbox_image, metrics_rows,  # noqa: F401
                   predict_transform, prep_frames, prep_image, rescale_boxes, write_results,
                   write_results_async)

__all__ = ["Darknet", "DetectionLayer", "EmptyLayer", "MaxPoolStride1", "bbox_iou",
           "bbox_iou_matrix", "confidence_mask", "letterbox_image", "metrics_rows", "predict_transform",
           "prep_frames", "prep_image", "rescale_boxes", "write_results", "write_results_async",
           "builtin_cfg", "parse_cfg"]
