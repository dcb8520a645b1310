"""Oracle (test infrastructure): CPU restatement of the reference's Darknet forward pass.

``DarknetPort`` consumes a Darknet ``.cfg`` and a parameter dictionary that uses the
reference's own ``state_dict`` key names (src/darknet.py:488-501: ``module_list.{i}.conv_{i}.weight``,
``module_list.{i}.batch_norm_{i}.{weight,bias,running_mean,running_var}``,
``module_list.{i}.conv_{i}.bias``) and evaluates the network with functional fp32
PyTorch ops in the reference's order (src/darknet.py:199-253).  BatchNorm uses the
running statistics (``bn_mode="eval"``, the parity oracle) or the batch statistics
(``bn_mode="batch"``: what the reference's scripts actually execute because they never
call ``.eval()``; used only for the timed CPU baseline).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .detect_port import predict_transform


def parse_cfg(path: str):
    """src/darknet.py:412-447 -- cfg text -> list of {key: str} blocks ([net] first).

    Blank lines and lines starting with '#' are dropped *before* stripping (so an
    indented comment would not be treated as one, like the reference); values stay strings.
    """
    with open(path, "r") as fh:
        raw = fh.read().split("\n")
    lines = [ln for ln in raw if len(ln) > 0]
    lines = [ln for ln in lines if ln[0] != "#"]
    lines = [ln.strip() for ln in lines]

    blocks, cur = [], {}
    for ln in lines:
        if ln[0] == "[":
            if cur:
                blocks.append(cur)
                cur = {}
            cur["type"] = ln[1:-1].rstrip()
        else:
            key, value = ln.split("=")
            cur[key.rstrip()] = value.lstrip()
    blocks.append(cur)
    return blocks


class DarknetPort:
    """Functional re-evaluation of ``Darknet.forward`` (src/darknet.py:199-303)."""

    def __init__(self, cfg_path: str, params: dict, bn_mode: str = "eval"):
        assert bn_mode in ("eval", "batch")
        self.blocks = parse_cfg(cfg_path)
        self.net_info = self.blocks[0]
        self.params = params
        self.bn_mode = bn_mode
        self.anchors = None
        self.num_classes = None

    # -- per block evaluators ------------------------------------------------------
    def _conv(self, i: int, blk: dict, x: torch.Tensor) -> torch.Tensor:
        """src/darknet.py:467-501 -- conv [+ BN] [+ leaky 0.1]."""
        p = self.params
        ksize = int(blk["size"])
        stride = int(blk["stride"])
        pad = (ksize - 1) // 2 if int(blk["pad"]) else 0                 # :482-485
        has_bn = False
        try:                                                           # :470-475
            has_bn = bool(int(blk["batch_normalize"]))
            bias = None
        except (KeyError, ValueError):
            bias = p[f"module_list.{i}.conv_{i}.bias"]
        y = F.conv2d(x, p[f"module_list.{i}.conv_{i}.weight"], bias, stride, pad)
        if has_bn:
            pre = f"module_list.{i}.batch_norm_{i}."
            if self.bn_mode == "eval":
                y = F.batch_norm(y, p[pre + "running_mean"], p[pre + "running_var"],
                                 p[pre + "weight"], p[pre + "bias"], False, 0.1, 1e-5)
            else:
                y = F.batch_norm(y, None, None, p[pre + "weight"], p[pre + "bias"],
                                 True, 0.1, 1e-5)
        if blk["activation"] == "leaky":                                # :499-501
            y = F.leaky_relu(y, 0.1)
        return y

    @staticmethod
    def _maxpool(blk: dict, x: torch.Tensor) -> torch.Tensor:
        """src/darknet.py:547-555 and MaxPoolStride1 :37-46."""
        size, stride = int(blk["size"]), int(blk["stride"])
        if stride != 1:
            return F.max_pool2d(x, size, stride)
        x = F.pad(x, (0, size - 1, 0, size - 1), mode="replicate")
        return F.max_pool2d(x, size, size - 1)

    @staticmethod
    def _route_sources(i: int, blk: dict):
        """src/darknet.py:270-283 -- absolute indices of the routed layers."""
        refs = [int(v) for v in blk["layers"].split(",")] \
            if isinstance(blk["layers"], str) else [int(v) for v in blk["layers"]]
        return [i + r if r <= 0 else r for r in refs]

    # -- forward -----------------------------------------------------------------------
    def forward(self, x: torch.Tensor, TRAIN: bool = False):
        outputs = {}
        heads = []
        self.anchors = None
        self.layer_outputs = outputs            # kept for layer-wise parity checks
        for i, blk in enumerate(self.blocks[1:]):
            kind = blk["type"]
            if kind == "convolutional":
                x = self._conv(i, blk, x)
            elif kind == "upsample":                                   # :591-592
                x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
            elif kind == "maxpool":
                x = self._maxpool(blk, x)
            elif kind == "route":                                      # :285-288
                src = self._route_sources(i, blk)
                x = outputs[src[0]] if len(src) == 1 else \
                    torch.cat((outputs[src[0]], outputs[src[1]]), 1)
            elif kind == "shortcut":                                   # :264-268
                x = outputs[i - 1] + outputs[i + int(blk["from"])]
            elif kind == "yolo":                                       # :226-247
                mask = [int(v) for v in blk["mask"].split(",")]
                flat = [int(v) for v in blk["anchors"].split(",")]
                pairs = [(flat[k], flat[k + 1]) for k in range(0, len(flat), 2)]
                anchors = [pairs[m] for m in mask]
                self.num_classes = int(blk["classes"])
                heads.append(predict_transform(x, int(self.net_info["height"]), anchors,
                                               self.num_classes, False, TRAIN=TRAIN))
                self.anchors = anchors.copy() if self.anchors is None \
                    else self.anchors + anchors
                outputs[i] = outputs[i - 1]
                x = heads[-1]
                continue
            else:
                raise AssertionError("unknown block " + kind)           # :525-526
            outputs[i] = x
        if not heads:
            return []
        return torch.cat(heads, 1)

    __call__ = forward
