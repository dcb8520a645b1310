"""Darknet ``.cfg`` handling for the detection hot path (host side, stays Python).

``parse_cfg`` is behaviour-identical to the reference parser
(src/darknet.py:412-447): blank lines and lines whose *first* character is '#' are
dropped, the rest is stripped, ``[type]`` opens a block, ``key=value`` pairs are kept as
strings.  ``builtin_cfg`` emits the two network descriptions the reference ships
(cfg/yolov3.cfg, cfg/yolov3-tiny.cfg) from a compact specification, so that the GPU
box -- where /root/reference does not exist -- can build the same networks; the
generated text parses to exactly the reference's blocks (checked by
tests/golden/make_golden.py and pinned by tests/golden/cfg_blocks.json).
"""
from __future__ import annotations

import os
import tempfile

_NET = """[net]
# Testing
batch=1
subdivisions=1
width=416
height=416
channels=3
momentum=0.9
decay=0.0005
angle=0
saturation = 1.5
exposure = 1.5
hue=.1

learning_rate=0.001
burn_in=1000
max_batches = 500200
policy=steps
steps=400000,450000
scales=.1,.1
"""

_V3_ANCHORS = "10,13,  16,30,  33,23,  30,61,  62,45,  59,119,  116,90,  156,198,  373,326"
_TINY_ANCHORS = "10,14,  23,27,  37,58,  81,82,  135,169,  344,319"


def parse_cfg(path: str):
    """cfg file -> list of dict blocks, ``[net]`` first (src/darknet.py:412-447)."""
    with open(path, "r") as fh:
        rows = fh.read().split("\n")
    rows = [r for r in rows if len(r) > 0]
    rows = [r for r in rows if r[0] != "#"]
    rows = [r.strip() for r in rows]
    blocks = []
    current = {}
    for row in rows:
        if row[0] == "[":
            if len(current) != 0:
                blocks.append(current)
                current = {}
            current["type"] = row[1:-1].rstrip()
        else:
            key, value = row.split("=")
            current[key.rstrip()] = value.lstrip()
    blocks.append(current)
    return blocks


def _conv(filters, size, stride=1, bn=True, act="leaky"):
    if bn:
        return ("[convolutional]\nbatch_normalize=1\nfilters=%d\nsize=%d\nstride=%d\npad=1\n"
                "activation=%s\n" % (filters, size, stride, act))
    return ("[convolutional]\nsize=%d\nstride=%d\npad=1\nfilters=%d\nactivation=%s\n"
            % (size, stride, filters, act))


def _yolo(mask, anchors, num, ignore):
    return ("[yolo]\nmask = %s\nanchors = %s\nclasses=80\nnum=%d\njitter=.3\n"
            "ignore_thresh = %s\ntruth_thresh = 1\nrandom=1\n" % (mask, anchors, num, ignore))


_SHORTCUT = "[shortcut]\nfrom=-3\nactivation=linear\n"
_UPSAMPLE = "[upsample]\nstride=2\n"


def _route(spec):
    return "[route]\nlayers = %s\n" % spec


def yolov3_cfg_text() -> str:
    """The 107-layer YOLOv3 description (75 conv, 23 shortcut, 4 route, 2 upsample, 3 yolo)."""
    out = [_NET, _conv(32, 3)]
    for filters, repeats in ((64, 1), (128, 2), (256, 8), (512, 8), (1024, 4)):
        out.append(_conv(filters, 3, 2))                       # stride-2 downsample
        for _ in range(repeats):                               # residual block
            out += [_conv(filters // 2, 1), _conv(filters, 3), _SHORTCUT]
    for k, (filters, mask) in enumerate(((512, "6,7,8"), (256, "3,4,5"), (128, "0,1,2"))):
        for _ in range(3):
            out += [_conv(filters, 1), _conv(filters * 2, 3)]
        out += [_conv(255, 1, bn=False, act="linear"), _yolo(mask, _V3_ANCHORS, 9, ".5")]
        if k < 2:
            out += [_route("-4"), _conv(filters // 2, 1), _UPSAMPLE,
                    _route("-1, 61" if k == 0 else "-1, 36")]
    return "\n".join(out)


def yolov3_tiny_cfg_text() -> str:
    """The 24-layer YOLOv3-tiny description (13 conv, 6 maxpool, 2 route, 1 upsample, 2 yolo)."""
    out = [_NET]
    for k, filters in enumerate((16, 32, 64, 128, 256, 512)):
        out += [_conv(filters, 3), "[maxpool]\nsize=2\nstride=%d\n" % (1 if k == 5 else 2)]
    out += [_conv(1024, 3), _conv(256, 1), _conv(512, 3),
            _conv(255, 1, bn=False, act="linear"), _yolo("3,4,5", _TINY_ANCHORS, 6, ".7"),
            _route("-4"), _conv(128, 1), _UPSAMPLE, _route("-1, 8"), _conv(256, 3),
            _conv(255, 1, bn=False, act="linear"), _yolo("0,1,2", _TINY_ANCHORS, 6, ".7")]
    return "\n".join(out)


_BUILTIN = {"yolov3": yolov3_cfg_text, "yolov3-tiny": yolov3_tiny_cfg_text}


def builtin_cfg(name: str, directory: str | None = None) -> str:
    """Write the named built-in network description to ``directory`` and return its path."""
    if name not in _BUILTIN:
        raise KeyError("unknown built-in cfg %r (have %s)" % (name, sorted(_BUILTIN)))
    directory = directory or os.path.join(tempfile.gettempdir(), "rtod_b200_cfg")
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, name + ".cfg")
    text = _BUILTIN[name]()
    if not os.path.exists(path) or open(path).read() != text:
        tmp = path + ".%d.tmp" % os.getpid()
        with open(tmp, "w") as fh:
            fh.write(text)
        os.replace(tmp, path)
    return path
