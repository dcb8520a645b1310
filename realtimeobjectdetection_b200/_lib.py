"""ctypes binding of ``librtod.so`` (C ABI declared in include/rtod.h).

There is no CPU fallback: importing a symbol from a missing library raises.  The library is
built in-tree by ``make -C realtimeobjectdetection_b200/csrc`` (see ``build_library``), so the
``.so`` sits next to this file and travels with the repository snapshot.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTOD_LIB") or os.path.join(_HERE, "librtod.so")    # RTOD_LIB: bring-up builds (make TRACE=1)
CSRC_DIR = os.path.join(_HERE, "csrc")
ABI_VERSION = 2

RTOD_MAX_ANCHORS = 8
LAYER_CONV, LAYER_SHORTCUT, LAYER_ROUTE, LAYER_UPSAMPLE, LAYER_MAXPOOL, LAYER_YOLO = range(6)
PLAN_KEEP_ALL = 1
PLAN_CONV_SIMT = 2
PLAN_NO_AUTOTUNE = 4
PLAN_BF16 = 8            # bf16 storage (default: fp16)
PLAN_NO_WSPLIT = 16      # fp16: no two-term weights in the HBM-bound early layers


class RtodError(RuntimeError):
    """A librtod call returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__("librtod error %d: %s" % (code, message))
        self.code = code


class RtodLayerDesc(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int32), ("filters", ctypes.c_int32), ("size", ctypes.c_int32),
                ("stride", ctypes.c_int32), ("pad", ctypes.c_int32),
                ("batch_normalize", ctypes.c_int32), ("leaky", ctypes.c_int32),
                ("src0", ctypes.c_int32), ("src1", ctypes.c_int32),
                ("num_anchors", ctypes.c_int32), ("classes", ctypes.c_int32),
                ("anchors", ctypes.c_float * (2 * RTOD_MAX_ANCHORS))]


_vp, _i, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
_PROTOTYPES = {
    # name: (restype, argtypes)
    "rtod_abi_version": (_i, []),
    "rtod_last_error": (ctypes.c_char_p, []),
    "rtod_plan_create": (_i, [ctypes.POINTER(RtodLayerDesc), _i, _i, _i, _i, _i, _i, ctypes.c_uint,
                              ctypes.POINTER(_vp)]),
    "rtod_plan_destroy": (None, [_vp]),
    "rtod_plan_workspace_bytes": (_sz, [_vp]),
    "rtod_plan_scratch_bytes": (_sz, [_vp]),
    "rtod_plan_weight_bytes": (_sz, [_vp]),
    "rtod_plan_num_rows": (_i, [_vp]),
    "rtod_plan_num_attrs": (_i, [_vp]),
    "rtod_plan_layer_shape": (_i, [_vp, _i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "rtod_plan_launch_count": (_i, [_vp]),
    "rtod_plan_conv_flops": (ctypes.c_double, [_vp]),
    "rtod_plan_bind": (_i, [_vp, _vp, _sz, _vp, _sz]),
    "rtod_plan_set_conv_weights": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp]),
    "rtod_plan_forward": (_i, [_vp, _vp, _vp, _i, _vp]),
    "rtod_plan_forward_u8": (_i, [_vp, _vp, _vp, _i, _vp]),
    "rtod_plan_forward_profile": (_i, [_vp, _vp, _vp, _i, _vp, ctypes.POINTER(_f), ctypes.POINTER(_i)]),
    "rtod_plan_forward_segments": (_i, [_vp, _vp, _vp, _i, _vp, ctypes.POINTER(_f), ctypes.POINTER(_f)]),
    "rtod_plan_layer_flops": (ctypes.c_double, [_vp, _i]),
    "rtod_plan_conv_backend": (_i, [_vp, _i]),
    "rtod_plan_conv_config": (_i, [_vp, _i, ctypes.POINTER(_i)]),
    "rtod_plan_read_layer": (_i, [_vp, _i, _vp, _vp]),
    "rtod_plan_check": (_i, [_vp, _vp]),
    "rtod_plan_set_error_sink": (_i, [_vp, _vp, _vp]),
    "rtod_plan_reset_errors": (_i, [_vp, _vp]),
    "rtod_plan_is_f16": (_i, [_vp]),
    "rtod_plan_conv_w_split": (_i, [_vp, _i]),
    "rtod_plan_conv_row_mode": (_i, [_vp, _i]),
    "rtod_yolo_decode": (_i, [_vp, _i, _i, _i, _i, _i, ctypes.POINTER(_f), _i, _vp, _vp]),
    "rtod_write_results_workspace_bytes": (_sz, [_i, _i, _i]),
    "rtod_write_results": (_i, [_vp, _i, _i, _i, _f, _f, _vp, _i, _vp, _vp, _sz, _vp]),
    "rtod_confidence_mask": (_i, [_vp, ctypes.c_longlong, _i, _f, _vp, _vp]),
    "rtod_bbox_iou": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp]),
    "rtod_sm_clock_probe": (_i, [_vp, _i, _i, _vp]),
    "rtod_letterbox_geometry": (_i, [_i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i),
                                     ctypes.POINTER(_i)]),
    "rtod_prep_image": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "rtod_rescale_boxes": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "rtod_pack_detections": (_i, [_vp, _i, _vp, _f, _i, _vp, _vp]),
    "rtod_bbox_iou_matrix": (_i, [_vp, _i, _i, _vp, _i, _i, _i, ctypes.c_double, _vp, _vp]),
}
EXPORTED_SYMBOLS = tuple(sorted(_PROTOTYPES))

_lock = threading.Lock()
_lib = None


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into ``librtod.so`` (in-tree)."""
    cmd = ["make", "-C", CSRC_DIR, "-j8"]
    if force:
        subprocess.run(["make", "-C", CSRC_DIR, "clean"], check=True, capture_output=not verbose)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building librtod.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Load (once) and type the library; raises if it is missing or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "librtod.so not found at %s -- build it with `make -C %s` or "
                "`python -c 'import __graft_entry__ as g; g.build()'`; there is no CPU fallback"
                % (LIB_PATH, CSRC_DIR))
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _PROTOTYPES.items():
            fn = getattr(lib, name)            # AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.rtod_abi_version() != ABI_VERSION:
            raise RuntimeError("librtod.so ABI %d != expected %d; rebuild" %
                               (lib.rtod_abi_version(), ABI_VERSION))
        _lib = lib
    return _lib


def check(code: int) -> None:
    if code != 0:
        raise RtodError(code, load().rtod_last_error().decode("utf-8", "replace"))
