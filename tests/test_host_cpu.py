"""CPU: host-side logic -- C-ABI exports, plan construction, weight files, module naming,
multi-rank gather (gloo, world_size 2).  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from helpers import make_network
from realtimeobjectdetection_b200 import Darknet, _lib, builtin_cfg, synth
from realtimeobjectdetection_b200.cfg import parse_cfg
from realtimeobjectdetection_b200.sharding import gather_detections, gather_detections_async, shard_bounds


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "rtod.h")).read()
    declared = sorted(set(re.findall(r"\b(rtod_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 20
    for name in declared:
        assert getattr(lib, name) is not None, name
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert lib.rtod_abi_version() == 2


def _create(lib, model, batch, hw, inp_dim, flags=0, in_c=3):
    descs = model._layer_descs()
    arr = (_lib.RtodLayerDesc * len(descs))(*descs)
    handle = ctypes.c_void_p()
    rc = lib.rtod_plan_create(arr, len(descs), batch, in_c, hw, hw, inp_dim, flags, ctypes.byref(handle))
    return rc, handle


@pytest.mark.parametrize("name,reso,rows,gflop,launches", [
    ("yolov3", 416, 10647, 65.864, 78), ("yolov3", 608, 22743, 140.692, 78),
    ("yolov3-tiny", 416, 2535, 5.565, 21), ("yolov3-tiny", 320, 1500, 3.293, 21)])
def test_plan_shapes_match_survey(lib, name, reso, rows, gflop, launches):
    model = Darknet(builtin_cfg(name), False)
    rc, h = _create(lib, model, 1, reso, reso)
    assert rc == 0, lib.rtod_last_error()
    assert lib.rtod_plan_num_rows(h) == rows
    assert lib.rtod_plan_num_attrs(h) == 85
    assert abs(lib.rtod_plan_conv_flops(h) / 1e9 - gflop) < 1e-3
    assert lib.rtod_plan_launch_count(h) == launches          # convs + upsample/maxpool + 1 decode
    c, hh, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.rtod_plan_layer_shape(h, 0, ctypes.byref(c), ctypes.byref(hh), ctypes.byref(w)) == 0
    assert (c.value, hh.value, w.value) == ((32 if name == "yolov3" else 16), reso, reso)
    # liveness planning: far less than keeping all outputs like the reference's outputs{} dict
    rc, h_all = _create(lib, model, 1, reso, reso, _lib.PLAN_KEEP_ALL)
    assert rc == 0
    act = lambda plan: lib.rtod_plan_workspace_bytes(plan) - lib.rtod_plan_scratch_bytes(plan)   # activations only
    assert act(h) < 0.5 * act(h_all)
    lib.rtod_plan_destroy(h)
    lib.rtod_plan_destroy(h_all)


def test_plan_error_paths(lib):
    model = Darknet(builtin_cfg("yolov3-tiny"), False)
    rc, _ = _create(lib, model, 1, 416, 300)                    # net_info height disagrees with grid
    assert rc == -1 and b"does not match" in lib.rtod_last_error()
    rc, _ = _create(lib, model, 0, 416, 416)
    assert rc == -1
    rc, h = _create(lib, model, 1, 416, 416)
    assert rc == 0
    assert lib.rtod_plan_forward(h, None, None, 0, None) == -5  # not bound
    assert lib.rtod_plan_layer_shape(h, 99, None, None, None) == -1
    lib.rtod_plan_destroy(h)
    assert lib.rtod_write_results_workspace_bytes(2, 10647, 80) > 2 * 16384 * 8
    assert lib.rtod_write_results(None, 1, 8, 3, 0.5, 0.4, None, 0, None, None, 0, None) == -1


def test_module_names_and_weight_file_roundtrip(tmp_path):
    cfg, blocks, stream, state = make_network("yolov3-tiny", 9, "calibrated")
    path = str(tmp_path / "tiny.weights")
    synth.write_weights_file(path, stream, seen=1234)
    assert os.path.getsize(path) == 20 + 4 * 8858734
    model = Darknet(cfg, False)
    model.load_weights(path)
    assert int(model.seen) == 1234 and model.header.tolist()[:3] == [0, 2, 0]
    sd = model.state_dict()
    for key, value in state.items():
        assert torch.equal(sd[key], value), key
    extra = [k for k in sd if k not in state and not k.endswith("num_batches_tracked")]
    assert extra == []
    # reference-visible structure
    assert len(model.module_list) == 24 and len(model.blocks) == 25
    assert model.blocks[21]["layers"] == ["-1", " 8"]             # split(',') like the reference
    assert model.module_list[16][0].anchors == [(81, 82), (135, 169), (344, 319)]
    assert isinstance(model.net_info, dict) and model.net_info["height"] == "416"
    with model.train_mode():
        assert model.TRAIN is True
    assert model.TRAIN is False


def test_yolov3_stream_size_matches_real_weights_file():
    blocks = parse_cfg(builtin_cfg("yolov3"))
    n = sum((r["cout"] * (4 if r["bn"] else 1) + r["cout"] * r["cin"] * r["size"] ** 2)
            for r in synth.layer_table(blocks) if r["type"] == "convolutional")
    assert n == 62001757                                         # yolov3.weights = 20 + 4*n bytes


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "librtod.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_forward_without_gpu_raises():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    model = Darknet(builtin_cfg("yolov3-tiny"), False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 3, 64, 64))


def test_shard_bounds_cover_batch():
    for total in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def rows_match(a, b):
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, int) or isinstance(b, int):
        return isinstance(a, int) and isinstance(b, int) and a == b
    return torch.equal(a, b)


def _gather_worker(rank, world, port, case, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = {0: torch.tensor([[0, 1, 2, 3, 4, .9, .8, 5.], [1, 5, 6, 7, 8, .7, .6, 2.]]),
                1: torch.tensor([[0, 9, 9, 9, 9, .5, .5, 1.]])}
        if case == "mixed":
            local = rows[rank]
        elif case == "rank1_empty":
            local = rows[0] if rank == 0 else 0
        else:
            local = 0
        out = gather_detections(local, first_frame=rank * 4)
        # the streaming variant: fixed capacity, count in-band, unsliced row buffer + device-style count tensor
        buf = torch.full((6, 8), 7.0)                          # stale content beyond the count must not leak
        n = 0 if isinstance(local, int) else local.size(0)
        if n:
            buf[:n] = local
        out2 = gather_detections_async(buf, torch.tensor([n], dtype=torch.int32), rank * 4, capacity=5).result()
        assert rows_match(out, out2)
        if rank == 0:
            try:
                gather_detections_async(buf, torch.tensor([n + 6], dtype=torch.int32), 0, capacity=5).result()
                overflow = False
            except RuntimeError:
                overflow = True
            assert overflow
        else:
            gather_detections_async(buf, torch.tensor([n + 6], dtype=torch.int32), 0, capacity=5).result()
        if rank == 0:
            ret.put(out if isinstance(out, int) else out.clone())
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["mixed", "rank1_empty", "all_empty"])
def test_gather_detections_world2_gloo(case):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + {"mixed": 0, "rank1_empty": 1, "all_empty": 2}[case]
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, case, ret)) for r in range(2)]
    for p in procs:
        p.start()
    out = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    if case == "all_empty":
        assert out == 0
    elif case == "rank1_empty":
        assert out.shape == (2, 8) and out[:, 0].tolist() == [0., 1.]
    else:
        assert out.shape == (3, 8)
        assert out[:, 0].tolist() == [0., 1., 4.]                  # rank 1's frame 0 is global frame 4
        assert out[2, 1:].tolist() == [9., 9., 9., 9., .5, .5, 1.]
