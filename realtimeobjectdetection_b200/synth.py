"""Synthetic Darknet weights (there is no network: get_weights.sh cannot run).

Two deterministic (numpy ``RandomState``) generators produce an fp32 parameter stream in
the order ``Darknet.load_weights`` consumes it (src/darknet.py:316-410: per conv block
``[bn.bias, bn.weight, bn.running_mean, bn.running_var]`` or ``[conv.bias]``, then the
conv weight ``[Cout, Cin, k, k]``):

* ``mode="default"`` mimics PyTorch's default initialisation of the reference modules
  (conv weight/bias ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)); BatchNorm identity).  In eval
  mode the signal decays layer by layer, so the head logits are ~ the head bias -- the
  degenerate network SURVEY.md "fact 4" describes.  This is the north-star contract case.
* ``mode="calibrated"`` tracks per-channel mean/variance through the network analytically
  and chooses BatchNorm statistics that keep activations O(1) with non-trivial
  gamma/beta/mean/var, and head biases such that about ``obj_pass`` of the rows clear an
  objectness of 0.5 -- a network that behaves like a trained detector for NMS load.

Nothing here runs the network; it is data generation only (numpy, host side).
"""
from __future__ import annotations

import math

import numpy as np

_SQRT2 = math.sqrt(2.0)


def layer_table(blocks):
    """Shape-infer the cfg like ``create_modules`` (src/darknet.py:449-533).

    Returns one dict per non-[net] block: type, index, in/out channels and, for convs,
    size/stride/pad/bn/leaky; for routes the absolute source indices.
    """
    table = []
    out_ch = []
    prev = 3                                   # the reference hard-wires 3 input channels (:601)
    filters = prev
    for i, blk in enumerate(blocks[1:]):
        kind = blk["type"]
        row = {"type": kind, "index": i, "cin": prev}
        if kind == "convolutional":
            try:
                bn = int(blk["batch_normalize"])
                has_bias = False
            except (KeyError, ValueError):
                bn, has_bias = 0, True
            filters = int(blk["filters"])
            size = int(blk["size"])
            row.update(cout=filters, size=size, stride=int(blk["stride"]),
                       pad=(size - 1) // 2 if int(blk["pad"]) else 0, bn=bool(bn),
                       bias=has_bias, leaky=blk["activation"] == "leaky")
        elif kind == "route":
            refs = blk["layers"]
            refs = [int(v) for v in (refs.split(",") if isinstance(refs, str) else refs)]
            src = [r if r > 0 else i + r for r in refs]
            row["sources"] = src
            filters = sum(out_ch[s] for s in src)
        elif kind == "shortcut":
            row["sources"] = [i - 1, i + int(blk["from"])]
        elif kind == "maxpool":
            row.update(size=int(blk["size"]), stride=int(blk["stride"]))
        elif kind == "yolo":
            mask = [int(v) for v in blk["mask"].split(",")]
            flat = [int(v) for v in blk["anchors"].split(",")]
            pairs = [(flat[k], flat[k + 1]) for k in range(0, len(flat), 2)]
            row.update(anchors=[pairs[m] for m in mask], classes=int(blk["classes"]))
        elif kind != "upsample":
            raise AssertionError("unknown block " + kind)
        row["cout"] = filters
        table.append(row)
        out_ch.append(filters)
        prev = filters
    return table


def _phi(t):
    return np.exp(-0.5 * t * t) / math.sqrt(2.0 * math.pi)


def _Phi(t):
    return 0.5 * (1.0 + np.vectorize(math.erf)(t / _SQRT2))


def _leaky_moments(a, s, slope=0.1):
    """Mean and variance of leaky_relu(z), z ~ N(a, s^2), element-wise."""
    s = np.maximum(s, 1e-12)
    t = a / s
    P, Q, d = _Phi(t), _Phi(-t), _phi(t)
    mean = a * (P + slope * Q) + s * d * (1.0 - slope)
    second = (a * a + s * s) * (P + slope * slope * Q) + a * s * d * (1.0 - slope * slope)
    return mean, np.maximum(second - mean * mean, 1e-12)


def synth_stream(blocks, seed: int = 0, mode: str = "calibrated", obj_pass: float = 0.01,
                 logit_std: float = 0.6) -> np.ndarray:
    """fp32 parameter stream for ``load_weights`` (see module docstring)."""
    assert mode in ("calibrated", "default")
    rng = np.random.RandomState(seed)
    table = layer_table(blocks)
    chunks = []
    # per-layer per-channel (mean, variance) of the activations, for mode="calibrated"
    stats = []
    in_m, in_v = np.full(3, 0.5), np.full(3, 1.0 / 12.0)         # input ~ U(0,1)
    pre_act = None                # (mean, std) before the leaky of the latest conv block
    for row in table:
        kind = row["type"]
        if kind == "convolutional":
            cin, cout, k = row["cin"], row["cout"], row["size"]
            fan_in = cin * k * k
            if mode == "default":
                bound = 1.0 / math.sqrt(fan_in)
                w = rng.uniform(-bound, bound, size=(cout, cin, k, k)).astype(np.float32)
                if row["bn"]:
                    chunks += [np.zeros(cout, np.float32), np.ones(cout, np.float32),
                               np.zeros(cout, np.float32), np.ones(cout, np.float32)]
                else:
                    chunks.append(rng.uniform(-bound, bound, size=cout).astype(np.float32))
                chunks.append(w.reshape(-1))
                stats.append((np.zeros(cout), np.ones(cout)))
                in_m, in_v = stats[-1]
                continue
            w = (rng.standard_normal((cout, cin, k, k)) / math.sqrt(fan_in)).astype(np.float32)
            w64 = w.astype(np.float64)
            z_mean = np.einsum("ocij,c->o", w64, in_m)                  # spatial mean of conv out
            z_var = np.einsum("ocij,c->o", w64 * w64, in_v)             # spatial variance
            z_std = np.sqrt(np.maximum(z_var, 1e-20))
            if row["bn"]:
                gamma = rng.uniform(0.8, 1.25, cout)
                beta = 0.15 * rng.standard_normal(cout)
                rho = rng.uniform(0.75, 1.35, cout)                     # var mis-estimate
                delta = 0.2 * rng.standard_normal(cout)                 # mean mis-estimate
                run_mean = z_mean + delta * z_std
                run_var = z_var * rho
                a = beta - gamma * delta / np.sqrt(rho)                 # pre-activation mean
                s = gamma / np.sqrt(rho)                                # pre-activation std
                chunks += [beta.astype(np.float32), gamma.astype(np.float32),
                           run_mean.astype(np.float32), run_var.astype(np.float32)]
                if row["leaky"]:
                    m, v = _leaky_moments(a, s)
                    pre_act = (a, s)
                else:
                    m, v = a, s * s
            else:
                # detection head: rescale to the wanted logit spread, cancel the DC term,
                # bias objectness so ~obj_pass of the rows pass 0.5
                scale = logit_std / z_std
                w *= scale[:, None, None, None].astype(np.float32)
                bias = -z_mean * scale
                attrs = cout // 3 if cout % 3 == 0 else cout
                zq = _normal_quantile(1.0 - obj_pass)
                for ch in range(cout):
                    if ch % attrs == 4:
                        bias[ch] -= zq * logit_std
                chunks.append(bias.astype(np.float32))
                m, v = np.zeros(cout), np.full(cout, logit_std ** 2)
            chunks.append(w.reshape(-1))
            stats.append((m, v))
        elif kind == "route":
            src = row["sources"]
            stats.append((np.concatenate([stats[s][0] for s in src]),
                          np.concatenate([stats[s][1] for s in src])))
        elif kind == "shortcut":
            a, b = row["sources"]
            stats.append((stats[a][0] + stats[b][0], stats[a][1] + stats[b][1]))
        elif kind == "upsample":
            stats.append((in_m, in_v * 0.7))                           # bilinear smoothing
        elif kind == "maxpool":
            stats.append(_maxpool_moments(pre_act, row["size"] ** 2) if pre_act is not None
                         else (in_m + 0.85 * np.sqrt(in_v), in_v * 0.6))
        else:                                                          # yolo
            stats.append((in_m, in_v))
        in_m, in_v = stats[-1]
    return np.concatenate(chunks).astype(np.float32, copy=False)


def _maxpool_moments(pre_act, window: int):
    """Mean/variance of max over ``window`` iid leaky(N(a, s^2)) samples, per channel.

    leaky is monotone, so the max commutes with it: integrate leaky(t) against the density
    of the maximum of ``window`` normals (window * phi * Phi^(window-1)) on a grid.
    """
    a, s = pre_act
    u = np.linspace(-8.0, 8.0, 3201)
    du = u[1] - u[0]
    dens = window * _phi(u) * _Phi(u) ** (window - 1) * du               # standardised max
    t = a[:, None] + s[:, None] * u[None, :]
    y = np.where(t > 0, t, 0.1 * t)
    mean = (y * dens).sum(1)
    second = (y * y * dens).sum(1)
    return mean, np.maximum(second - mean * mean, 1e-12)


def _normal_quantile(p: float) -> float:
    lo, hi = -10.0, 10.0
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if 0.5 * (1.0 + math.erf(mid / _SQRT2)) < p:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def write_weights_file(path: str, stream: np.ndarray, seen: int = 0) -> None:
    """Darknet binary layout: int32[5] header (major, minor, revision, seen lo, seen hi)
    followed by the fp32 stream (src/darknet.py:397-410)."""
    header = np.array([0, 2, 0, seen & 0x7FFFFFFF, 0], dtype=np.int32)
    with open(path, "wb") as fh:
        fh.write(header.tobytes())
        fh.write(np.ascontiguousarray(stream, dtype=np.float32).tobytes())


def stream_to_state(blocks, stream: np.ndarray) -> dict:
    """Split a parameter stream into numpy arrays keyed like the reference's state_dict
    (``module_list.{i}.conv_{i}.weight`` ...), following ``load_weights`` order."""
    state = {}
    pos = 0

    def take(n, shape):
        nonlocal pos
        arr = np.asarray(stream[pos:pos + n], dtype=np.float32).reshape(shape)
        pos += n
        return arr

    for row in layer_table(blocks):
        if row["type"] != "convolutional":
            continue
        i, cout, cin, k = row["index"], row["cout"], row["cin"], row["size"]
        if row["bn"]:
            pre = "module_list.%d.batch_norm_%d." % (i, i)
            state[pre + "bias"] = take(cout, (cout,))
            state[pre + "weight"] = take(cout, (cout,))
            state[pre + "running_mean"] = take(cout, (cout,))
            state[pre + "running_var"] = take(cout, (cout,))
        else:
            state["module_list.%d.conv_%d.bias" % (i, i)] = take(cout, (cout,))
        state["module_list.%d.conv_%d.weight" % (i, i)] = take(cout * cin * k * k,
                                                               (cout, cin, k, k))
    if pos != len(stream):
        raise ValueError("parameter stream has %d values, network consumes %d" % (len(stream), pos))
    return state
