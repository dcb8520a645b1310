"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np
import torch

import oracle
from realtimeobjectdetection_b200 import synth
from realtimeobjectdetection_b200.cfg import builtin_cfg, parse_cfg


def synth_pred(seed, B, N, C, density, clustered, span=416.0):
    """[B,N,5+C] prediction tensor, exactly round(density*N) rows per image above 0.5, all
    objectness values distinct.  Mirrors tests/golden/make_golden.py."""
    rng = np.random.RandomState(seed)
    pred = np.zeros((B, N, 5 + C), np.float32)
    K = int(round(density * N))
    for b in range(B):
        pred[b, :, 0:2] = rng.uniform(0, span, size=(N, 2))
        pred[b, :, 2:4] = np.exp(rng.uniform(2, 5, size=(N, 2)))
        pred[b, :, 5:] = rng.uniform(0, 1, size=(N, C))
        hi = 0.5 + 0.4995 * (np.arange(K) + 0.5) / max(K, 1)
        lo = 0.4995 * (np.arange(N - K) + 0.5) / max(N - K, 1)
        perm = rng.permutation(N)
        pred[b, perm, 4] = np.concatenate([hi, lo]).astype(np.float32)
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


def make_network(cfg_name, weight_seed, mode):
    """(cfg path, blocks, parameter stream, oracle state dict) for a built-in network."""
    cfg = builtin_cfg(cfg_name)
    blocks = parse_cfg(cfg)
    stream = synth.synth_stream(blocks, weight_seed, mode)
    state = {k: torch.from_numpy(v) for k, v in synth.stream_to_state(blocks, stream).items()}
    return cfg, blocks, stream, state


def oracle_forward(cfg, state, x, reso):
    port = oracle.DarknetPort(cfg, state)
    port.net_info["height"] = reso
    with torch.no_grad():
        return port(x)


def rows_equal(a, b):
    """bit-exact comparison of write_results outputs (tensor or int 0)."""
    if isinstance(a, int) or isinstance(b, int):
        return isinstance(a, int) and isinstance(b, int) and a == b
    return a.shape == b.shape and torch.equal(a.cpu(), b.cpu())
