"""Seeded synthetic inputs for the warp + cost-volume path.

Features follow the reference's pyramid statistics: N(0,1) through leaky-ReLU(0.1)
(tfoptflow/model_pwcnet.py:1099); flows are smooth fields at level resolution
(the network's `up_flow * scaler`, tfoptflow/model_pwcnet.py:1560-1561), with an
i.i.d. stress variant.  Seeds are 1969 + level (the reference's random_seed,
tfoptflow/dataset_base.py:39).  Generated with numpy so that CPU tests, the GPU
tests and bench.py see bit-identical inputs without needing a GPU.
"""
from __future__ import annotations

import numpy as np

PYR_CHANNELS = {1: 16, 2: 32, 3: 64, 4: 96, 5: 128, 6: 196}  # model_pwcnet.py:1084
FLOW_SCALERS = {l: 20.0 / 2 ** l for l in range(1, 7)}  # model_pwcnet.py:1560


def level_shapes(height, width, levels=(6, 5, 4, 3, 2)):
    return [(l, height >> l, width >> l, PYR_CHANNELS[l]) for l in levels]


def _smooth_field(rng, b, h, w, sigma):
    ch, cw = max(2, (h + 7) // 8 + 1), max(2, (w + 7) // 8 + 1)
    coarse = rng.standard_normal((b, ch, cw, 2)).astype(np.float32)
    ys = np.linspace(0, ch - 1, h, dtype=np.float32)
    xs = np.linspace(0, cw - 1, w, dtype=np.float32)
    y0 = np.clip(np.floor(ys).astype(np.int64), 0, ch - 2)
    x0 = np.clip(np.floor(xs).astype(np.int64), 0, cw - 2)
    ay = (ys - y0)[None, :, None, None]
    ax = (xs - x0)[None, None, :, None]
    top = coarse[:, y0][:, :, x0] * (1 - ax) + coarse[:, y0][:, :, x0 + 1] * ax
    bot = coarse[:, y0 + 1][:, :, x0] * (1 - ax) + coarse[:, y0 + 1][:, :, x0 + 1] * ax
    return np.ascontiguousarray(((top * (1 - ay) + bot * ay) * sigma).astype(np.float32))


def make_level(b, h, w, c, seed, dtype=np.float32, flow="smooth", sigma=2.0, with_grad=False, search_range=4):
    """Returns dict(c1, c2, flow[, g]) as numpy arrays of `dtype` (NHWC)."""
    rng = np.random.default_rng(seed)
    def feat():
        x = rng.standard_normal((b, h, w, c)).astype(np.float32)
        return np.maximum(0.1 * x, x)
    out = {"c1": feat().astype(dtype), "c2": feat().astype(dtype)}
    if flow == "smooth":
        f = _smooth_field(rng, b, h, w, sigma)
    elif flow == "stress":
        f = rng.uniform(-8.0, 8.0, (b, h, w, 2)).astype(np.float32)
    elif flow == "zero":
        f = np.zeros((b, h, w, 2), np.float32)
    elif flow is None or flow == "none":
        f = None
    else:
        raise ValueError(flow)
    out["flow"] = None if f is None else np.ascontiguousarray(f.astype(dtype))
    if with_grad:
        d = (2 * search_range + 1) ** 2
        out["g"] = rng.standard_normal((b, h, w, d)).astype(np.float32).astype(dtype)
    return out


def make_pyramid(b, height, width, dtype=np.float32, levels=(6, 5, 4, 3, 2), flow="smooth", top_has_flow=False,
                 with_grad=False, search_range=4):
    """Inputs for one batch through the pyramid loop (model_pwcnet.py:1550-1567):
    list of dict(lvl, c1, c2, flow|None, scaler).  The top level has no flow unless
    `top_has_flow` (operator sweep, BASELINE.json configs[1])."""
    out = []
    for (l, h, w, c) in level_shapes(height, width, levels):
        has_flow = top_has_flow or l != max(levels)
        d = make_level(b, h, w, c, 1969 + l, dtype, flow if has_flow else None,
                       with_grad=with_grad, search_range=search_range)
        # the op sees up_flow (pre-scaler); sigma is chosen post-scaler, so divide it out
        if d["flow"] is not None:
            d["flow"] = np.ascontiguousarray((d["flow"].astype(np.float32) / FLOW_SCALERS[l]).astype(dtype))
        d["lvl"] = l
        d["scaler"] = FLOW_SCALERS[l] if d["flow"] is not None else 1.0
        out.append(d)
    return out


def fwd_bytes_per_pair(height, width, elt, search_range=4, levels=(6, 5, 4, 3, 2), top_has_flow=False):
    """Algorithmic HBM bytes of one pair's forward: (2C + 2[has_flow] + D) * elt per pixel."""
    d = (2 * search_range + 1) ** 2
    tot = 0
    for (l, h, w, c) in level_shapes(height, width, levels):
        hf = 1 if (top_has_flow or l != max(levels)) else 0
        tot += h * w * (2 * c + 2 * hf + d) * elt
    return tot


def bwd_bytes_per_pair(height, width, elt, search_range=4, levels=(6, 5, 4, 3, 2), top_has_flow=False):
    """(2D + 4C + 4[has_flow]) * elt per pixel."""
    d = (2 * search_range + 1) ** 2
    tot = 0
    for (l, h, w, c) in level_shapes(height, width, levels):
        hf = 1 if (top_has_flow or l != max(levels)) else 0
        tot += h * w * (2 * d + 4 * c + 4 * hf) * elt
    return tot
