"""CPU oracle for PWC-Net's warp + cost-volume hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement, statement by statement, of the two reference
operators on the hot path (paths relative to /root/reference):

  * tfoptflow/core_warp.py:42-150   `_interpolate_bilinear`
  * tfoptflow/core_warp.py:153-202  `dense_image_warp`
  * tfoptflow/core_costvol.py:20-40 `cost_volume`

plus the analytic backward that TensorFlow 1.10 autodiff would generate for
them (the reference has no backward code of its own, it
is selected by `optim.minimize`, tfoptflow/model_pwcnet.py:547-550).

PARITY PINNED TO THE REFERENCE'S OWN SOURCE: TensorFlow 1.10.1 (tfoptflow/setup/dlubu36.yml:153) cannot be
imported or installed in this image (Python 3.12, no network) and the reference ships no tests for this path, but
its two hot-path files are plain Python over ~25 TF symbols.  tests/golden/tf_numpy_shim.py supplies those symbols
on numpy, imports tfoptflow/core_warp.py and core_costvol.py UNMODIFIED and executes them;
tests/test_oracle_vs_reference.py asserts np.array_equal(oracle, reference) on the level shapes of every
BASELINE.json config (fp32 + fp16, r = 4 and 8) and checks the backward below against float64 torch.autograd
through the same unmodified source; tests/golden/*.npz are generated from that import
(tests/golden/make_golden_from_reference.py).  What stays an assumption is kernel-level TF behaviour the Python
source does not show: fp16 reduce_mean accumulating in fp32, and the gradient routing of Maximum/Minimum at exact
ties (TF 1.x `_MaximumMinimumGrad`), which only matters at integer coordinates / zero pre-activations.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
may import this module.  The product path (tfoptflow_b200/) never does.

Every array op below is performed in the dtype the reference graph would use,
so that `float16` inputs reproduce the reference's all-fp16 regime
(tfoptflow/model_pwcnet.py:390-394) op by op.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "interpolate_bilinear",
    "dense_image_warp",
    "cost_volume",
    "warp_cost_volume",
    "cost_volume_bwd",
    "dense_image_warp_bwd",
    "warp_cost_volume_bwd",
    "level_shapes",
    "PYR_CHANNELS",
]

# tfoptflow/model_pwcnet.py:1084 - feature channels of pyramid levels 1..6
PYR_CHANNELS = {1: 16, 2: 32, 3: 64, 4: 96, 5: 128, 6: 196}


def level_shapes(height: int, width: int, levels=(6, 5, 4, 3, 2)):
    """(lvl, h, w, C) of every pyramid level for a padded HxW input
    (tfoptflow/model_pwcnet.py:1012-1101: each level halves the resolution)."""
    return [(l, height >> l, width >> l, PYR_CHANNELS[l]) for l in levels]


# --------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------
def interpolate_bilinear(grid: np.ndarray, query_points: np.ndarray, indexing: str = "ij") -> np.ndarray:
    """core_warp.py:42-150.  grid [B,H,W,C], query_points [B,N,2] -> [B,N,C]."""
    if indexing != "ij" and indexing != "xy":  # core_warp.py:64-65
        raise ValueError("Indexing mode must be 'ij' or 'xy'")
    grid = np.asarray(grid)
    query_points = np.asarray(query_points)
    if grid.ndim != 4:  # core_warp.py:71-73
        raise ValueError("Grid must be 4 dimensional. Received: " + str(grid.shape))
    if query_points.ndim != 3:  # core_warp.py:80-82
        raise ValueError("Query points must be 3 dimensional. Received: " + str(query_points.shape))
    batch_size, height, width, channels = grid.shape
    query_type = query_points.dtype
    grid_type = grid.dtype
    num_queries = query_points.shape[1]

    alphas, floors, ceils = [], [], []
    index_order = [0, 1] if indexing == "ij" else [1, 0]  # core_warp.py:90
    for dim in index_order:  # core_warp.py:93-119
        queries = query_points[:, :, dim]
        size_in_indexing_dimension = grid.shape[dim + 1]
        max_floor = np.asarray(size_in_indexing_dimension - 2).astype(query_type)  # :101
        min_floor = np.asarray(0.0).astype(query_type)
        floor = np.minimum(np.maximum(min_floor, np.floor(queries)), max_floor)  # :103-104
        int_floor = floor.astype(np.int32)  # :105
        floors.append(int_floor)
        ceils.append(int_floor + 1)  # :107
        alpha = (queries - floor).astype(grid_type)  # :112
        min_alpha = np.asarray(0.0).astype(grid_type)
        max_alpha = np.asarray(1.0).astype(grid_type)
        alpha = np.minimum(np.maximum(min_alpha, alpha), max_alpha)  # :115
        alphas.append(alpha[:, :, None])  # :119

    flattened_grid = grid.reshape(batch_size * height * width, channels)  # :122
    batch_offsets = (np.arange(batch_size, dtype=np.int32) * height * width).reshape(batch_size, 1)  # :124

    def gather(y_coords, x_coords):  # core_warp.py:131-136
        linear_coordinates = batch_offsets + y_coords * width + x_coords
        return flattened_grid[linear_coordinates].reshape(batch_size, num_queries, channels)

    top_left = gather(floors[0], floors[1])  # :139-142
    top_right = gather(floors[0], ceils[1])
    bottom_left = gather(ceils[0], floors[1])
    bottom_right = gather(ceils[0], ceils[1])

    interp_top = alphas[1] * (top_right - top_left) + top_left  # :146
    interp_bottom = alphas[1] * (bottom_right - bottom_left) + bottom_left  # :147
    interp = alphas[0] * (interp_bottom - interp_top) + interp_top  # :148
    return interp.astype(grid_type, copy=False)


def dense_image_warp(image: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """core_warp.py:153-202.  image [B,H,W,C], flow [B,H,W,2] -> [B,H,W,C].
    Channel 0 of `flow` displaces ROWS, the sign is minus (core_warp.py:189-194)."""
    image = np.asarray(image)
    flow = np.asarray(flow)
    if image.ndim != 4 or flow.ndim != 4:
        raise ValueError("image and flow must be 4 dimensional")
    batch_size, height, width, channels = image.shape
    grid_x, grid_y = np.meshgrid(np.arange(width), np.arange(height))  # :189-190
    stacked_grid = np.stack([grid_y, grid_x], axis=2).astype(flow.dtype)  # :191-192
    batched_grid = stacked_grid[None]  # :193
    query_points_on_grid = batched_grid - flow  # :194
    query_points_flattened = query_points_on_grid.reshape(batch_size, height * width, 2)  # :195-196
    interpolated = interpolate_bilinear(image, query_points_flattened)  # :199
    return interpolated.reshape(batch_size, height, width, channels)  # :200-201


def _leaky_relu(x: np.ndarray, alpha: float) -> np.ndarray:
    """TF 1.10 tf.nn.leaky_relu == maximum(alpha * x, x) with alpha cast to x.dtype."""
    a = np.asarray(alpha).astype(x.dtype)
    return np.maximum(a * x, x)


def cost_volume(c1: np.ndarray, warp: np.ndarray, search_range: int) -> np.ndarray:
    """core_costvol.py:20-40.  c1, warp [B,H,W,C] -> [B,H,W,(2r+1)^2], channel
    index (dy+r)*(2r+1)+(dx+r), dy outer (core_costvol.py:32-33)."""
    c1 = np.asarray(c1)
    warp = np.asarray(warp)
    r = int(search_range)
    padded_lvl = np.pad(warp, [[0, 0], [r, r], [r, r], [0, 0]])  # :27
    _, h, w, _ = c1.shape  # :28
    max_offset = r * 2 + 1  # :29
    cost_vol = []
    for y in range(0, max_offset):  # :32
        for x in range(0, max_offset):  # :33
            slc = padded_lvl[:, y:y + h, x:x + w, :]  # :34
            prod = c1 * slc
            # reduce_mean in the tensor dtype (numpy uses fp32 intermediates for fp16 input)
            cost = np.mean(prod, axis=3, keepdims=True).astype(c1.dtype, copy=False)  # :35
            cost_vol.append(cost)
    cost_vol = np.concatenate(cost_vol, axis=3)  # :37
    return _leaky_relu(cost_vol, 0.1)  # :38


def warp_cost_volume(c1, c2, flow, search_range: int, flow_scale: float = 1.0):
    """One pyramid level of tfoptflow/model_pwcnet.py:1550-1567:
    corr(c1, warp(c2, up_flow * scaler)) for levels with a flow, corr(c1, c2) at the top."""
    if flow is None:
        return cost_volume(c1, c2, search_range)
    flow = np.asarray(flow)
    scaled = flow * np.asarray(flow_scale).astype(flow.dtype)  # model_pwcnet.py:1560-1561
    return cost_volume(c1, dense_image_warp(c2, scaled), search_range)


# --------------------------------------------------------------------------
# backward (what TF 1.10 autodiff emits for the graph above)
# --------------------------------------------------------------------------
def cost_volume_bwd(g, out, c1, warp, search_range: int):
    """Gradients of `cost_volume` w.r.t. c1 and warp.  `out` is the forward result.
    maximum(0.1x, x) routes ties (x == 0) to the 0.1x branch (TF _MaximumMinimumGrad
    uses x >= y for the first argument)."""
    r = int(search_range)
    acc = np.float64 if np.asarray(c1).dtype == np.float64 else np.float32
    g = np.asarray(g).astype(acc)
    out = np.asarray(out)
    c1a = np.asarray(c1).astype(acc)
    wa = np.asarray(warp).astype(acc)
    B, H, W, C = c1a.shape
    n = 2 * r + 1
    gp = g * np.where(out > 0, acc(1.0), acc(0.1))
    wz = np.pad(wa, [[0, 0], [r, r], [r, r], [0, 0]])
    dc1 = np.zeros_like(c1a)
    dwz = np.zeros_like(wz)
    inv_c = acc(1.0) / acc(C)
    for y in range(n):
        for x in range(n):
            gd = gp[..., y * n + x][..., None] * inv_c
            dc1 += gd * wz[:, y:y + H, x:x + W, :]
            dwz[:, y:y + H, x:x + W, :] += gd * c1a
    dwarp = dwz[:, r:r + H, r:r + W, :]
    return dc1, dwarp


def dense_image_warp_bwd(gwarp, image, flow):
    """Gradients of `dense_image_warp` w.r.t. image and flow.
    floor() has no gradient; the alpha clamp min(max(0, a), 1) passes gradient only
    where 0 < a <= 1 (TF 1.x _MaximumMinimumGrad selectors), so exactly-integer
    coordinates get zero flow gradient."""
    image = np.asarray(image)
    flow = np.asarray(flow)
    acc = np.float64 if image.dtype == np.float64 else np.float32
    gw = np.asarray(gwarp).astype(acc)
    B, H, W, C = image.shape
    gy, gx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    grid = np.stack([gy, gx], axis=2).astype(flow.dtype)[None]
    q = grid - flow
    dimg = np.zeros((B * H * W, C), dtype=acc)
    fl, al, mask = [], [], []
    for dim, size in ((0, H), (1, W)):
        qs = q[..., dim]
        f = np.minimum(np.maximum(np.asarray(0.0, flow.dtype), np.floor(qs)), np.asarray(size - 2, flow.dtype))
        a_raw = (qs - f).astype(image.dtype)
        mask.append(((a_raw > 0) & (a_raw <= 1)).astype(acc))
        al.append(np.minimum(np.maximum(np.asarray(0.0, image.dtype), a_raw), np.asarray(1.0, image.dtype)).astype(acc))
        fl.append(f.astype(np.int64))
    ay, ax = al[0][..., None], al[1][..., None]
    boff = (np.arange(B, dtype=np.int64) * H * W).reshape(B, 1, 1)
    i_tl = boff + fl[0] * W + fl[1]
    i_tr = i_tl + 1
    i_bl = i_tl + W
    i_br = i_bl + 1
    flat = image.reshape(B * H * W, C).astype(acc)
    tl, tr, bl, br = flat[i_tl], flat[i_tr], flat[i_bl], flat[i_br]
    # out = ay*(bot-top)+top ; top = ax*(tr-tl)+tl ; bot = ax*(br-bl)+bl
    w_tl = (1 - ay) * (1 - ax)
    w_tr = (1 - ay) * ax
    w_bl = ay * (1 - ax)
    w_br = ay * ax
    for idx, wgt in ((i_tl, w_tl), (i_tr, w_tr), (i_bl, w_bl), (i_br, w_br)):
        np.add.at(dimg, idx.reshape(-1), (gw * wgt).reshape(-1, C))
    top = ax * (tr - tl) + tl
    bot = ax * (br - bl) + bl
    d_ax = np.sum(gw * ((1 - ay) * (tr - tl) + ay * (br - bl)), axis=-1)
    d_ay = np.sum(gw * (bot - top), axis=-1)
    dflow = np.stack([-d_ay * mask[0], -d_ax * mask[1]], axis=-1)
    return dimg.reshape(B, H, W, C), dflow


def warp_cost_volume_bwd(g, c1, c2, flow, search_range: int, flow_scale: float = 1.0):
    """Backward of `warp_cost_volume`: returns (dc1, dc2, dflow); dflow is None when flow is None."""
    if flow is None:
        out = cost_volume(c1, c2, search_range)
        dc1, dc2 = cost_volume_bwd(g, out, c1, c2, search_range)
        return dc1, dc2, None
    flow = np.asarray(flow)
    scaled = flow * np.asarray(flow_scale).astype(flow.dtype)
    warp = dense_image_warp(c2, scaled)
    out = cost_volume(c1, warp, search_range)
    dc1, dwarp = cost_volume_bwd(g, out, c1, warp, search_range)
    dc2, dscaled = dense_image_warp_bwd(dwarp, c2, scaled)
    return dc1, dc2, dscaled * dscaled.dtype.type(flow_scale)
