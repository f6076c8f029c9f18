"""Op-for-op torch-CPU port of the reference's TensorFlow graph for the hot path.
TEST / BASELINE INFRASTRUCTURE ONLY (same rules as oracle/pwc_oracle.py).

TensorFlow 1.10.1 is not installable here, so the "reference CPU path" that
bench.py times is this port: it keeps the reference's graph *structure* - one
pad, (2r+1)^2 x {slice, multiply, reduce_mean}, one concat, one leaky-relu
(tfoptflow/core_costvol.py:27-38); a meshgrid, four flat gathers and three lerps
(tfoptflow/core_warp.py:86-148,185-202) - with each TF op replaced by the
matching multi-threaded torch CPU kernel.  It is labelled "port", never "TF".

It is also a second, independently written implementation used to cross-check
oracle/pwc_oracle.py (tests/test_oracle.py) and, through autograd, its backward
at non-integer flows (torch.clamp and TF's Maximum/Minimum gradients differ only
at exact ties).
"""
from __future__ import annotations

import torch

__all__ = ["dense_image_warp", "cost_volume", "pyramid_forward"]


def _interpolate_bilinear(grid: torch.Tensor, query_points: torch.Tensor) -> torch.Tensor:
    # tfoptflow/core_warp.py:42-150, indexing='ij'
    if grid.dim() != 4:
        raise ValueError("Grid must be 4 dimensional. Received: " + str(tuple(grid.shape)))
    if query_points.dim() != 3:
        raise ValueError("Query points must be 3 dimensional. Received: " + str(tuple(query_points.shape)))
    batch_size, height, width, channels = grid.shape
    num_queries = query_points.shape[1]
    alphas, floors, ceils = [], [], []
    for dim in (0, 1):
        queries = query_points[:, :, dim]
        size = grid.shape[dim + 1]
        max_floor = torch.tensor(size - 2, dtype=query_points.dtype)
        min_floor = torch.tensor(0.0, dtype=query_points.dtype)
        floor = torch.minimum(torch.maximum(min_floor, torch.floor(queries)), max_floor)
        int_floor = floor.to(torch.int32)
        floors.append(int_floor)
        ceils.append(int_floor + 1)
        alpha = (queries - floor.detach()).to(grid.dtype)
        alpha = torch.minimum(torch.maximum(torch.tensor(0.0, dtype=grid.dtype), alpha),
                              torch.tensor(1.0, dtype=grid.dtype))
        alphas.append(alpha.unsqueeze(2))
    flattened_grid = grid.reshape(batch_size * height * width, channels)
    batch_offsets = (torch.arange(batch_size, dtype=torch.int32) * height * width).reshape(batch_size, 1)

    def gather(y_coords, x_coords):
        linear_coordinates = (batch_offsets + y_coords * width + x_coords).long()
        return flattened_grid[linear_coordinates.reshape(-1)].reshape(batch_size, num_queries, channels)

    top_left = gather(floors[0], floors[1])
    top_right = gather(floors[0], ceils[1])
    bottom_left = gather(ceils[0], floors[1])
    bottom_right = gather(ceils[0], ceils[1])
    interp_top = alphas[1] * (top_right - top_left) + top_left
    interp_bottom = alphas[1] * (bottom_right - bottom_left) + bottom_left
    return alphas[0] * (interp_bottom - interp_top) + interp_top


def dense_image_warp(image: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    # tfoptflow/core_warp.py:153-202
    batch_size, height, width, channels = image.shape
    grid_y, grid_x = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
    stacked_grid = torch.stack([grid_y, grid_x], dim=2).to(flow.dtype)
    query_points_on_grid = stacked_grid.unsqueeze(0) - flow
    query_points_flattened = query_points_on_grid.reshape(batch_size, height * width, 2)
    interpolated = _interpolate_bilinear(image, query_points_flattened)
    return interpolated.reshape(batch_size, height, width, channels)


def cost_volume(c1: torch.Tensor, warp: torch.Tensor, search_range: int) -> torch.Tensor:
    # tfoptflow/core_costvol.py:20-40
    r = int(search_range)
    padded_lvl = torch.nn.functional.pad(warp, (0, 0, r, r, r, r))
    _, h, w, _ = c1.shape
    max_offset = r * 2 + 1
    cost_vol = []
    for y in range(0, max_offset):
        for x in range(0, max_offset):
            slc = padded_lvl[:, y:y + h, x:x + w, :]
            cost = torch.mean(c1 * slc, dim=3, keepdim=True)
            cost_vol.append(cost)
    cost_vol = torch.cat(cost_vol, dim=3)
    return torch.maximum(0.1 * cost_vol, cost_vol)  # TF 1.10 leaky_relu


def pyramid_forward(levels, search_range: int):
    """One image-pair batch through the pyramid loop's hot path
    (tfoptflow/model_pwcnet.py:1550-1567): `levels` is a list of
    (c1, c2, flow_or_None, scaler) from the top level down."""
    outs = []
    for c1, c2, flow, scaler in levels:
        if flow is None:
            outs.append(cost_volume(c1, c2, search_range))
        else:
            outs.append(cost_volume(c1, dense_image_warp(c2, flow * scaler), search_range))
    return outs
