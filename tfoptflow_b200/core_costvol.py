"""Drop-in for tfoptflow/core_costvol.py: `cost_volume(c1, warp, search_range, name)`
(reference: tfoptflow/core_costvol.py:20-40), backed by sm_100a CUDA kernels."""
from __future__ import annotations

from . import ops
from .core_warp import WarpedTensor


def cost_volume(c1, warp, search_range, name=None, out=None):
    """Build cost volume for associating a pixel from Image1 with its corresponding pixels in Image2.
    Args:
        c1: Level of the feature pyramid of Image1, [B,H,W,C]
        warp: Warped level of the feature pyramid of Image2 (a tensor, or the lazy result of
              `dense_image_warp`, in which case warp and correlation run as one kernel)
        search_range: Search range (maximum displacement)
    Returns [B,H,W,(2r+1)^2], leaky-ReLU(0.1) applied, channel (dy+r)(2r+1)+(dx+r).
    `out` (extension): write into a channel slice of a wider NHWC buffer instead of allocating."""
    if isinstance(warp, WarpedTensor) and warp._value is None:
        return ops.warp_cost_volume(c1, warp.image, warp.flow, search_range, warp.flow_scale, out=out)
    if isinstance(warp, WarpedTensor):
        warp = warp.materialize()
    return ops.warp_cost_volume(c1, warp, None, search_range, 1.0, out=out)
