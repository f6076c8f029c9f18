"""tfoptflow_b200 - B200-native (sm_100a) implementation of PWC-Net's per-pyramid-level warp + cost-volume
hot path, behind the reference's own signatures (philferriere/tfoptflow: core_costvol.cost_volume,
core_warp.dense_image_warp).  PyTorch supplies device memory, streams and torch.distributed only; the
compute is hand-written CUDA reached through a C ABI (include/pwc_b200.h)."""
from .core_costvol import cost_volume  # noqa: F401
from .core_warp import WarpedTensor, dense_image_warp  # noqa: F401
from .ops import warp_cost_volume  # noqa: F401

__all__ = ["cost_volume", "dense_image_warp", "warp_cost_volume", "WarpedTensor"]
