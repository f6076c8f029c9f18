"""torch.autograd wrappers around the C ABI.  Shapes are NHWC like the reference."""
from __future__ import annotations

import torch

from . import _ffi
from ._tensors import as_cuda_tensor, dtype_code, ptr, stream_ptr


def _check4(t: torch.Tensor, what: str):
    if t.dim() != 4:
        raise ValueError(f"{what} must be 4 dimensional. Received: {list(t.shape)}")


def _ws(nbytes: int, device) -> torch.Tensor | None:
    return torch.empty(nbytes, dtype=torch.uint8, device=device) if nbytes else None


def warp_cost_volume_forward(c1, c2, flow, search_range: int, flow_scale: float = 1.0, out=None):
    """cost_volume(c1, dense_image_warp(c2, flow * flow_scale), r) in one kernel; flow=None: no warp.
    `out` may be a [B,H,W,>=D] channel-slice view of a wider NHWC buffer (the concat buffer of
    tfoptflow/model_pwcnet.py:1424)."""
    lib = _ffi.load()
    _check4(c1, "c1"), _check4(c2, "c2")
    if c1.shape != c2.shape or c1.dtype != c2.dtype:
        raise ValueError(f"c1 {list(c1.shape)}/{c1.dtype} and c2 {list(c2.shape)}/{c2.dtype} must match")
    B, H, W, C = c1.shape
    r = int(search_range)
    D = (2 * r + 1) ** 2
    if flow is not None:
        _check4(flow, "flow")
        if tuple(flow.shape) != (B, H, W, 2):
            raise ValueError(f"flow must be [B,H,W,2]={[B, H, W, 2]}. Received: {list(flow.shape)}")
    if out is None:
        out = torch.empty((B, H, W, D), dtype=c1.dtype, device=c1.device)
        ops = D
    else:
        if out.dim() != 4 or tuple(out.shape[:3]) != (B, H, W) or out.shape[3] != D or out.dtype != c1.dtype:
            raise ValueError("out must be a [B,H,W,(2r+1)^2] view with the feature dtype")
        if out.stride(3) != 1 or out.stride(1) != W * out.stride(2) or out.stride(0) != H * W * out.stride(2):
            raise ValueError("out must be a channel slice of a dense NHWC buffer")
        ops = out.stride(2)
    with torch.cuda.device(c1.device):
        _ffi.check(lib.pwc_warp_cost_volume_fwd(
            ptr(c1), ptr(c2), ptr(flow), float(flow_scale), ptr(out), B, H, W, C, r,
            dtype_code(c1), dtype_code(flow) if flow is not None else dtype_code(c1), ops, stream_ptr(c1.device)))
    return out


class _WarpCostVolumeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c1, c2, flow, search_range, flow_scale):
        c1, c2 = c1.contiguous(), c2.contiguous()
        flow = None if flow is None else flow.contiguous()
        out = warp_cost_volume_forward(c1, c2, flow, search_range, flow_scale)
        ctx.save_for_backward(c1, c2, flow, out)
        ctx.r, ctx.flow_scale = int(search_range), float(flow_scale)
        return out

    @staticmethod
    def backward(ctx, g):
        c1, c2, flow, out = ctx.saved_tensors
        lib = _ffi.load()
        B, H, W, C = c1.shape
        g = g.contiguous()
        if g.dtype != c1.dtype:
            g = g.to(c1.dtype)
        dc1, dc2 = torch.empty_like(c1), torch.empty_like(c2)
        dflow = torch.empty_like(flow) if flow is not None else None
        nb = lib.pwc_bwd_workspace_bytes(B, H, W, C, ctx.r, dtype_code(c1), 1 if flow is not None else 0)
        ws = _ws(nb, c1.device)
        with torch.cuda.device(c1.device):
            _ffi.check(lib.pwc_warp_cost_volume_bwd(
                ptr(g), ptr(out), ptr(c1), ptr(c2), ptr(flow), ctx.flow_scale, ptr(dc1), ptr(dc2), ptr(dflow),
                B, H, W, C, ctx.r, dtype_code(c1), dtype_code(flow) if flow is not None else dtype_code(c1),
                0, 0, ptr(ws), nb, stream_ptr(c1.device)))
        return dc1, dc2, dflow, None, None


class _DenseImageWarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, flow):
        lib = _ffi.load()
        image, flow = image.contiguous(), flow.contiguous()
        B, H, W, C = image.shape
        out = torch.empty_like(image)
        with torch.cuda.device(image.device):
            _ffi.check(lib.pwc_dense_image_warp_fwd(ptr(image), ptr(flow), ptr(out), B, H, W, C,
                                                    dtype_code(image), dtype_code(flow), stream_ptr(image.device)))
        ctx.save_for_backward(image, flow)
        return out

    @staticmethod
    def backward(ctx, g):
        image, flow = ctx.saved_tensors
        lib = _ffi.load()
        B, H, W, C = image.shape
        g = g.contiguous().to(image.dtype)
        dimg, dflow = torch.empty_like(image), torch.empty_like(flow)
        nb = lib.pwc_bwd_workspace_bytes(B, H, W, C, 0, dtype_code(image), 1)
        ws = _ws(nb, image.device)
        with torch.cuda.device(image.device):
            _ffi.check(lib.pwc_dense_image_warp_bwd(ptr(g), ptr(image), ptr(flow), ptr(dimg), ptr(dflow), B, H, W, C,
                                                    dtype_code(image), dtype_code(flow), ptr(ws), nb,
                                                    stream_ptr(image.device)))
        return dimg, dflow


def warp_cost_volume(c1, c2, flow, search_range: int, flow_scale: float = 1.0, out=None):
    """Differentiable fused op.  With `out=` (inference: write into a concat buffer) no graph is recorded."""
    c1, c2 = as_cuda_tensor(c1, "c1"), as_cuda_tensor(c2, "c2")
    flow = None if flow is None else as_cuda_tensor(flow, "flow")
    if out is not None:
        return warp_cost_volume_forward(c1, c2, flow, search_range, flow_scale, out=out)
    return _WarpCostVolumeFn.apply(c1, c2, flow, int(search_range), float(flow_scale))


def dense_image_warp_eager(image, flow):
    image, flow = as_cuda_tensor(image, "image"), as_cuda_tensor(flow, "flow")
    _check4(image, "image"), _check4(flow, "flow")
    B, H, W, C = image.shape
    if tuple(flow.shape) != (B, H, W, 2):
        raise ValueError(f"flow must be [B,H,W,2]={[B, H, W, 2]}. Received: {list(flow.shape)}")
    if H < 2 or W < 2:
        raise ValueError("dense_image_warp needs height >= 2 and width >= 2")
    return _DenseImageWarpFn.apply(image, flow)
