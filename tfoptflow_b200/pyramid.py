"""The hot-path calls of PWC-Net's coarse-to-fine loop (tfoptflow/model_pwcnet.py:1550-1567) and the two
thin wrappers around them (`ModelPWCNet.warp` :1106-1178, `ModelPWCNet.corr` :1226-1277)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _ffi, ops
from ._tensors import as_cuda_tensor, dtype_code, ptr, stream_ptr
from .core_costvol import cost_volume
from .core_warp import WarpedTensor, dense_image_warp

DEFAULT_OPTS = {"pyr_lvls": 6, "flow_pred_lvl": 2, "search_range": 4}  # model_pwcnet.py:33-62


def flow_scaler(lvl: int) -> float:
    """`scaler = 20. / 2**lvl` (model_pwcnet.py:1560): 0.625, 1.25, 2.5, 5.0 for levels 5..2."""
    return 20.0 / 2 ** lvl


def warp(c2, sc_up_flow, lvl, name="warp"):
    """ModelPWCNet.warp (model_pwcnet.py:1106-1178): dense_image_warp(c2, sc_up_flow, name=f'warp{lvl}')."""
    return dense_image_warp(c2, sc_up_flow, name=f"{name}{lvl}")


def corr(c1, warp_, lvl, search_range=4, name="corr", out=None):
    """ModelPWCNet.corr (model_pwcnet.py:1226-1277): cost_volume(c1, warp, search_range, f'corr{lvl}')."""
    return cost_volume(c1, warp_, search_range, f"{name}{lvl}", out=out)


def level_cost_volume(c1, c2, up_flow, lvl, opts=None, out=None):
    """One iteration of the loop body's hot path: top level correlates un-warped features
    (model_pwcnet.py:1552-1554), lower levels warp c2 by up_flow * 20/2^lvl first (:1560-1564).
    The scaler multiply is folded into the kernel."""
    opts = opts or DEFAULT_OPTS
    if up_flow is None:
        return corr(c1, c2, lvl, opts["search_range"], out=out)
    w = WarpedTensor(as_cuda_tensor(c2, "c2"), as_cuda_tensor(up_flow, "up_flow"), flow_scaler(lvl), f"warp{lvl}")
    return corr(c1, w, lvl, opts["search_range"], out=out)


def _levels_array(levels, dev_ptrs=True):
    arr = (_ffi.LevelT * len(levels))()
    for i, L in enumerate(levels):
        arr[i].c1, arr[i].c2, arr[i].flow, arr[i].out = L["c1"], L["c2"], L["flow"], L["out"]
        arr[i].H, arr[i].W, arr[i].C = L["H"], L["W"], L["C"]
        arr[i].flow_scale = L["flow_scale"]
        arr[i].out_pixel_stride = L.get("out_pixel_stride", 0)
    return arr


class PyramidPlan:
    """Pre-validated, pre-marshalled argument block for `pwc_pyramid_fwd`: all hot-path kernels of one
    forward pass in ONE C call (inference; no autograd).  `levels` is a list of dicts
    (c1, c2, flow|None, scaler) from the top of the pyramid down; outputs are allocated once."""

    def __init__(self, levels, search_range=4, outs=None):
        self.r = int(search_range)
        D = (2 * self.r + 1) ** 2
        self.tensors = []
        descs = []
        self.outs = []
        if outs is not None and len(outs) != len(levels):
            raise ValueError("outs must have one tensor per level")
        for i, L in enumerate(levels):
            c1, c2 = as_cuda_tensor(L["c1"], "c1"), as_cuda_tensor(L["c2"], "c2")
            flow = None if L.get("flow") is None else as_cuda_tensor(L["flow"], "flow")
            if c1.dim() != 4 or c2.shape != c1.shape or c2.dtype != c1.dtype or c2.device != c1.device:
                raise ValueError(f"level {i}: c1 and c2 must be [B,H,W,C] tensors of one dtype on one device")
            B, H, W, C = c1.shape
            if flow is not None and (tuple(flow.shape) != (B, H, W, 2) or flow.device != c1.device):
                raise ValueError(f"level {i}: flow must be [B,H,W,2]={[B, H, W, 2]} on the features' device")
            if outs is not None:
                out = outs[i]
                if out.dim() != 4 or tuple(out.shape) != (B, H, W, D) or out.dtype != c1.dtype or out.device != c1.device:
                    raise ValueError(f"level {i}: out must be a [B,H,W,{D}] view with the feature dtype and device")
                if out.stride(3) != 1 or out.stride(1) != W * out.stride(2) or out.stride(0) != H * W * out.stride(2):
                    raise ValueError(f"level {i}: out must be a channel slice of a dense NHWC buffer")
            else:
                out = torch.empty((B, H, W, D), dtype=c1.dtype, device=c1.device)
            self.tensors.append((c1, c2, flow))
            self.outs.append(out)
            descs.append(dict(c1=ptr(c1), c2=ptr(c2), flow=ptr(flow), out=ptr(out), H=H, W=W, C=C,
                              flow_scale=float(L.get("scaler", 1.0)), out_pixel_stride=out.stride(2)))
        first = self.tensors[0][0]
        self.B, self.device, self.dtype = first.shape[0], first.device, dtype_code(first)
        flows = [t[2] for t in self.tensors if t[2] is not None]
        self.flow_dtype = dtype_code(flows[0]) if flows else self.dtype
        for c1, _, flow in self.tensors:  # one (B, dtype, flow dtype, device) per call: the C entry point takes them once
            if c1.shape[0] != self.B or dtype_code(c1) != self.dtype or c1.device != self.device or \
                    (flow is not None and dtype_code(flow) != self.flow_dtype):
                raise ValueError("all levels of a plan must share batch size, dtypes and device")
        self.arr = _levels_array(descs)
        self.n = len(descs)
        self.lib = _ffi.load()

    def run(self):
        with torch.cuda.device(self.device):  # the library launches on the calling thread's current device
            _ffi.check(self.lib.pwc_pyramid_fwd(self.arr, self.n, self.B, self.r, self.dtype, self.flow_dtype,
                                                stream_ptr(self.device)))
        return self.outs


class HostPyramid:
    """End-to-end path with HOST buffers (numpy arrays, ideally pinned): H2D copies, kernels and D2H copies
    are issued by `pwc_pyramid_fwd_host` on the library's own streams and overlap across levels."""

    _NP = {np.dtype(np.float32): _ffi.PWC_F32, np.dtype(np.float16): _ffi.PWC_F16}

    def __init__(self):
        self.lib = _ffi.load()
        h = ctypes.c_void_p()
        _ffi.check(self.lib.pwc_host_ctx_create(ctypes.byref(h)))
        self.h = h

    def run(self, levels, outs, search_range=4):
        descs = []
        B0, dt0, fdt0 = levels[0]["c1"].shape[0], levels[0]["c1"].dtype, None
        if dt0 not in self._NP:
            raise ValueError(f"unsupported feature dtype {dt0} (float32 / float16)")
        for L, o in zip(levels, outs):
            B, H, W, C = L["c1"].shape
            f = L.get("flow")
            for name, arr in (("c1", L["c1"]), ("c2", L["c2"]), ("flow", f), ("out", o)):
                if arr is not None and not (isinstance(arr, np.ndarray) and arr.flags["C_CONTIGUOUS"]):
                    raise ValueError(f"{name} must be a C-contiguous numpy array (NHWC)")
            if L["c2"].shape != L["c1"].shape or (f is not None and f.shape != (B, H, W, 2)) or \
                    o.shape != (B, H, W, (2 * int(search_range) + 1) ** 2):
                raise ValueError("c1/c2/flow/out shapes do not match")
            if B != B0 or L["c1"].dtype != dt0 or L["c2"].dtype != dt0 or o.dtype != dt0:
                raise ValueError("every level must share the batch size, and c1/c2/out the feature dtype")
            if f is not None:
                if f.dtype not in self._NP:
                    raise ValueError(f"unsupported flow dtype {f.dtype} (float32 / float16)")
                if fdt0 is None:
                    fdt0 = f.dtype
                elif f.dtype != fdt0:
                    raise ValueError("all flows of one call must share a dtype")
            descs.append(dict(c1=L["c1"].ctypes.data, c2=L["c2"].ctypes.data, flow=None if f is None else f.ctypes.data,
                              out=o.ctypes.data, H=H, W=W, C=C, flow_scale=float(L.get("scaler", 1.0)),
                              out_pixel_stride=0))
        arr = _levels_array(descs)
        _ffi.check(self.lib.pwc_pyramid_fwd_host(self.h, arr, len(descs), B0, int(search_range), self._NP[dt0],
                                                 self._NP[fdt0 if fdt0 is not None else dt0]))
        return outs

    def close(self):
        if self.h:
            self.lib.pwc_host_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hot_path_forward(c1_pyr, c2_pyr, up_flows, opts=None):
    """Runs the loop's hot-path calls for externally supplied up-sampled flows (differentiable):
    returns {lvl: cost volume}.  `c1_pyr`/`c2_pyr`/`up_flows` are dicts keyed by level."""
    opts = opts or DEFAULT_OPTS
    out = {}
    for lvl in range(opts["pyr_lvls"], opts["flow_pred_lvl"] - 1, -1):
        uf = None if lvl == opts["pyr_lvls"] else up_flows[lvl]
        out[lvl] = level_cost_volume(c1_pyr[lvl], c2_pyr[lvl], uf, lvl, opts)
    return out
