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
        for i, L in enumerate(levels):
            c1, c2 = as_cuda_tensor(L["c1"], "c1"), as_cuda_tensor(L["c2"], "c2")
            flow = None if L.get("flow") is None else as_cuda_tensor(L["flow"], "flow")
            B, H, W, C = c1.shape
            out = outs[i] if outs is not None else torch.empty((B, H, W, D), dtype=c1.dtype, device=c1.device)
            self.tensors.append((c1, c2, flow))
            self.outs.append(out)
            descs.append(dict(c1=ptr(c1), c2=ptr(c2), flow=ptr(flow), out=ptr(out), H=H, W=W, C=C,
                              flow_scale=float(L.get("scaler", 1.0)), out_pixel_stride=out.stride(2)))
        self.B = self.tensors[0][0].shape[0]
        self.device = self.tensors[0][0].device
        self.dtype = dtype_code(self.tensors[0][0])
        flows = [t[2] for t in self.tensors if t[2] is not None]
        self.flow_dtype = dtype_code(flows[0]) if flows else self.dtype
        self.arr = _levels_array(descs)
        self.n = len(descs)
        self.lib = _ffi.load()

    def run(self):
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
        for L, o in zip(levels, outs):
            B, H, W, C = L["c1"].shape
            f = L.get("flow")
            for name, arr in (("c1", L["c1"]), ("c2", L["c2"]), ("flow", f), ("out", o)):
                if arr is not None and not (isinstance(arr, np.ndarray) and arr.flags["C_CONTIGUOUS"]):
                    raise ValueError(f"{name} must be a C-contiguous numpy array (NHWC)")
            if L["c2"].shape != L["c1"].shape or (f is not None and f.shape != (B, H, W, 2)) or \
                    o.shape != (B, H, W, (2 * int(search_range) + 1) ** 2) or o.dtype != L["c1"].dtype:
                raise ValueError("c1/c2/flow/out shapes or dtypes do not match")
            descs.append(dict(c1=L["c1"].ctypes.data, c2=L["c2"].ctypes.data, flow=None if f is None else f.ctypes.data,
                              out=o.ctypes.data, H=H, W=W, C=C, flow_scale=float(L.get("scaler", 1.0)),
                              out_pixel_stride=0))
        dt = self._NP[levels[0]["c1"].dtype]
        arr = _levels_array(descs)
        _ffi.check(self.lib.pwc_pyramid_fwd_host(self.h, arr, len(descs), levels[0]["c1"].shape[0], int(search_range),
                                                 dt, dt))
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
